"""Pins the oracle (oracle/broker_oracle.cpp) against every known-answer test the reference holds
for the hot path.  CPU only.  If these pass, the oracle reproduces the reference's
routing outcomes, table semantics and framing; the GPU tests then compare the CUDA path with it.
"""
import ctypes as C

import pytest

import scenarios
from harness import OracleBackend
from oracle import oracle as orc


@pytest.mark.parametrize("scenario", scenarios.ALL, ids=lambda f: f.__name__)
def test_reference_scenario_on_oracle(scenario):
    scenario(OracleBackend())


# ---- relational_map.rs:132-346 ---------------------------------------------------------------
class Rel:
    def __init__(self):
        self.L = orc.lib()
        self.h = self.L.orc_rel_new()

    def assoc(self, k, vs):
        self.L.orc_rel_assoc(self.h, k.encode(), (C.c_uint64 * max(1, len(vs)))(*vs), len(vs))

    def dissoc(self, k, vs):
        self.L.orc_rel_dissoc(self.h, k.encode(), (C.c_uint64 * max(1, len(vs)))(*vs), len(vs))

    def remove_key(self, k):
        self.L.orc_rel_remove_key(self.h, k.encode())

    def keys(self, v):
        buf = C.create_string_buffer(4096)
        n = self.L.orc_rel_keys_by_value(self.h, v, buf, 4096)
        return buf.value.decode().split("\n")[:-1] if n else []

    def values(self):
        out = (C.c_uint64 * 64)()
        n = self.L.orc_rel_values(self.h, out, 64)
        return [out[i] for i in range(n)]

    def values_of(self, k):
        out = (C.c_uint64 * 64)()
        n = self.L.orc_rel_values_of_key(self.h, k.encode(), out, 64)
        return None if n == 0xFFFFFFFF else [out[i] for i in range(n)]

    def nkeys(self):
        return self.L.orc_rel_num_keys(self.h)

    def nvalues(self):
        return self.L.orc_rel_num_values(self.h)


def test_relational():
    """relational_map.rs:132-207 test_relational"""
    m = Rel()
    m.assoc("user0", [0, 1, 2])
    m.assoc("user1", [1, 2])
    assert m.keys(0) == ["user0"]
    assert m.keys(1) == ["user0", "user1"]
    m.dissoc("user0", [1])
    assert m.keys(1) == ["user1"]
    m.remove_key("user1")
    assert m.keys(2) == ["user0"]
    m.dissoc("user0", [2])
    assert m.keys(1) == []
    assert 1 not in m.values()
    assert m.nkeys() == 1 and m.nvalues() == 1
    m.dissoc("user0", [0])
    assert m.nkeys() == 0 and m.nvalues() == 0


def test_relational_association():
    """relational_map.rs:210-280 test_relational_association"""
    m = Rel()
    m.assoc("user0", [0, 1, 2])
    m.assoc("user1", [1, 2])
    m.dissoc("user0", [1])
    assert m.keys(0) == ["user0"] and m.keys(1) == ["user1"] and m.keys(2) == ["user0", "user1"]
    assert m.values_of("user1") == [1, 2] and m.values_of("user0") == [0, 2]
    assert len(m.values()) == 3 and m.nkeys() == 2
    m.dissoc("user0", [0])
    assert m.keys(0) == [] and m.keys(1) == ["user1"] and m.keys(2) == ["user0", "user1"]
    assert m.values_of("user1") == [1, 2] and m.values_of("user0") == [2]
    assert m.values() == [1, 2]
    m.dissoc("user1", [1, 2])
    assert m.keys(0) == [] and m.keys(1) == [] and m.keys(2) == ["user0"]
    assert m.values_of("user1") is None and m.values_of("user0") == [2]


def test_relational_remove():
    """relational_map.rs:283-346 test_relational_remove"""
    m = Rel()
    m.assoc("user0", [0, 1, 2])
    m.assoc("user1", [1, 2, 3])
    assert m.values_of("user0") == [0, 1, 2] and m.values_of("user1") == [1, 2, 3]
    assert m.keys(0) == ["user0"] and m.keys(1) == ["user0", "user1"]
    assert m.keys(2) == ["user0", "user1"] and m.keys(3) == ["user1"]
    m.remove_key("user1")
    assert m.values_of("user0") == [0, 1, 2] and m.values_of("user1") is None
    assert m.keys(0) == ["user0"] and m.keys(1) == ["user0"] and m.keys(2) == ["user0"] and m.keys(3) == []
    assert 3 not in m.values()
    m.remove_key("user0")
    assert m.values_of("user0") is None and m.nvalues() == 0


# ---- versioned_map.rs:277-376 ----------------------------------------------------------------
class Ver:
    def __init__(self, ident=0, h=None):
        self.L = orc.lib()
        self.h = h if h is not None else self.L.orc_ver_new(ident)

    def insert(self, k, v):
        self.L.orc_ver_insert(self.h, k.encode(), v.encode())

    def remove(self, k):
        self.L.orc_ver_remove(self.h, k.encode())

    def get(self, k):
        buf = C.create_string_buffer(256)
        return buf.value.decode() if self.L.orc_ver_get(self.h, k.encode(), buf, 256) else None

    def get_full(self):
        return Ver(h=self.L.orc_ver_get_full(self.h))

    def diff(self):
        return Ver(h=self.L.orc_ver_diff(self.h))

    def merge(self, other):
        return self.L.orc_ver_merge(self.h, other.h)

    def purge(self, v):
        self.L.orc_ver_remove_by_value_no_modify(self.h, v.encode())

    def __len__(self):
        return self.L.orc_ver_len(self.h)


def test_versioned_insert_remove():
    """versioned_map.rs:277-289"""
    m = Ver(0)
    m.insert("user0", "broker0")
    assert m.get("user0") == "broker0"
    m.remove("user0")
    assert m.get("user0") is None


def test_versioned_conflict():
    """versioned_map.rs:292-308 — equal versions: the higher conflict identity wins"""
    m0, m1 = Ver(0), Ver(1)
    m0.insert("user0", "broker0")
    m1.insert("user0", "broker1")
    m0.merge(m1.get_full())
    m1.merge(m0.get_full())
    assert m0.get("user0") == "broker1" and m1.get("user0") == "broker1"


def test_versioned_partial():
    """versioned_map.rs:311-344"""
    m0, m1 = Ver(0), Ver(1)
    m0.insert("user0", "broker0")
    m0.diff()
    m0.insert("user1", "broker0")
    d = m0.diff()
    m1.merge(d)
    assert m1.get("user0") is None and m1.get("user1") == "broker0"
    m1.merge(m0.get_full())
    assert m1.get("user0") == "broker0"
    m1.remove("user0")
    m0.merge(m1.diff())
    assert m0.get("user0") is None


def test_versioned_purge():
    """versioned_map.rs:347-376"""
    m = Ver(0)
    m.insert("user0", "broker0")
    m.insert("user1", "broker0")
    m.insert("user2", "broker1")
    m.purge("broker0")
    assert m.get("user0") is None and m.get("user1") is None and m.get("user2") == "broker1"
    assert len(m.diff()) == 1


# ---- connections/mod.rs:410-526 --------------------------------------------------------------
def _pair():
    local = orc.Oracle("test/local")
    remote = orc.Oracle("test/remote")
    local.add_broker("test/remote")
    remote.add_broker("test/local")
    return local, remote


def test_topic_sync():
    """connections/mod.rs:410-470 test_topic_sync"""
    local, remote = _pair()
    remote.subscribe_user_to(bytes([1]), [0, 1])
    assert not remote.topic_sync_to(local, "test/remote", full=True, apply=False)  # full sync is None
    assert remote.topic_sync_to(local, "test/remote")
    rb = local.broker_conn("test/remote")
    assert local.interested([0]) == [rb] and local.interested_counts([0]) == (1, 0)
    assert local.interested([1]) == [rb]
    remote.unsubscribe_user_from(bytes([1]), [0])
    assert remote.topic_sync_to(local, "test/remote")
    assert local.interested([0]) == []
    assert local.interested([1]) == [rb]


def test_topic_sync_out_of_order():
    """connections/mod.rs:474-526 test_topic_sync_out_of_order"""
    local, remote = _pair()
    remote.subscribe_user_to(bytes([1]), [0, 1])
    assert remote.topic_sync_to(local, "test/remote", apply=False)  # computed, not applied
    remote.unsubscribe_user_from(bytes([1]), [0])
    remote.unsubscribe_user_from(bytes([1]), [1])
    assert remote.topic_sync_to(local, "test/remote")
    remote.subscribe_user_to(bytes([1]), [1])
    assert remote.topic_sync_to(local, "test/remote")
    assert remote.topic_sync_to(local, "test/remote", full=True)
    rb = local.broker_conn("test/remote")
    assert local.interested([0]) == []
    assert local.interested([1]) == [rb]


# ---- framing + wire (cdn-proto/src/connection/protocols/mod.rs:354-394, message.rs:397-457) ----
def test_frame_is_be_length_then_raw():
    o = orc.Oracle("/")
    c = o.add_user(b"k" * 8, [0])
    raw = orc.broadcast_frame([0], b"x" * 21)
    o.handle_broadcast_message([0], raw)
    assert o.stream(c) == len(raw).to_bytes(4, "big") + raw
    assert o.bytes_sent() == len(raw) and o.deliveries() == 1


def test_serialization_parity():
    """message.rs:397-457 test_serialization_parity (routed kinds): serialize → deserialize"""
    cases = [
        (orc.KIND_DIRECT, bytes(range(3)), bytes([3, 4, 5])),
        (orc.KIND_BROADCAST, bytes([0, 1]), bytes([0, 1, 2])),
        (orc.KIND_SUBSCRIBE, bytes([0, 1]), b""),
        (orc.KIND_UNSUBSCRIBE, bytes([0, 1]), b""),
        (orc.KIND_USER_SYNC, b"", bytes([0, 1, 2])),
        (orc.KIND_TOPIC_SYNC, b"", bytes([0, 1, 2])),
        (orc.KIND_BROADCAST, b"", b""),
        (orc.KIND_BROADCAST, bytes([7]), bytes(10000)),       # benches/broadcast.rs:26: 2-segment
        (orc.KIND_DIRECT, bytes(128), bytes(range(256)) * 40),
    ]
    for kind, f0, pl in cases:
        raw = orc.serialize(kind, f0, pl)
        assert len(raw) % 8 == 0
        assert orc.deserialize(raw) == (kind, f0, pl)


def test_worked_example_bytes():
    """SURVEY.md Appendix B worked example (derived from the Cap'n Proto spec — PARITY UNPINNED:
    the reference has no golden bytes): topics [0], message "test broadcast global"."""
    raw = orc.broadcast_frame([0], b"test broadcast global")
    expect = bytes.fromhex(
        "00000000" "09000000"                      # 1 segment, 9 words
        "00000000" "01000100"                      # root: struct, 1 data word, 1 pointer
        "0400000000000000"                         # union tag 4 = broadcast
        "00000000" "00000200"                      # → Broadcast: 0 data, 2 pointers
        "05000000" "0a000000"                      # topics: list, off 1, byte elems, count 1
        "05000000" "aa000000"                      # message: list, off 1, byte elems, count 21
        "0000000000000000"                         # topics [0] padded
    ) + b"test broadcast global" + b"\0\0\0"
    assert raw == expect
    assert len(raw) == 80
    # sizes used by BASELINE configs: L = 8*(6 + ceil(n/8) + ceil(K/8))
    assert len(orc.broadcast_frame([0], bytes(1024))) == 1080
    assert len(orc.broadcast_frame([0], bytes(4096))) == 4152
    assert len(orc.direct_frame(bytes(128), bytes(512))) == 688


def test_two_segment_layout():
    """payload that does not fit the 1024-word first segment goes behind a far pointer"""
    raw = orc.broadcast_frame([0], bytes(10000))
    assert int.from_bytes(raw[0:4], "little") == 1  # nseg-1
    s0 = int.from_bytes(raw[4:8], "little")
    s1 = int.from_bytes(raw[8:12], "little")
    assert s0 == 6 and s1 == 1 + 1250
    assert len(raw) == 16 + 8 * (s0 + s1)


def test_malformed_frames_rejected():
    good = orc.broadcast_frame([0], b"hello")
    assert orc.deserialize(good[:7]) is None
    assert orc.deserialize(good[:-8]) is None               # premature end of segment
    bad = bytearray(good); bad[0:4] = (600).to_bytes(4, "little")
    assert orc.deserialize(bytes(bad)) is None              # too many segments
    bad = bytearray(good); bad[16] = 9
    assert orc.deserialize(bytes(bad)) is None              # tag not in schema
    bad = bytearray(good); bad[36] = 0xFF; bad[37] = 0xFF   # topics list count beyond segment
    assert orc.deserialize(bytes(bad)) is None
    assert orc.deserialize(good + b"trailing") is not None   # reader ignores trailing bytes


def test_send_failure_removes_peer():
    """R13: tasks/user/sender.rs:23-29, tasks/broker/sender.rs:35-42"""
    o = orc.Oracle("/")
    a = o.add_user(b"a" * 8, [0])
    b = o.add_user(b"b" * 8, [0])
    o.close_conn(a)
    raw = orc.broadcast_frame([0], b"m1")
    o.handle_broadcast_message([0], raw)
    assert o.stream(a) == b"" and o.frames(b) == [raw]
    assert o.num_users() == 1 and o.route(b"a" * 8) == (0, -1)


@pytest.mark.parametrize("n_users,k", [(128, 1024), (2, 10000)])
def test_c1_reference_bench_shape(n_users, k):
    """BASELINE config C1 — the reference's own CPU-runnable case (cdn-broker/benches/broadcast.rs:50-75:
    users subscribed to one topic, user 0 sends a broadcast, every subscriber INCLUDING the sender
    receives the identical bytes) at 128 x 1 KiB, and the bench's literal 2 x 10 000 B shape"""
    o = orc.Oracle("/", 0)
    conns = [o.add_user(i.to_bytes(8, "little"), [0]) for i in range(n_users)]   # tests/mod.rs:111-115 keys
    raw = orc.broadcast_frame([0], bytes((i * 7 + 1) & 0xFF for i in range(k)))
    for _ in range(3):
        assert o.user_receive((0).to_bytes(8, "little"), raw) == 0               # user 0 is the sender
    L = len(raw)
    if k <= 8000:   # single segment; larger payloads leave the encoder in 2 segments
        assert L == 8 * (6 + 1 + (k + 7) // 8)
    for c in conns:
        assert o.frames(c) == [raw] * 3
        assert o.stream(c) == (L.to_bytes(4, "big") + raw) * 3                   # protocols/mod.rs:366-385
    assert o.bytes_sent() == 3 * n_users * L
