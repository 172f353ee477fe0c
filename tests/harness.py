"""Restatement of the reference's in-process broker test harness
(cdn-broker/src/tests/mod.rs:117-412: TestUser / TestBroker / TestDefinition / TestRun and the
`send_message_as!` / `assert_received!` / `at_index!` macros) over two interchangeable backends:

* ``OracleBackend`` — oracle/ (CPU restatement of the reference broker), one message at a time;
* ``EngineBackend`` — the product, through the C ABI (include/pcdn_fanout.h) on a CUDA device.

The reference injects users with ``Connections::add_user`` and brokers with ``add_broker`` followed by
a TopicSync and a UserSync frame (tests/mod.rs:258-389); both backends do exactly that.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

from oracle import oracle as orc

GLOBAL, DA = 0, 1  # TestTopic, cdn-proto/src/def.rs:22-28
IDENTITY = "/"     # broker under test: both advertise endpoints are String::new() (tests/mod.rs:229-232)


def at_index(i: int) -> bytes:
    """at_index! — (index as usize).to_le_bytes() (tests/mod.rs:108-115)"""
    return int(i).to_bytes(8, "little")


@dataclass
class TestUser:
    __test__ = False
    index: int
    subscribed_topics: Sequence[int] = ()

    @property
    def public_key(self) -> bytes:
        return at_index(self.index)


@dataclass
class TestBroker:
    __test__ = False
    connected_users: List[TestUser] = field(default_factory=list)


class OracleBackend:
    """oracle/: processes every frame immediately."""

    def __init__(self, n_valid_topics: int = 2):
        self.o = orc.Oracle(IDENTITY, n_valid_topics)
        self._taken = {}

    def add_user(self, key, topics):
        return self.o.add_user(key, topics)

    def remove_user(self, key):
        self.o.remove_user(key)

    def add_broker(self, ident):
        return self.o.add_broker(ident)

    def remove_broker(self, ident):
        self.o.remove_broker(ident)

    def subscribe_broker_to(self, ident, topics):
        self.o.subscribe_broker_to(ident, topics)

    def unsubscribe_broker_from(self, ident, topics):
        self.o.unsubscribe_broker_from(ident, topics)

    def subscribe_user_to(self, key, topics):
        self.o.subscribe_user_to(key, topics)

    def unsubscribe_user_from(self, key, topics):
        self.o.unsubscribe_user_from(key, topics)

    def apply_user_sync(self, remote_identity, entries):
        self.o.apply_user_sync(remote_identity, entries)

    def user_receive(self, key, raw):
        return self.o.user_receive(key, raw)

    def broker_receive(self, ident, raw):
        return self.o.broker_receive(raw)

    def handle_broadcast_message(self, topics, raw, to_users_only=False):
        self.o.handle_broadcast_message(topics, raw, to_users_only)

    def handle_direct_message(self, recipient, raw, to_user_only=False):
        self.o.handle_direct_message(recipient, raw, to_user_only)

    def flush(self):
        pass

    def take_frames(self, conn):
        """frames delivered to `conn` since the previous take (FIFO)"""
        fr = self.o.frames(conn)
        k = self._taken.get(conn, 0)
        self._taken[conn] = len(fr)
        return fr[k:]

    def stream(self, conn):
        return self.o.stream(conn)

    def interested(self, topics, to_users_only=False):
        return self.o.interested(topics, to_users_only)

    def route(self, key):
        return self.o.route(key)

    def user_conn(self, key):
        return self.o.user_conn(key)

    def num_users(self):
        return self.o.num_users()


class EngineBackend:
    """The product through the C ABI.  Frames are batched; ``flush`` launches the batch on the
    device, polls it, reads every span back and releases it."""

    def __init__(self, pcdn, n_valid_topics: int = 2, **cfg):
        kw = dict(max_conns=256, max_topics=256, max_keys=1024, ring_bytes_per_conn=1 << 16,
                  max_batch_msgs=1024, max_batch_bcast=256, max_batch_bytes=1 << 22,
                  max_batch_deliveries=1 << 16, identity=IDENTITY, n_valid_topics=n_valid_topics)
        kw.update(cfg)
        self.e = pcdn.Engine(**kw)
        self._frames = {}

    def add_user(self, key, topics):
        self.flush()
        return self.e.add_user(key, topics)

    def remove_user(self, key):
        self.flush()
        self.e.remove_user(key)

    def add_broker(self, ident):
        self.flush()
        return self.e.add_broker(ident)

    def remove_broker(self, ident):
        self.flush()
        self.e.remove_broker(ident)

    def subscribe_broker_to(self, ident, topics):
        self.flush()
        self.e.subscribe_broker_to(ident, topics)

    def unsubscribe_broker_from(self, ident, topics):
        self.flush()
        self.e.unsubscribe_broker_from(ident, topics)

    def subscribe_user_to(self, key, topics):
        self.flush()
        self.e.subscribe_user_to(key, topics)

    def unsubscribe_user_from(self, key, topics):
        self.flush()
        self.e.unsubscribe_user_from(key, topics)

    def apply_user_sync(self, remote_identity, entries):
        self.flush()
        self.e.apply_user_sync(remote_identity, entries)

    def user_receive(self, key, raw):
        return self.e.user_receive(key, raw)

    def broker_receive(self, ident, raw):
        return self.e.broker_receive(ident, raw)

    def handle_broadcast_message(self, topics, raw, to_users_only=False):
        self.e.handle_broadcast_message(topics, raw, to_users_only)

    def handle_direct_message(self, recipient, raw, to_user_only=False):
        self.e.handle_direct_message(recipient, raw, to_user_only)

    def flush(self):
        for conn, frames in self.e.drain().items():
            self._frames.setdefault(conn, []).extend(frames)

    def take_frames(self, conn):
        self.flush()
        return self._frames.pop(conn, [])

    def interested(self, topics, to_users_only=False):
        return self.e.debug_interested(topics, to_users_only)

    def route(self, key):
        return self.e.debug_route(key)

    def num_users(self):
        return self.e.num_users()[0]


class TestRun:
    """TestRun (tests/mod.rs:176-183): the actors' ends of the connections."""
    __test__ = False

    def __init__(self, backend, users: Sequence[TestUser], brokers: Sequence[TestBroker]):
        self.b = backend
        self.user_keys = [u.public_key for u in users]
        self.connected_users = []    # conn handles, same order as the definition
        self.connected_brokers = []
        self.broker_ids = []
        # inject_users (tests/mod.rs:258-300)
        for u in users:
            self.connected_users.append(backend.add_user(u.public_key, list(u.subscribed_topics)))
        # inject_brokers (tests/mod.rs:308-389)
        for i, br in enumerate(brokers):
            ident = f"{i}/{i}"
            self.broker_ids.append(ident)
            self.connected_brokers.append(backend.add_broker(ident))
            topics = [t for u in br.connected_users for t in u.subscribed_topics]
            # TopicSync of a fresh TopicSyncMap: every listed topic becomes Subscribed
            # (apply_topic_sync → subscribe_broker_to, connections/mod.rs:165-191)
            seen = []
            for t in topics:
                if t not in seen:
                    seen.append(t)
            if seen:
                backend.subscribe_broker_to(ident, seen)
            # UserSync: DirectMap::new(identifier) with user → identifier, version 1 (its diff)
            if br.connected_users:
                backend.apply_user_sync(ident, [(u.public_key, 1, ident) for u in br.connected_users])

    # send_message_as! (tests/mod.rs:48-56): the actor's frame enters the broker's receive loop
    def send_as_user(self, idx: int, raw: bytes) -> int:
        return self.b.user_receive(self.user_keys[idx], raw)

    def send_as_broker(self, idx: int, raw: bytes) -> int:
        return self.b.broker_receive(self.broker_ids[idx], raw)

    # assert_received!(yes, actor, message) (tests/mod.rs:88-105)
    def assert_received(self, conn: int, raw: bytes) -> None:
        got = self._pending(conn)
        assert got, "timed out trying to receive message"
        first = got.pop(0)
        assert first == raw, "was supposed to receive a message but did not"

    # assert_received!(no, all, ...) (tests/mod.rs:62-86)
    def assert_nothing_more(self) -> None:
        for c in list(self.connected_users) + list(self.connected_brokers):
            assert not self._pending(c), "wasn't supposed to receive a message but did"

    def _pending(self, conn):
        if not hasattr(self, "_q"):
            self._q = {}
        self._q.setdefault(conn, []).extend(self.b.take_frames(conn))
        return self._q[conn]


@dataclass
class TestDefinition:
    __test__ = False
    connected_users: List[TestUser] = field(default_factory=list)
    connected_brokers: List[TestBroker] = field(default_factory=list)

    def into_run(self, backend) -> TestRun:
        return TestRun(backend, self.connected_users, self.connected_brokers)


def Broadcast(topics, message: bytes) -> bytes:
    """Message::Broadcast{topics, message}.serialize()"""
    return orc.broadcast_frame(list(topics), message)


def Direct(recipient: bytes, message: bytes) -> bytes:
    return orc.direct_frame(recipient, message)
