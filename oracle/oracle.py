"""ORACLE — test infrastructure only.

ctypes front-end of ``oracle/broker_oracle.cpp`` (the C++ CPU restatement of the cdn-broker routing
hot path).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
module; the product package must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
TIMED_PATH = os.path.join(_HERE, "cpu_broker_timed")

KIND_DIRECT, KIND_BROADCAST, KIND_SUBSCRIBE, KIND_UNSUBSCRIBE, KIND_USER_SYNC, KIND_TOPIC_SYNC = 3, 4, 5, 6, 7, 8


def build(force: bool = False) -> None:
    """Compile the oracle (g++, a few seconds).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, "broker_oracle.cpp"), os.path.join(_HERE, "capnp_lite.hpp")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB_PATH, srcs[0]], cwd=_HERE
        )
    tsrc = os.path.join(_HERE, "cpu_broker_timed.cpp")
    if os.path.exists(tsrc) and (
        force or not os.path.exists(TIMED_PATH) or os.path.getmtime(tsrc) > os.path.getmtime(TIMED_PATH)
    ):
        subprocess.check_call(
            ["g++", "-O3", "-march=native", "-std=c++17", "-pthread", "-o", TIMED_PATH, tsrc], cwd=_HERE
        )


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u8p, u16p, u32, u64, ci, cp = C.c_void_p, C.c_char_p, C.POINTER(C.c_uint16), C.c_uint32, C.c_uint64, C.c_int, C.c_char_p
        sig = {
            "orc_create": (vp, [cp, u32]),
            "orc_destroy": (None, [vp]),
            "orc_add_user": (ci, [vp, u8p, u32, u16p, u32]),
            "orc_remove_user": (None, [vp, u8p, u32]),
            "orc_subscribe_user_to": (None, [vp, u8p, u32, u16p, u32]),
            "orc_unsubscribe_user_from": (None, [vp, u8p, u32, u16p, u32]),
            "orc_add_broker": (ci, [vp, cp]),
            "orc_remove_broker": (None, [vp, cp]),
            "orc_subscribe_broker_to": (None, [vp, cp, u16p, u32]),
            "orc_unsubscribe_broker_from": (None, [vp, cp, u16p, u32]),
            "orc_dmap_new": (vp, [cp]),
            "orc_dmap_free": (None, [vp]),
            "orc_dmap_put": (None, [vp, u8p, u32, u64, cp]),
            "orc_apply_user_sync": (None, [vp, vp]),
            "orc_topic_sync": (ci, [vp, vp, cp, ci, ci]),
            "orc_user_sync": (ci, [vp, vp, ci, ci]),
            "orc_apply_topic_list": (None, [vp, cp, u16p, u32]),
            "orc_handle_broadcast_message": (None, [vp, u16p, u32, u8p, u32, ci]),
            "orc_handle_direct_message": (None, [vp, u8p, u32, u8p, u32, ci]),
            "orc_user_receive": (ci, [vp, u8p, u32, u8p, u32]),
            "orc_broker_receive": (ci, [vp, u8p, u32]),
            "orc_num_conns": (u32, [vp]),
            "orc_stream_len": (u64, [vp, ci]),
            "orc_stream_ptr": (C.c_void_p, [vp, ci]),
            "orc_stream_frames": (u32, [vp, ci]),
            "orc_stream_clear": (None, [vp, ci]),
            "orc_stream_clear_all": (None, [vp]),
            "orc_close_conn": (None, [vp, ci]),
            "orc_conn_removed": (ci, [vp, ci]),
            "orc_user_conn": (ci, [vp, u8p, u32]),
            "orc_broker_conn": (ci, [vp, cp]),
            "orc_num_users": (u32, [vp]),
            "orc_bytes_sent": (u64, [vp]),
            "orc_deliveries": (u64, [vp]),
            "orc_interested": (u32, [vp, u16p, u32, ci, C.POINTER(C.c_int), u32]),
            "orc_interested_counts": (None, [vp, u16p, u32, ci, C.POINTER(u32), C.POINTER(u32)]),
            "orc_route": (ci, [vp, u8p, u32, C.POINTER(C.c_int)]),
            "orc_serialize": (u64, [ci, u8p, u32, u8p, u32, C.c_void_p, u64]),
            "orc_deserialize": (ci, [u8p, u64, C.c_void_p, u32, C.POINTER(u32), C.c_void_p, u32, C.POINTER(u32)]),
            "orc_rel_new": (vp, []),
            "orc_rel_free": (None, [vp]),
            "orc_rel_assoc": (None, [vp, cp, C.POINTER(u64), u32]),
            "orc_rel_dissoc": (None, [vp, cp, C.POINTER(u64), u32]),
            "orc_rel_remove_key": (None, [vp, cp]),
            "orc_rel_keys_by_value": (u32, [vp, u64, C.c_char_p, u32]),
            "orc_rel_values": (u32, [vp, C.POINTER(u64), u32]),
            "orc_rel_values_of_key": (u32, [vp, cp, C.POINTER(u64), u32]),
            "orc_rel_num_keys": (u32, [vp]),
            "orc_rel_num_values": (u32, [vp]),
            "orc_ver_new": (vp, [u64]),
            "orc_ver_free": (None, [vp]),
            "orc_ver_insert": (None, [vp, cp, cp]),
            "orc_ver_remove": (None, [vp, cp]),
            "orc_ver_get": (ci, [vp, cp, C.c_char_p, u32]),
            "orc_ver_get_full": (vp, [vp]),
            "orc_ver_diff": (vp, [vp]),
            "orc_ver_merge": (u32, [vp, vp]),
            "orc_ver_remove_by_value_no_modify": (None, [vp, cp]),
            "orc_ver_len": (u32, [vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _t16(topics: Iterable[int]):
    t = list(topics)
    arr = (C.c_uint16 * max(1, len(t)))(*t)
    return arr, len(t)


# ------------------------------------------------------------------------------------- capnp-lite
def serialize(kind: int, field0: bytes = b"", payload: bytes = b"") -> bytes:
    """Message::serialize (cdn-proto/src/message.rs:116-204) for the routed kinds."""
    L = lib()
    n = L.orc_serialize(kind, field0, len(field0), payload, len(payload), None, 0)
    buf = C.create_string_buffer(int(n))
    L.orc_serialize(kind, field0, len(field0), payload, len(payload), C.cast(buf, C.c_void_p), n)
    return buf.raw


def broadcast_frame(topics: Sequence[int], message: bytes) -> bytes:
    return serialize(KIND_BROADCAST, bytes(topics), message)


def direct_frame(recipient: bytes, message: bytes) -> bytes:
    return serialize(KIND_DIRECT, recipient, message)


def deserialize(raw: bytes):
    """Returns (kind, field0, payload) or None on Error::Deserialize."""
    L = lib()
    f0 = C.create_string_buffer(max(1, len(raw)))
    pl = C.create_string_buffer(max(1, len(raw)))
    n0, n1 = C.c_uint32(0), C.c_uint32(0)
    k = L.orc_deserialize(raw, len(raw), C.cast(f0, C.c_void_p), len(raw), C.byref(n0), C.cast(pl, C.c_void_p), len(raw), C.byref(n1))
    if k < 0:
        return None
    return k, f0.raw[: n0.value], pl.raw[: n1.value]


# ------------------------------------------------------------------------------------- broker
class Oracle:
    """The reference broker's `Inner` + `Connections`, one message at a time."""

    def __init__(self, identity: str = "self/self", n_valid_topics: int = 0):
        self.L = lib()
        self.h = self.L.orc_create(identity.encode(), n_valid_topics)
        self.identity = identity

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # state -----------------------------------------------------------------------------------
    def add_user(self, key: bytes, topics: Iterable[int] = ()) -> int:
        t, n = _t16(topics)
        return self.L.orc_add_user(self.h, key, len(key), t, n)

    def remove_user(self, key: bytes) -> None:
        self.L.orc_remove_user(self.h, key, len(key))

    def subscribe_user_to(self, key: bytes, topics: Iterable[int]) -> None:
        t, n = _t16(topics)
        self.L.orc_subscribe_user_to(self.h, key, len(key), t, n)

    def unsubscribe_user_from(self, key: bytes, topics: Iterable[int]) -> None:
        t, n = _t16(topics)
        self.L.orc_unsubscribe_user_from(self.h, key, len(key), t, n)

    def add_broker(self, ident: str) -> int:
        return self.L.orc_add_broker(self.h, ident.encode())

    def remove_broker(self, ident: str) -> None:
        self.L.orc_remove_broker(self.h, ident.encode())

    def subscribe_broker_to(self, ident: str, topics: Iterable[int]) -> None:
        t, n = _t16(topics)
        self.L.orc_subscribe_broker_to(self.h, ident.encode(), t, n)

    def unsubscribe_broker_from(self, ident: str, topics: Iterable[int]) -> None:
        t, n = _t16(topics)
        self.L.orc_unsubscribe_broker_from(self.h, ident.encode(), t, n)

    def apply_user_sync(self, remote_identity: str, entries) -> None:
        """entries: iterable of (key, version, owner_or_None) — a remote DirectMap."""
        m = self.L.orc_dmap_new(remote_identity.encode())
        for key, version, owner in entries:
            self.L.orc_dmap_put(m, key, len(key), version, None if owner is None else owner.encode())
        self.L.orc_apply_user_sync(self.h, m)
        self.L.orc_dmap_free(m)

    def apply_topic_list(self, ident: str, topics: Iterable[int]) -> None:
        t, n = _t16(topics)
        self.L.orc_apply_topic_list(self.h, ident.encode(), t, n)

    def user_sync_to(self, other: "Oracle", full: bool = False, apply: bool = True) -> bool:
        return bool(self.L.orc_user_sync(self.h, other.h, 1 if full else 0, 1 if apply else 0))

    def topic_sync_to(self, other: "Oracle", my_id_in_other: str, full: bool = False, apply: bool = True) -> bool:
        return bool(self.L.orc_topic_sync(self.h, other.h, my_id_in_other.encode(), 1 if full else 0, 1 if apply else 0))

    # data ------------------------------------------------------------------------------------
    def handle_broadcast_message(self, topics: Iterable[int], raw: bytes, to_users_only: bool = False) -> None:
        t, n = _t16(topics)
        self.L.orc_handle_broadcast_message(self.h, t, n, raw, len(raw), int(to_users_only))

    def handle_direct_message(self, recipient: bytes, raw: bytes, to_user_only: bool = False) -> None:
        self.L.orc_handle_direct_message(self.h, recipient, len(recipient), raw, len(raw), int(to_user_only))

    def user_receive(self, sender: bytes, raw: bytes) -> int:
        return self.L.orc_user_receive(self.h, sender, len(sender), raw, len(raw))

    def broker_receive(self, raw: bytes) -> int:
        return self.L.orc_broker_receive(self.h, raw, len(raw))

    # observation -----------------------------------------------------------------------------
    def num_conns(self) -> int:
        return self.L.orc_num_conns(self.h)

    def stream(self, conn: int) -> bytes:
        n = self.L.orc_stream_len(self.h, conn)
        if n == 0:
            return b""
        return C.string_at(self.L.orc_stream_ptr(self.h, conn), n)

    def stream_len(self, conn: int) -> int:
        return self.L.orc_stream_len(self.h, conn)

    def frames(self, conn: int):
        """Split a connection's byte stream back into raw frames (read_length_delimited)."""
        s, out, p = self.stream(conn), [], 0
        while p < len(s):
            ln = int.from_bytes(s[p : p + 4], "big")
            out.append(s[p + 4 : p + 4 + ln])
            p += 4 + ln
        return out

    def clear(self, conn: Optional[int] = None) -> None:
        if conn is None:
            self.L.orc_stream_clear_all(self.h)
        else:
            self.L.orc_stream_clear(self.h, conn)

    def close_conn(self, conn: int) -> None:
        self.L.orc_close_conn(self.h, conn)

    def conn_removed(self, conn: int) -> bool:
        return bool(self.L.orc_conn_removed(self.h, conn))

    def user_conn(self, key: bytes) -> int:
        return self.L.orc_user_conn(self.h, key, len(key))

    def broker_conn(self, ident: str) -> int:
        return self.L.orc_broker_conn(self.h, ident.encode())

    def num_users(self) -> int:
        return self.L.orc_num_users(self.h)

    def bytes_sent(self) -> int:
        return self.L.orc_bytes_sent(self.h)

    def deliveries(self) -> int:
        return self.L.orc_deliveries(self.h)

    def interested(self, topics: Iterable[int], to_users_only: bool = False):
        t, n = _t16(topics)
        cap = self.num_conns() + 1
        out = (C.c_int * cap)()
        k = self.L.orc_interested(self.h, t, n, int(to_users_only), out, cap)
        return sorted(out[i] for i in range(k))

    def interested_counts(self, topics: Iterable[int], to_users_only: bool = False):
        t, n = _t16(topics)
        nb, nu = C.c_uint32(0), C.c_uint32(0)
        self.L.orc_interested_counts(self.h, t, n, int(to_users_only), C.byref(nb), C.byref(nu))
        return nb.value, nu.value

    def route(self, key: bytes):
        c = C.c_int(-1)
        k = self.L.orc_route(self.h, key, len(key), C.byref(c))
        return k, c.value
