"""BASELINE.json configs C3 / C4 / C5 at sizes the oracle finishes in seconds — bit-exact against the
oracle through the C ABI (the full-size runs of the same shapes are in bench_configs.py / bench.py)."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_parity import World

pytestmark = pytest.mark.gpu


def test_c4_direct_many_keys(pcdn):
    """config 4 shape: 128-byte keys, uniform recipients, 512-byte payloads, 10 % unknown keys"""
    rng = random.Random(4)
    w = World(pcdn, max_conns=32768, max_keys=32768, max_batch_msgs=65536, max_batch_bytes=64 << 20,
              ring_bytes_per_conn=1 << 15, max_batch_deliveries=1 << 17)
    keys = [rng.getrandbits(1024).to_bytes(128, "little") for _ in range(30000)]
    for k in keys:
        w.add_user(k, [])
    payload = bytes(range(256)) * 2
    for j in range(40000):
        rc = rng.choice(keys) if rng.random() < 0.9 else rng.getrandbits(1024).to_bytes(128, "little")
        w.direct(rc, orc.direct_frame(rc, payload))
    n = w.check()
    assert 35000 < n < 37000
    assert w.e.last_result.n_direct_dropped == 40000 - n


def test_c3_zipf_mixed_sizes_extended_topics(pcdn):
    """config 3 shape: Zipf-0.99 subscriptions over MORE than 256 topics (engine topic ids are u16;
    the wire's List(UInt8) cannot carry them, the ABI takes topics separately), payloads 256 B-64 KiB"""
    rng = np.random.default_rng(3)
    T, n, M = 600, 3000, 48
    p = 1.0 / np.arange(1, T + 1) ** 0.99
    p /= p.sum()
    w = World(pcdn, max_conns=4096, max_topics=1024, ring_bytes_per_conn=4 << 20, max_batch_bytes=16 << 20,
              max_batch_deliveries=1 << 18)
    for i in range(n):
        w.add_user(i.to_bytes(8, "little") * 4, [int(t) for t in rng.choice(T, size=8, replace=False, p=p)])
    for j in range(M):
        t = int(rng.choice(T, p=p))
        k = int(rng.choice([256 << i for i in range(9)]))
        raw = orc.broadcast_frame([t & 0xFF], bytes(rng.integers(0, 256, size=k, dtype=np.uint8)))
        w.bcast([t], raw)
    assert w.check() > 5000


@pytest.mark.parametrize("dense", [True, False])
def test_c5_4k_broadcast(pcdn, dense):
    """config 5 shard shape: 4 KiB payloads (records > 4 KiB: message-major staged path), all
    subscribed / 4 of 64 topics"""
    rng = random.Random(5)
    w = World(pcdn, max_conns=8192, ring_bytes_per_conn=1 << 17, max_batch_deliveries=1 << 18)
    for i in range(6000):
        w.add_user(i.to_bytes(8, "little"), [0] if dense else rng.sample(range(64), 4))
    for m in range(8 if dense else 40):
        t = 0 if dense else rng.randrange(64)
        w.bcast([t], orc.broadcast_frame([t], bytes([m]) * 4096))
    n = w.check()
    assert n == 48000 if dense else n > 10000


@pytest.mark.parametrize("n_users,k", [(128, 1024), (2, 10000)])
def test_c1_reference_bench_shape(pcdn, n_users, k):
    """config 1 — the reference's own CPU bench shape (cdn-broker/benches/broadcast.rs:50-75) through the
    GPU engine: every subscriber including the sender gets the identical bytes, one message per batch"""
    w = World(pcdn, max_conns=256, ring_bytes_per_conn=1 << 16)
    for i in range(n_users):
        w.add_user(i.to_bytes(8, "little"), [0])
    raw = orc.broadcast_frame([0], bytes((i * 7 + 1) & 0xFF for i in range(k)))
    for _ in range(3):
        assert w.e.user_receive((0).to_bytes(8, "little"), raw) == 0
        assert w.o.user_receive((0).to_bytes(8, "little"), raw) == 0
        assert w.check() == n_users
