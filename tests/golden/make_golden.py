#!/usr/bin/env python
"""Generates tests/golden/*.json.

The reference (Rust) cannot be run in this environment, and its own tests hold no golden byte
vectors for this path: what is pinned against the REFERENCE are the routing
scenarios in tests/scenarios.py.  The fixtures written here pin the ORACLE and the product against
silent regressions: wire frames for the sizes BASELINE.json uses (derived from the Cap'n Proto
spec — "parity unpinned", see oracle/capnp_lite.hpp) and the delivered byte streams of one seeded
mixed workload (sha256 per connection).  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def frames():
    out = []
    cases = [
        (orc.KIND_BROADCAST, bytes([0]), b"test broadcast global"),            # broadcast.rs:58-61
        (orc.KIND_BROADCAST, bytes([1]), b"test broadcast DA"),
        (orc.KIND_DIRECT, (0).to_bytes(8, "little"), b"test direct 0"),        # direct.rs:51-54
        (orc.KIND_SUBSCRIBE, bytes([0, 1]), b""),
        (orc.KIND_UNSUBSCRIBE, bytes([1]), b""),
        (orc.KIND_BROADCAST, bytes([0]), bytes(1024)),                          # C2: L = 1080
        (orc.KIND_BROADCAST, bytes([0]), bytes(4096)),                          # C5: L = 4152
        (orc.KIND_DIRECT, bytes(128), bytes(512)),                              # C4: L = 688
        (orc.KIND_BROADCAST, bytes([0]), bytes(10000)),                         # benches/broadcast.rs:26 (2 segments)
        (orc.KIND_BROADCAST, b"", b""),
    ]
    for kind, f0, pl in cases:
        raw = orc.serialize(kind, f0, pl)
        out.append({"kind": kind, "field0": f0.hex(), "payload_len": len(pl), "payload_sha256": hashlib.sha256(pl).hexdigest(),
                    "len": len(raw), "sha256": hashlib.sha256(raw).hexdigest(), "hex": raw.hex() if len(raw) <= 256 else None})
    return out


def workload(seed=20240921):
    """a seeded mixed workload; returns (setup ops, frames list, per-connection stream digests)"""
    rng = random.Random(seed)
    o = orc.Oracle("/", 8)
    users = []
    ops = []
    for i in range(300):
        key = rng.getrandbits(64).to_bytes(8, "little") * rng.choice([1, 4, 16])
        topics = [t for t in range(8) if rng.random() < (0.5 if t == 0 else 0.1)]
        users.append(key)
        ops.append(["add_user", key.hex(), topics])
        o.add_user(key, topics)
    ops.append(["add_broker", "9/9", [0, 3]])
    o.add_broker("9/9"); o.subscribe_broker_to("9/9", [0, 3])
    ops.append(["user_sync", "9/9", [[b"elsewhere".hex(), 1, "9/9"]]])
    o.apply_user_sync("9/9", [(b"elsewhere", 1, "9/9")])
    fr = []
    for j in range(120):
        r = rng.random()
        size = 17000 if j in (17, 83) else rng.choice([0, 7, 64, 600, 2500])
        pl = bytes([(j * 7 + k) & 0xFF for k in range(size)])
        if r < 0.6:
            topics = [rng.randrange(10) for _ in range(rng.randrange(1, 4))]
            raw = orc.broadcast_frame(topics, pl)
        else:
            rc = rng.choice(users) if rng.random() < 0.8 else rng.choice([b"elsewhere", b"nobody"])
            raw = orc.direct_frame(rc, pl)
        origin = 1 if rng.random() < 0.2 else 0
        sender = rng.choice(users)
        rc = o.broker_receive(raw) if origin else o.user_receive(sender, raw)
        fr.append({"sender": sender.hex(), "origin": origin, "raw": raw.hex(), "rc": rc})
    digests = {}
    for c in range(o.num_conns()):
        s = o.stream(c)
        if s:
            digests[str(c)] = {"len": len(s), "sha256": hashlib.sha256(s).hexdigest()}
    return {"seed": seed, "n_valid_topics": 8, "ops": ops, "frames": fr, "streams": digests,
            "deliveries": o.deliveries(), "bytes_sent": o.bytes_sent()}


if __name__ == "__main__":
    json.dump(frames(), open(os.path.join(HERE, "frames.json"), "w"), indent=1)
    json.dump(workload(), open(os.path.join(HERE, "workload.json"), "w"))
    print("wrote frames.json, workload.json")
