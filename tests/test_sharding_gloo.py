"""N>1 host path on CPU: two processes, gloo backend, no GPU.  Each process creates the SAME sharded
broker through the C ABI as a host-only mirror (`world_shards=2, first_shard=rank, device=-1`) and
replays the same control-plane calls (the SPMD contract of a multi-process group,
include/pcdn_fanout.h).  Checks: (a) every process hands out the same connection ids and balances
the shards, (b) the ingest collective replicates a batch bit-exactly (gloo stands in for NCCL),
(c) the union over shards of what each shard would deliver — the slice of the recipient set in its
id range, the direct route if the target lives there — equals ONE unsharded oracle broker, each
recipient on exactly one shard."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    from oracle import oracle as orc

    pcdn = ge.load_package()
    eng = pcdn.Engine(device=-1, max_conns=400, max_keys=8192, identity="/", world_shards=world, first_shard=rank)
    d = eng.shard_info(0)
    stride = d.shard_stride
    assert d.global_index == rank and d.conn_base == rank * stride and eng.num_shards() == (0, world)
    lo, hi = rank * stride, (rank + 1) * stride
    o = orc.Oracle("/")  # the unsharded reference broker, replayed identically on every rank
    rng = random.Random(5)  # same seed everywhere: identical control-plane stream
    keys = [rng.getrandbits(64).to_bytes(8, "little") * rng.choice([1, 4]) for _ in range(600)]
    m = {}
    for k in keys:
        t = [x for x in range(6) if rng.random() < 0.3]
        m[eng.add_user(k, t)] = o.add_user(k, t)
    m[eng.add_broker("peer/peer")] = o.add_broker("peer/peer")
    eng.subscribe_broker_to("peer/peer", [1, 4]); o.subscribe_broker_to("peer/peer", [1, 4])
    for k in rng.sample(keys, 80):
        eng.remove_user(k); o.remove_user(k)
    for k in rng.sample(keys, 80):
        t = [rng.randrange(6)]
        eng.subscribe_user_to(k, t); o.subscribe_user_to(k, t)
    for k in rng.sample(keys, 40):   # reconnects: kicked, new id on the least-loaded shard
        t = [rng.randrange(6)]
        m[eng.add_user(k, t)] = o.add_user(k, t)
    ents = [(b"remote-%d" % i, 1, "peer/peer") for i in range(5)] + [(keys[3], 9, "peer/peer")]
    eng.apply_user_sync("peer/peer", ents); o.apply_user_sync("peer/peer", ents)

    # (a) identical ids in every process, shards balanced, ids inside the usable part of a shard
    ids = torch.tensor(sorted(m), dtype=torch.int64)
    ref = ids.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ids, ref)
    per_shard = [int(((ids // stride) == s).sum()) for s in range(world)]
    assert all(c % stride < 400 for c in m) and max(per_shard) - min(per_shard) <= 2, per_shard

    # (b) ingest replication is bit-exact
    arena = torch.zeros(4096, dtype=torch.uint8)
    if rank == 0:
        arena = torch.randint(0, 256, (4096,), dtype=torch.uint8)
    want = arena.clone()
    dist.broadcast(want, src=0)
    dist.broadcast(arena, src=0)
    assert torch.equal(arena, want)

    # (c) what THIS shard would deliver: the recipients in its id range; a direct message if its
    #     target connection lives here (the direct map is replicated, the target's shard packs)
    local = []
    for topics in ([0], [1], [2, 3], [4, 5, 0], [1, 4]):
        for flag in (False, True):
            local.append(sorted(m[c] for c in eng.debug_interested(topics, flag) if lo <= c < hi))
    direct = []
    for k in keys[:200] + [e[0] for e in ents] + [b"nobody"]:
        kind, conn = eng.debug_route(k)
        mine = conn >= 0 and lo <= conn < hi
        direct.append((kind, m[conn]) if mine else (0, -1))
    torch.save({"local": local, "direct": direct}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    if rank == 0:
        parts = [torch.load(os.path.join(out_dir, f"rank{r}.pt")) for r in range(world)]
        i = 0
        for topics in ([0], [1], [2, 3], [4, 5, 0], [1, 4]):
            for flag in (False, True):
                union = sorted(x for p in parts for x in p["local"][i])
                assert union == o.interested(topics, flag), (topics, flag)  # disjoint + complete
                i += 1
        for j, k in enumerate(keys[:200] + [e[0] for e in ents] + [b"nobody"]):
            hits = [p["direct"][j] for p in parts if p["direct"][j][0] != 0]
            okind, oconn = o.route(k)
            if okind == 0 or oconn < 0:
                assert hits == [], k
            else:
                assert hits == [(okind, oconn)], (k, hits, okind, oconn)   # exactly one shard delivers
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded_oracle(tmp_path):
    import __graft_entry__ as ge

    ge.build()
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def test_sharded_mirror_in_one_process(pcdn):
    """the same check without processes: a 3-shard host-only mirror against the oracle"""
    from oracle import oracle as orc

    eng = pcdn.Engine(device=-1, max_conns=100, max_keys=2048, identity="/", world_shards=3, first_shard=0)
    o = orc.Oracle("/")
    rng = random.Random(2)
    m = {}
    keys = [b"k%d" % i for i in range(240)]
    for k in keys:
        t = [rng.randrange(4)]
        m[eng.add_user(k, t)] = o.add_user(k, t)
    stride = eng.shard_info(0).shard_stride
    assert sorted({c // stride for c in m}) == [0, 1, 2]
    for topics in ([0], [1, 2], [3]):
        assert sorted(m[c] for c in eng.debug_interested(topics)) == o.interested(topics, False)
    # the id space is full at 3 x 100: the next user is refused, a removed id comes back
    for i in range(60):
        eng.add_user(b"x%d" % i, [])
    with pytest.raises(pcdn.PcdnError) as ei:
        eng.add_user(b"one-too-many", [])
    assert ei.value.code == -5
    eng.remove_user(b"x7")
    eng.add_user(b"one-too-many", [])
