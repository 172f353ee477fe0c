"""N>1 host path on CPU: two processes, gloo backend.  Checks that (a) every rank derives the same
owner for every key, (b) the ingest collective replicates the batch bit-exactly, (c) the union of
the shards' routing tables (host mirrors of the engines) resolves every broadcast / direct exactly
like ONE unsharded oracle broker — each recipient on exactly one shard."""
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
    import importlib.util

    spec = importlib.util.spec_from_file_location("pcdn_shard", os.path.join(ROOT, "push-cdn_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)

    eng = pcdn.Engine(device=-1, max_conns=4096, max_keys=8192, identity="/")
    sb = shard.ShardedBroker(eng, rank, world)
    o = orc.Oracle("/")  # the unsharded reference broker, replayed identically on every rank
    rng = random.Random(5)  # same seed everywhere: identical control-plane stream
    keys = [rng.getrandbits(64).to_bytes(8, "little") * rng.choice([1, 4]) for _ in range(600)]
    mine = {}
    for k in keys:
        t = [x for x in range(6) if rng.random() < 0.3]
        c = sb.add_user(k, t)
        oc = o.add_user(k, t)
        if c is not None:
            mine[c] = oc
    bc = sb.add_broker("peer/peer")
    obc = o.add_broker("peer/peer")
    if bc is not None:
        mine[bc] = obc
    sb.subscribe_broker_to("peer/peer", [1, 4])
    o.subscribe_broker_to("peer/peer", [1, 4])
    for k in rng.sample(keys, 80):
        sb.remove_user(k); o.remove_user(k)
    for k in rng.sample(keys, 80):
        t = [rng.randrange(6)]
        sb.subscribe_user_to(k, t); o.subscribe_user_to(k, t)
    ents = [(b"remote-%d" % i, 1, "peer/peer") for i in range(5)] + [(keys[3], 9, "peer/peer")]
    sb.apply_user_sync("peer/peer", ents); o.apply_user_sync("peer/peer", ents)

    # (a) owners agree across ranks
    owners = torch.tensor([shard.owner_of(k, world) for k in keys], dtype=torch.int64)
    ref = owners.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(owners, ref)
    assert 0.3 < float((owners == 0).float().mean()) < 0.7

    # (b) ingest replication is bit-exact
    arena = torch.zeros(4096, dtype=torch.uint8)
    if rank == 0:
        arena = torch.randint(0, 256, (4096,), dtype=torch.uint8)
    want = arena.clone()
    dist.broadcast(want, src=0)
    sb.ingest(arena, src=0)
    assert torch.equal(arena, want)

    # (c) local recipients of every broadcast/direct, mapped to oracle connection ids
    local = []
    for topics in ([0], [1], [2, 3], [4, 5, 0], [1, 4]):
        for flag in (False, True):
            local.append(sorted(mine[c] for c in eng.debug_interested(topics, flag)))
    direct = []
    for k in keys[:200] + [e[0] for e in ents] + [b"nobody"]:
        kind, conn = eng.debug_route(k)
        # remote routes resolve on rank 0 only (where the peer broker connection lives)
        if kind == 2 and rank != 0:
            kind, conn = 0, -1
        direct.append((kind, mine.get(conn, -1) if conn >= 0 else -1))
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
            if okind == 0:
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
