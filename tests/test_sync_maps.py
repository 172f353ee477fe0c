"""Inter-broker sync on the product's tables, CPU only (host-only engines): the
reference's topic-sync unit tests replayed on two engines, and a randomized two-broker mesh where
the product's user/topic sync exchange must leave both brokers resolving routes exactly like two
oracle brokers exchanging the reference's maps."""
import random

import pytest

from oracle import oracle as orc


def _pair(pcdn):
    local = pcdn.Engine(device=-1, max_conns=64, identity="test/local")
    remote = pcdn.Engine(device=-1, max_conns=64, identity="test/remote")
    rb = local.add_broker("test/remote")
    remote.add_broker("test/local")
    return local, remote, rb


def _sync_topics(frm, to, frm_ident, full=False, apply=True):
    ents = frm.get_topic_sync(full)
    if not ents:
        return False
    if apply:
        to.apply_topic_sync(frm_ident, ents)
    return True


def test_topic_sync(pcdn):
    """cdn-broker/src/connections/mod.rs:410-470"""
    local, remote, rb = _pair(pcdn)
    remote.subscribe_user_to(bytes([1]), [0, 1])
    assert remote.get_topic_sync(full=True) == []                   # full sync is None
    assert _sync_topics(remote, local, "test/remote")
    assert local.debug_interested([0]) == [rb] and local.debug_interested([1]) == [rb]
    remote.unsubscribe_user_from(bytes([1]), [0])
    assert _sync_topics(remote, local, "test/remote")
    assert local.debug_interested([0]) == [] and local.debug_interested([1]) == [rb]


def test_topic_sync_out_of_order(pcdn):
    """cdn-broker/src/connections/mod.rs:474-526"""
    local, remote, rb = _pair(pcdn)
    remote.subscribe_user_to(bytes([1]), [0, 1])
    assert _sync_topics(remote, local, "test/remote", apply=False)   # computed but never delivered
    remote.unsubscribe_user_from(bytes([1]), [0])
    remote.unsubscribe_user_from(bytes([1]), [1])
    assert _sync_topics(remote, local, "test/remote")
    remote.subscribe_user_to(bytes([1]), [1])
    assert _sync_topics(remote, local, "test/remote")
    assert _sync_topics(remote, local, "test/remote", full=True)
    assert local.debug_interested([0]) == [] and local.debug_interested([1]) == [rb]


@pytest.mark.parametrize("seed", range(4))
def test_two_broker_mesh_matches_oracle(pcdn, seed):
    rng = random.Random(seed)
    names = ["a/a", "b/b"]
    E = [pcdn.Engine(device=-1, max_conns=256, max_keys=1024, identity=n) for n in names]
    O = [orc.Oracle(n) for n in names]
    maps = [{}, {}]
    for i in (0, 1):
        maps[i][E[i].add_broker(names[1 - i])] = O[i].add_broker(names[1 - i])
    keys = [bytes([k]) * rng.choice([1, 8, 32]) for k in range(40)]

    def check():
        for i in (0, 1):
            for k in keys:
                kind, conn = E[i].debug_route(k)
                okind, oconn = O[i].route(k)
                assert kind == okind and (maps[i][conn] if conn >= 0 else -1) == oconn, (i, k)
            for t in range(6):
                for flag in (False, True):
                    assert sorted(maps[i][c] for c in E[i].debug_interested([t], flag)) == O[i].interested([t], flag)
            assert E[i].num_users()[0] == O[i].num_users()

    for step in range(300):
        i = rng.randrange(2)
        k = rng.choice(keys)
        t = [rng.randrange(6) for _ in range(rng.randrange(0, 3))]
        op = rng.randrange(8)
        if op < 2:
            maps[i][E[i].add_user(k, t)] = O[i].add_user(k, t)     # a user (re)connects to broker i
        elif op == 2:
            E[i].remove_user(k); O[i].remove_user(k)
        elif op == 3:
            E[i].subscribe_user_to(k, t); O[i].subscribe_user_to(k, t)
        elif op == 4:
            E[i].unsubscribe_user_from(k, t); O[i].unsubscribe_user_from(k, t)
        elif op == 5:   # partial user sync i → other (sync.rs:70-89)
            ents = E[i].get_user_sync(False)
            sent = O[i].user_sync_to(O[1 - i], full=False)
            assert bool(ents) == sent
            if ents:
                E[1 - i].apply_user_sync(names[i], ents)
        elif op == 6:   # partial topic sync i → other (sync.rs:113-128)
            ents = E[i].get_topic_sync(False)
            sent = O[i].topic_sync_to(O[1 - i], names[i], full=False)
            assert bool(ents) == sent
            if ents:
                E[1 - i].apply_topic_sync(names[i], ents)
        else:           # full syncs (on connect, handler.rs:97-117)
            ents = E[i].get_user_sync(True)
            sent = O[i].user_sync_to(O[1 - i], full=True)
            assert bool(ents) == sent
            if ents:
                E[1 - i].apply_user_sync(names[i], ents)
            ents = E[i].get_topic_sync(True)
            sent = O[i].topic_sync_to(O[1 - i], names[i], full=True)
            assert bool(ents) == sent
            if ents:
                E[1 - i].apply_topic_sync(names[i], ents)
        if step % 15 == 0:
            check()
    check()
