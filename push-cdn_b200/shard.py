"""Sharding of one broker's connections over the GPUs of a box.

The reference scales out by sharding *users over brokers* (the marshal hands a user to the
least-loaded broker, cdn-proto/src/connection/auth/marshal.rs:108-118) and forwarding a broadcast
once to every interested peer broker, which then fans out locally with `to_users_only = true`
(cdn-broker/src/tasks/broker/handler.rs:156-160,262-271).  Inside one box the same structure maps
to: every GPU owns a shard of the connections (tables, rings, direct map entries of its own users
only), the ingested batch is replicated to all GPUs with ONE collective (NCCL broadcast over NVLink;
gloo in the CPU tests) and each GPU runs the fan-out pipeline over its shard.  Nothing else crosses
GPUs: a connection lives on exactly one shard, so its FIFO order is the batch order (R9).

`torch.distributed` is plumbing here; the data path of every shard is the C-ABI engine.
"""
from __future__ import annotations

from typing import Iterable, List, Optional


def owner_of(key: bytes, world: int) -> int:
    """Deterministic owner shard of a user key: 64-bit FNV-1a over the key bytes, finalised with
    the murmur3 mixer (plain FNV keeps byte parity in its low bit), mod world.  Every rank computes
    the same owner without communication."""
    M = 0xFFFFFFFFFFFFFFFF
    h = 0xCBF29CE484222325
    for b in key:
        h = ((h ^ b) * 0x100000001B3) & M
    h ^= h >> 33
    h = (h * 0xFF51AFD7ED558CCD) & M
    h ^= h >> 33
    h = (h * 0xC4CEB9FE1A85EC53) & M
    h ^= h >> 33
    return h % world


class ShardedBroker:
    """One rank's view of a broker whose connections are sharded over `world` ranks.

    State calls are issued identically on every rank (they are tiny control-plane messages); only
    the owner applies them to its engine.  Peer brokers (cross-host mesh) are pinned to rank 0.
    """

    def __init__(self, engine, rank: int, world: int, group=None):
        self.e, self.rank, self.world, self.group = engine, rank, world, group

    # ---- state (Connections::*), applied on the owner only -----------------------------------
    def owns(self, key: bytes) -> bool:
        return owner_of(key, self.world) == self.rank

    def add_user(self, key: bytes, topics: Iterable[int] = ()) -> Optional[int]:
        return self.e.add_user(key, topics) if self.owns(key) else None

    def remove_user(self, key: bytes) -> None:
        if self.owns(key):
            self.e.remove_user(key)

    def subscribe_user_to(self, key: bytes, topics: Iterable[int]) -> None:
        if self.owns(key):
            self.e.subscribe_user_to(key, topics)

    def unsubscribe_user_from(self, key: bytes, topics: Iterable[int]) -> None:
        if self.owns(key):
            self.e.unsubscribe_user_from(key, topics)

    def add_broker(self, ident: str) -> Optional[int]:
        return self.e.add_broker(ident) if self.rank == 0 else None

    def subscribe_broker_to(self, ident: str, topics: Iterable[int]) -> None:
        if self.rank == 0:
            self.e.subscribe_broker_to(ident, topics)

    def apply_user_sync(self, remote_identity: str, entries) -> None:
        """remote users: the direct-map entry lives on rank 0 (where the peer broker connection
        is); an entry that moves one of OUR users away must reach that user's owner too."""
        ents = list(entries)
        mine = [e for e in ents if self.rank == 0 or self.owns(e[0])]
        if mine:
            self.e.apply_user_sync(remote_identity, mine)

    # ---- data: the one exchange step -----------------------------------------------------------
    def ingest(self, tensor, src: int = 0):
        """replicate the ingested batch (frames or descriptors) from `src` to every shard"""
        if self.world > 1:
            import torch.distributed as dist

            dist.broadcast(tensor, src=src, group=self.group)
        return tensor

    def route_local(self, msgs: List[tuple]) -> int:
        """route one replicated batch over the local shard (see Engine.submit); every rank passes
        the same ordered list.  A direct message whose recipient is remote (owned by a peer
        broker) is forwarded by rank 0 only; local recipients resolve on their owner, all other
        ranks drop it as 'unknown' — exactly one delivery box-wide."""
        return self.e.submit(msgs)
