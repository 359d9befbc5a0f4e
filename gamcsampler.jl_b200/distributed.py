"""One-process-per-GPU plumbing for row-sharded jobs.  torch.distributed is used only to
ship the 128-byte NCCL unique id and to agree on shard boundaries; the data-path collective (one
all-reduce of [loglik, gradient, metric] per target evaluation) is issued by the C library on its own
communicator and stream."""
from __future__ import annotations

import numpy as np


def shard_rows(N: int, world: int, rank: int):
    """Contiguous row block [r0, r1) of rank `rank` (sizes differ by at most one row)."""
    base, rem = divmod(int(N), int(world))
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def broadcast_unique_id(engine_cls, rank: int, world: int):
    """Rank 0 creates the NCCL unique id (gamc_comm_unique_id); everybody receives it."""
    import torch.distributed as dist
    payload = [engine_cls.comm_unique_id().tobytes() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    return np.frombuffer(payload[0], dtype=np.uint8).copy()


def init_comm(engine, rank: int, world: int):
    uid = broadcast_unique_id(type(engine), rank, world)
    engine.comm_init(uid, world, rank)
    return uid


def allreduce_partials_host(parts):
    """Host-side stand-in for the NCCL all-reduce used by CPU-only tests: sums per-shard (ll, g, G) tuples."""
    ll = sum(p[0] for p in parts)
    g = sum(p[1] for p in parts) if parts[0][1] is not None else None
    G = sum(p[2] for p in parts) if parts[0][2] is not None else None
    return ll, g, G


def shard_chains(nchains: int, world: int, rank: int):
    """Block of independent chains [c0, c1) owned by `rank` (chains shard trivially, no collective)."""
    return shard_rows(nchains, world, rank)
