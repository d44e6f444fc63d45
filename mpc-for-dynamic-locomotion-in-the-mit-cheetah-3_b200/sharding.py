"""Batch-index sharding across the GPUs of one box.

MPC problems are independent (one robot / state / gait each; the only state that survives
a tick is the per-problem warm start, which lives on the GPU that owns the problem), so a
batch shards by contiguous batch-index ranges with NO collective on the solve path.  A
collective (NCCL on the GPU box, gloo in the CPU tests) is used only to gather the
first-stage forces and solver statistics when a caller wants them in one place.
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int):
    """Contiguous, balanced partition of range(B): first (B % world) ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(B: int, world: int):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_results(u0, iters, status, B: int, group=None):
    """all_gather of per-rank results into global-batch order.

    u0 [b_r,12] float32, iters [b_r] int32, status [b_r] int32 torch tensors of this rank's
    shard (shard_range(B, rank, world)).  Returns (u0 [B,12], iters [B], status [B]) on every
    rank.  64 bytes per problem - irrelevant to throughput; shards may be ragged."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    bmax = max(sizes)
    dev = u0.device
    pack = torch.zeros((bmax, 14), dtype=torch.float32, device=dev)
    n = u0.shape[0]
    pack[:n, :12] = u0
    pack[:n, 12] = iters.to(torch.float32)
    pack[:n, 13] = status.to(torch.float32)
    bufs = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(bufs, pack, group=group)
    full = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    return full[:, :12].contiguous(), full[:, 12].to(torch.int32), full[:, 13].to(torch.int32)


def solve_sharded(mpc, pb, rank: int, world: int, device=None):
    """Solve this rank's shard of a ProblemBatch on its GPU.  Returns (lo, hi, U, X, stats)."""
    import torch
    lo, hi = shard_range(pb.B, rank, world)
    sub = pb.slice(lo, hi)
    dev = torch.device("cuda", mpc.device) if device is None else device
    args = [torch.from_numpy(a).to(dev) for a in sub.f32()]
    U, X, st = mpc.solve(*args)
    return lo, hi, U, X, st
