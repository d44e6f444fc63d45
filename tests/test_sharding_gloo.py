"""Multi-process host logic of the N>1 path on CPU: world_size-2 gloo.  The solve itself
needs a GPU, so the per-problem results are stood in for by a deterministic function of
the problem data; what is tested is the partition (contiguous, balanced, ragged, empty)
and the gather back into global batch order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mpc_b200 as pkg
from mpc_b200.sharding import shard_range, shard_sizes, gather_results


def test_shard_range_partitions():
    for B in (0, 1, 7, 8, 4096, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(B, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sz = shard_sizes(B, world)
            assert max(sz) - min(sz) <= 1 and sum(sz) == B
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _fake_solve(pb):
    """stand-in for the GPU solve: depends on every input of the problem."""
    u0 = (pb.x0[:, :12] * 3.0 + pb.r[:, 0].reshape(pb.B, 12) + pb.mu[:, None]).astype(np.float32)
    iters = (pb.stance.sum((1, 2)) + pb.tick % 7).astype(np.int32)
    status = (pb.gait_id % 2).astype(np.int32)
    return u0, iters, status


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb = pkg.problems.synthetic_batch(B, N=5, gaits=pkg.problems.GAIT_NAMES, seed=123)
    lo, hi = shard_range(B, rank, world)
    u0, iters, status = _fake_solve(pb.slice(lo, hi))
    g = gather_results(torch.from_numpy(u0), torch.from_numpy(iters), torch.from_numpy(status), B)
    if rank == 0:
        q.put([t.numpy() for t in g])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [37, 2])
def test_gather_matches_single_process(B):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pb = pkg.problems.synthetic_batch(B, N=5, gaits=pkg.problems.GAIT_NAMES, seed=123)
    u0, iters, status = _fake_solve(pb)
    assert np.array_equal(got[0], u0) and np.array_equal(got[1], iters) and np.array_equal(got[2], status)
