"""GPU experiment (round 2): how much of config 2's batch time is the straggler tail, and what a
separate latency-oriented launch of the hardest problems could gain.  Needs a build with
-DCMPC_EXTRA_LAYOUTS for the wide layouts (kernel_variant 1 = <10,2,4>, 2 = <10,4,2>)."""
import json, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, ProblemBatch

def sub(pb, idx):
    return ProblemBatch(pb.x0[idx], pb.r[idx], pb.stance[idx], pb.x_des[idx], pb.mu[idx], pb.gait_id[idx], pb.tick[idx])

def timeit(pb, reps=20, **opts):
    dev = torch.device("cuda", 0)
    args = [torch.from_numpy(a).to(dev) for a in pb.f32()]
    mpc = pkg.BatchedMPC(N=pb.N, max_batch=pb.B, warm_mode=0, time_kernel=1, **opts)
    out = mpc.alloc_outputs(pb.B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for i in range(reps + 3):
        flush.fill_(1)
        mpc.solve(*args, out=out)
        if i >= 3:
            ms.append(mpc.last_kernel_ms)
    it = out[2].cpu().numpy()
    mpc.close()
    return float(np.median(ms)), it

if __name__ == "__main__":
    pb = synthetic_batch(4096, N=10, seed=0)
    t_all, iters = timeit(pb)
    order = np.argsort(-iters)
    print(json.dumps(dict(what="full batch", ms=t_all, iters_top=iters[order[:12]].tolist(), mean=float(iters.mean()),
                          n_gt50=int((iters > 50).sum()), n_gt100=int((iters > 100).sum()))), flush=True)
    np.save("gpurun_out/config2_iters.npy", iters)
    for K in (8, 32, 64, 148, 296):
        rest = np.sort(order[K:])
        t_rest, _ = timeit(sub(pb, rest))
        row = dict(K=K, min_iters_in_top=int(iters[order[K - 1]]), ms_rest=t_rest)
        hard = sub(pb, np.sort(order[:K]))
        for v in (0, 1, 2):
            if pkg._capi.has_variant(10, v):
                row[f"ms_hard_v{v}"], _ = timeit(hard, kernel_variant=v, lpt_schedule=0)
        print(json.dumps(row), flush=True)
    # iteration caps: batch time if nobody ran longer than cap (lower bound of a two-phase scheme's phase 1)
    for cap in (40, 60, 100):
        t_cap, it = timeit(pb, max_iter=cap)
        print(json.dumps(dict(cap=cap, ms=t_cap, unfinished=int((it >= cap).sum()))), flush=True)
