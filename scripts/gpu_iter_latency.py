"""GPU experiment (round 2): latency of one ADMM iteration of solve_kernel<10,...> as a function of SM
load (problems per SM) and of the refresh / check periods, from fixed-iteration launches (eps = 0):
us per iteration = (T(K2) - T(K1)) / (K2 - K1)."""
import json, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch

def t_fixed(pb, K, reps=10, **opts):
    dev = torch.device("cuda", 0)
    args = [torch.from_numpy(a).to(dev) for a in pb.f32()]
    o = dict(max_iter=K, eps_abs=0.0, eps_rel=0.0, warm_mode=0, time_kernel=1, lpt_schedule=0, adaptive_rho_interval=0,
             check_every=1000000, refresh_every=5)
    o.update(opts)
    mpc = pkg.BatchedMPC(N=pb.N, max_batch=pb.B, **o)
    out = mpc.alloc_outputs(pb.B)
    ms = []
    for i in range(reps + 2):
        mpc.solve(*args, out=out)
        if i >= 2:
            ms.append(mpc.last_kernel_ms)
    mpc.close()
    return float(np.median(ms))

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
V = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for B in (1, 148, 148 * 4, 148 * 8, 148 * 16):
    pb = synthetic_batch(B, N=N, seed=1)
    row = dict(N=N, B=B, per_sm=B / 148, variant=V)
    for name, opts in (("default", {}), ("no_refresh", dict(refresh_every=0))):
        opts = dict(opts, kernel_variant=V)
        t1, t2 = t_fixed(pb, 100, **opts), t_fixed(pb, 300, **opts)
        row[name] = dict(cycles_per_iter=round(1.965e6 * (t2 - t1) / 200, 1), ms_K0=round(t1 - 100 * (t2 - t1) / 200, 4))
    print(json.dumps(row), flush=True)
