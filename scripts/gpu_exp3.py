"""B=1 per-iteration latency of the N=10 thread layouts (fixed iteration counts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
pb = synthetic_batch(8, N=10, seed=0)
args = [torch.from_numpy(a).cuda()[:1].contiguous() for a in pb.f32()]
for v in (0, 1, 2, 4):
    res = []
    for K in (0, 100, 1000):
        mpc = pkg.BatchedMPC(N=10, max_batch=1, warm_mode=0, kernel_variant=v, adaptive_rho_interval=0,
                             max_iter=K, check_every=5, eps_abs=0.0, eps_rel=0.0)
        out = mpc.alloc_outputs(1)
        for _ in range(5): mpc.solve(*args, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mpc.solve(*args, out=out); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(float(np.median(ts)))
    print(f"variant {v}: K=0 {res[0]:.1f} us, K=100 {res[1]:.1f} us, K=1000 {res[2]:.1f} us -> {(res[2]-res[1])/900:.3f} us/iter", flush=True)
