import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
pb = synthetic_batch(4096, N=30, seed=0)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
for v in (0, 2):
    for name, extra in (("default", dict()), ("K=0", dict(adaptive_rho_interval=0, max_iter=0, check_every=100000)), ("K=25", dict(adaptive_rho_interval=0, max_iter=25, check_every=100000))):
        mpc = pkg.BatchedMPC(N=30, max_batch=4096, warm_mode=0, kernel_variant=v, **extra)
        out = mpc.alloc_outputs(4096)
        for _ in range(2): mpc.solve(*args, out=out)
        torch.cuda.synchronize(); ts=[]
        for _ in range(3):
            flush.fill_(1); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mpc.solve(*args, out=out); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        it = out[2].cpu().numpy()
        print(f"N=30 variant {v} {name}: {np.median(ts):.3f} ms {4096/np.median(ts):.1f} k solves/s iters {it.mean():.1f} max {it.max()}", flush=True)
