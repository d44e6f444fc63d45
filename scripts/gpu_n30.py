import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
def timeit(mpc, args, out, n=3):
    for _ in range(2): mpc.solve(*args, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): mpc.solve(*args, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for N, B, variants in ((30, 4096, (0, 2)), (20, 4096, (0, 2)), (60, 1024, (0,)), (40, 1024, (0,))):
    pb = synthetic_batch(B, N=N, seed=0)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    for v in variants:
        for extra in (dict(), dict(adaptive_rho_interval=0, max_iter=0, check_every=100000, eps_abs=0., eps_rel=0.)):
            mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, kernel_variant=v, **extra)
            out = mpc.alloc_outputs(B)
            ms = timeit(mpc, args, out)
            it = out[2].cpu().numpy(); st = out[5].cpu().numpy()
            print(f"N={N} variant {v} {'K=0' if extra else 'default'}: {ms:.2f} ms/batch {B/ms*1e3/1e3:.1f} k solves/s iters mean {it.mean():.1f} max {it.max()} solved {np.mean(st==1):.4f}", flush=True)
    one = [t[:1].contiguous() for t in args]
    mpc = pkg.BatchedMPC(N=N, max_batch=1, warm_mode=0, kernel_variant=variants[-1])
    out = mpc.alloc_outputs(1)
    print(f"N={N} B=1 latency (variant {variants[-1]}): {timeit(mpc, one, out, n=20)*1e3:.1f} us, iters {int(out[2][0])}", flush=True)
