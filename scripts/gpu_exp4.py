"""Register-blocked (R rows per thread) layouts vs the default: agreement and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def dev_time(mpc, args, out, n=10):
    for _ in range(3): mpc.solve(*args, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.solve(*args, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
cases = [(10, 4096, (0, 3, 4), 10)]
if len(sys.argv) > 1: cases.append((30, 4096, (0, 1, 3, 4), 3))
for N, B, variants, n in cases:
    pb = synthetic_batch(B, N=N, seed=0)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    ref = None
    for v in variants:
        for name, extra in (("default", dict()), ("K=0", dict(adaptive_rho_interval=0, max_iter=0, check_every=100000)),
                            ("K=25", dict(adaptive_rho_interval=0, max_iter=25, check_every=100000))):
            try:
                mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, kernel_variant=v, **extra)
                out = mpc.alloc_outputs(B)
                ms = dev_time(mpc, args, out, n)
            except Exception as e:
                print(f"N={N} variant {v} {name}: FAILED {e}", flush=True); continue
            it = out[2].cpu().numpy(); st = out[5].cpu().numpy(); U = out[0].cpu().numpy()
            msg = ""
            if name == "default":
                if ref is None: ref = (U.copy(), it.copy())
                else: msg = f" | vs variant 0: max|dU| {np.abs(U-ref[0]).max():.2e}, iters equal {np.mean(it==ref[1]):.4f}"
            print(f"N={N} variant {v} {name}: {ms:.4f} ms {B/ms/1e3:.3f} M solves/s iters mean {it.mean():.1f} max {it.max()} solved {np.mean(st==1):.4f}{msg}", flush=True)
