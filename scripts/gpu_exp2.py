"""GPU experiments (round 1, second session): thread-layout variants, host-path modes.
usage: python scripts/gpu_exp2.py [variants|host|n30 ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch

what = set(sys.argv[1:]) or {"variants", "host", "n30"}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def dev_time(mpc, args, out, n=10):
    for _ in range(3):
        mpc.solve(*args, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.solve(*args, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


if "variants" in what:
    pb = synthetic_batch(4096, N=10, seed=0)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    for v in (0, 1, 2, 3, 4):
        for lpt in (1024, 0):
            mpc = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, kernel_variant=v, lpt_schedule=lpt)
            out = mpc.alloc_outputs(4096)
            med, mn = dev_time(mpc, args, out)
            it = out[2].cpu().numpy()
            print(f"N=10 variant {v} lpt {lpt}: median {med:.4f} ms min {mn:.4f} ms  {4096/med/1e3:.2f} M solves/s iters mean {it.mean():.1f} max {it.max()}", flush=True)
    # K = 0 / fixed K probes of the default layout
    for name, extra in (("K=0", dict(adaptive_rho_interval=0, max_iter=0, check_every=100000)),
                        ("K=25 fixed", dict(adaptive_rho_interval=0, max_iter=25, check_every=100000))):
        mpc = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, **extra)
        out = mpc.alloc_outputs(4096)
        med, mn = dev_time(mpc, args, out)
        print(f"N=10 variant 0 {name}: median {med:.4f} ms min {mn:.4f}", flush=True)

if "host" in what:
    pb = synthetic_batch(4096, N=10, seed=0)
    hin = [torch.from_numpy(a).pin_memory().numpy() for a in pb.f32()]
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    hout = (pin((4096, 10, 12), torch.float32), None, pin((4096,), torch.int32),
            pin((4096,), torch.float32), pin((4096,), torch.float32), pin((4096,), torch.int32))
    for zc in (1, 0):
        for lpt in (1024, 0):
            mpc = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, host_zero_copy=zc, lpt_schedule=lpt)
            for _ in range(5):
                mpc.solve_host(*hin, want_X=False, out=hout)
            torch.cuda.synchronize()
            ts = []
            for _ in range(30):
                t0 = time.perf_counter()
                mpc.solve_host(*hin, want_X=False, out=hout)
                ts.append(time.perf_counter() - t0)
            med = float(np.median(ts)) * 1e3
            print(f"host path zero_copy {zc} lpt {lpt}: median {med:.4f} ms min {min(ts)*1e3:.4f} ms  {4096/med/1e3:.2f} M solves/s e2e", flush=True)

if "n30" in what:
    for N, B, variants in ((30, 4096, (0, 1, 3)),):
        pb = synthetic_batch(B, N=N, seed=0)
        args = [torch.from_numpy(a).cuda() for a in pb.f32()]
        for v in variants:
            for name, extra in (("default", dict()), ("K=0", dict(adaptive_rho_interval=0, max_iter=0, check_every=100000))):
                mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, kernel_variant=v, **extra)
                out = mpc.alloc_outputs(B)
                med, mn = dev_time(mpc, args, out, n=3)
                it = out[2].cpu().numpy(); st = out[5].cpu().numpy()
                print(f"N={N} variant {v} {name}: {med:.2f} ms/batch {B/med:.1f} k solves/s iters mean {it.mean():.1f} max {it.max()} solved {np.mean(st==1):.4f}", flush=True)
