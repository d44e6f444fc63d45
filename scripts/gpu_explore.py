"""Exploration on the GPU box: accuracy vs refresh period, iteration statistics, timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, DT
from oracle import condensed_admm as ca

def timeit(mpc, args, out, n=10):
    for _ in range(3): mpc.solve(*args, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): mpc.solve(*args, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

pb = synthetic_batch(48, N=10, seed=5)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
K = 40
refs = None
for refresh in (1, 2, 5, 10, 25, 0):
    mpc = pkg.BatchedMPC(N=10, max_batch=48, max_iter=K, check_every=100000, eps_abs=0.0, eps_rel=0.0, warm_mode=0,
                         adaptive_rho_interval=0, refresh_every=refresh)
    U, X, st = mpc.solve(*args); torch.cuda.synchronize()
    U = U.cpu().numpy().astype(np.float64)
    if refs is None:
        refs = [ca.solve_problem(*pb.problem(b)[:4], pb.problem(b)[4], DT, fixed_iters=K, rho=0.3) for b in range(48)]
    errs = np.array([np.abs(U[b] - refs[b]['U']).max() for b in range(48)])
    viol = np.array([(np.abs(U[b] - refs[b]['U']) - 1e-3 * np.abs(refs[b]['U'])).max() for b in range(48)])
    print(f"refresh {refresh:3d}: max err {errs.max():.2e} median {np.median(errs):.2e}  max(err - 1e-3|U|) {viol.max():.2e}", flush=True)

pb = synthetic_batch(4096, N=10, seed=0)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
for name, opts in [("default (lpt)", dict()), ("no lpt", dict(lpt_schedule=0)), ("lpt rho .5", dict(rho=0.5)), ("lpt tol3", dict(adaptive_rho_tolerance=3.0)),
                   ("K=25 fixed", dict(adaptive_rho_interval=0, max_iter=25, check_every=100000, eps_abs=0., eps_rel=0.)),
                   ("K=0", dict(adaptive_rho_interval=0, max_iter=0, check_every=100000, eps_abs=0., eps_rel=0.))]:
    mpc = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, **opts)
    out = mpc.alloc_outputs(4096)
    ms = timeit(mpc, args, out)
    it = out[2].cpu().numpy(); stt = out[5].cpu().numpy()
    print(f"{name:20s}: {ms:.3f} ms/batch  {4096/ms*1e3/1e6:.2f} M solves/s  iters mean {it.mean():.1f} p99 {np.percentile(it,99):.0f} max {it.max()}  solved {np.mean(stt==1):.4f} nan {np.mean(stt==-1):.4f}", flush=True)
# lone-CTA latency: B=1, fixed K
one = [t[:1].contiguous() for t in args]
for v in (0,):
    res = []
    for K in (0, 100, 1000):
        mpc = pkg.BatchedMPC(N=10, max_batch=1, warm_mode=0, kernel_variant=v, adaptive_rho_interval=0, max_iter=K, check_every=5, eps_abs=0., eps_rel=0.)
        out = mpc.alloc_outputs(1)
        res.append(timeit(mpc, one, out, n=20))
    print(f"variant {v}: B=1 K=0 {res[0]*1e3:.1f} us, K=100 {res[1]*1e3:.1f} us, K=1000 {res[2]*1e3:.1f} us -> {(res[2]-res[1])/900*1e3:.3f} us/iter", flush=True)
