import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
def run(N, B, gaits, mu, seeds, optlist):
    for opts in optlist:
        ts=[]; its=[]; mx=[]; sol=[]
        for seed in seeds:
            pb = synthetic_batch(B, N=N, gaits=gaits, seed=seed, mu=mu)
            args = [torch.from_numpy(a).cuda() for a in pb.f32()]
            mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, **opts)
            out = mpc.alloc_outputs(B)
            mpc.solve(*args, out=out); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mpc.solve(*args, out=out); mpc.solve(*args, out=out); e1.record(); torch.cuda.synchronize()
            it = out[2].cpu().numpy(); st = out[5].cpu().numpy()
            ts.append(e0.elapsed_time(e1)/2); its.append(it.mean()); mx.append(it.max()); sol.append(np.mean(st==1))
        print(f"N={N} {gaits[0] if len(gaits)==1 else 'mixed'} mu{mu} {opts}: {np.mean(ts):.3f} ms iters mean {np.mean(its):.1f} max {max(mx)} solved {min(sol):.4f}", flush=True)
grid10 = [dict(rho=r, rho_max=300.) for r in (0.3, 0.5, 0.7, 1.0)] + [dict(rho=0.5, rho_max=300., adaptive_rho_tolerance=3.0), dict(rho=0.5, rho_max=300., adaptive_rho_interval=15), dict(rho=0.5, rho_max=30.)]
run(10, 4096, ("trot",), (1.0, 1.0), (0, 1, 2), grid10)
run(10, 4096, GAIT_NAMES, (0.3, 1.0), (3,), grid10[:4])
run(30, 2048, ("trot",), (1.0, 1.0), (0,), [dict(rho=r, rho_max=300.) for r in (1.0, 2.0, 3.0, 5.0)])
run(60, 256, ("trot",), (1.0, 1.0), (0,), [dict(rho=r, rho_max=300.) for r in (2.0, 4.0, 8.0)])
run(20, 2048, ("trot",), (1.0, 1.0), (0,), [dict(rho=r, rho_max=300.) for r in (0.5, 1.0, 2.0)])
