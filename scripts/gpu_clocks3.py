"""Per-phase cycle totals of the Riccati kernel (CMPC_DEBUG_CLOCKS=1), CTA 0, alone and in a full wave."""
import os, sys
os.environ["CMPC_DEBUG_CLOCKS"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for B in (1, 148 * 4):
    pb = synthetic_batch(B, N=N, seed=0)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, lpt_schedule=0, max_iter=50, eps_abs=0.0, eps_rel=0.0, check_every=5,
                         adaptive_rho_interval=0, kernel_variant=5)
    for _ in range(2):
        print(f"B={B} (50 iterations, 1 factorisation):", file=sys.stderr, flush=True)
        mpc.solve(*args)
        torch.cuda.synchronize()
