"""Phase clocks of CTA 0 (CMPC_DEBUG_CLOCKS=1): B=1 alone on the GPU, and inside a full batch."""
import sys, os
os.environ["CMPC_DEBUG_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
V = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pb = synthetic_batch(4096 if N == 10 else 1024, N=N, seed=0)
full = [torch.from_numpy(a).cuda() for a in pb.f32()]
for B in (1, pb.B):
    args = [t[:B].contiguous() for t in full]
    mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, lpt_schedule=0, kernel_variant=V)
    out = mpc.alloc_outputs(B)
    for i in range(2):
        print(f"B={B} run {i}: iters[0]={int(out[2][0])}", file=sys.stderr, flush=True)
        mpc.solve(*args, out=out); torch.cuda.synchronize()
