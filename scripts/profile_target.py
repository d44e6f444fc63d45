"""Short program for ncu: 2 warm-up + 3 timed launches of the solve kernel on config 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pb = synthetic_batch(B, N=N, seed=0)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0)
out = mpc.alloc_outputs(B)
for _ in range(5):
    mpc.solve(*args, out=out)
torch.cuda.synchronize()
print("iters mean", float(out[2].float().mean()), "solved", float((out[5] == 1).float().mean()))
