import sys, os, json
sys.path.insert(0, '/root/repo/scripts'); sys.path.insert(0, '/root/repo')
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
from gpu_riccati_exp import run
pb = synthetic_batch(16384, N=30, seed=0)
for v in (0, 4, 5):
    r = run(pb, v)
    print(json.dumps(dict(variant=v, ms=r["ms"], solves_s=pb.B / r["ms"] * 1e3, iters=float(r["it"].mean()), solved=float((r["st"] == 1).mean()))), flush=True)
