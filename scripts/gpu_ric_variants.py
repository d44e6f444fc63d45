"""Riccati kernel at N=30 (config 4) with different launch bounds (resident CTAs per SM), against the dense kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
from gpu_riccati_exp import run
pb = synthetic_batch(16384, N=30, seed=0)
for v, name in ((5, "dense <30,6,1,3>"), (0, "riccati, 4 CTAs/SM (default)"), (4, "riccati, 5 CTAs/SM"), (3, "riccati, 6 CTAs/SM")):
    if not pkg._capi.has_variant(30, v):
        continue
    r = run(pb, v)
    print(json.dumps(dict(variant=v, kernel=name, ms=r["ms"], solves_s=pb.B / r["ms"] * 1e3, iters=float(r["it"].mean()), solved=float((r["st"] == 1).mean()))), flush=True)
