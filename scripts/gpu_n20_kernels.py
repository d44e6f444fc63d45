"""GPU experiment (round 2): dense against Riccati kernel at N = 12 / 16 / 20 with two batch sizes and rho policies."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
from gpu_riccati_exp import run
for N in (12, 16, 20):
    ric = 0 if N >= 20 else 5
    for B in (4096, 16384):
        pb = synthetic_batch(B, N=N, seed=0)
        for name, v, opts in (("dense tol3", 5 - ric, dict(adaptive_rho_tolerance=3.0)), ("riccati tol3", ric, dict(adaptive_rho_tolerance=3.0)),
                              ("riccati tol1.5", ric, dict(adaptive_rho_tolerance=1.5))):
            if not pkg._capi.has_variant(N, v):
                continue
            r = run(pb, v, **opts)
            print(json.dumps(dict(N=N, B=B, kernel=name, ms=r["ms"], solves_s=B / r["ms"] * 1e3, iters=float(r["it"].mean()), max_it=int(r["it"].max()))), flush=True)
