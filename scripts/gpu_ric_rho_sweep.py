"""GPU experiment (round 2): rho policy for the Riccati kernel.  Its refactorisation costs ~6 iterations (dense
kernel at N=30: ~80), so adapting rho earlier / more often may pay.  Config 4 subset (4096 x N=30) and N=60."""
import json, sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
from gpu_riccati_exp import run
for N, B, gaits in ((30, 4096, ("trot",)), (60, 512, ("pseudo_gallop",))):
    pb = synthetic_batch(B, N=N, gaits=gaits, seed=0)
    base = run(pb, 0)
    print(json.dumps(dict(N=N, cfg="default", ms=base["ms"], iters=float(base["it"].mean()), max_it=int(base["it"].max()), solved=float((base["st"] == 1).mean()))), flush=True)
    for rho_s, interval, tol in itertools.product((0.5, 1.0, 2.0), (10, 15, 25), (1.5, 2.0, 3.0)):
        r = run(pb, 0, reps=3, rho=0.05 * N * rho_s, adaptive_rho_interval=interval, adaptive_rho_tolerance=tol, rho_min=0.1 * 0.05 * N * rho_s)
        print(json.dumps(dict(N=N, rho0=0.05 * N * rho_s, interval=interval, tol=tol, ms=r["ms"], iters=float(r["it"].mean()), max_it=int(r["it"].max()),
                              solved=float((r["st"] == 1).mean()))), flush=True)
