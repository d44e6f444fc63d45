"""GPU experiment (round 2): rho adaptation policy of the dense N=10 kernel on config 2 (4096 trot) and a
config-3 shard (8192 mixed gaits): kernel ms, mean / max iterations per (interval, tolerance)."""
import json, sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
from gpu_riccati_exp import run
for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)), ("config3_shard", synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0)))):
    base = run(pb, 0)
    print(json.dumps(dict(cfg=name, policy="default", ms=base["ms"], iters=float(base["it"].mean()), max_it=int(base["it"].max()), solved=float((base["st"] == 1).mean()))), flush=True)
    for interval, tol in itertools.product((10, 15, 20, 25, 35), (1.5, 2.0, 3.0, 5.0)):
        r = run(pb, 0, reps=3, adaptive_rho_interval=interval, adaptive_rho_tolerance=tol)
        same = (r["st"] == 1) & (base["st"] == 1)
        print(json.dumps(dict(cfg=name, interval=interval, tol=tol, ms=r["ms"], iters=float(r["it"].mean()), max_it=int(r["it"].max()),
                              solved=float((r["st"] == 1).mean()), max_dX=float(np.abs(r["X"][same] - base["X"][same]).max()))), flush=True)
