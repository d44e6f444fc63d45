"""GPU experiment (round 2): dense N=10 kernel with the SIMT sweep (variant 0) against the tensor-core sweep
(variant 1) on config 2 and a config-3 shard: kernel ms, iterations, agreement of the results."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
from gpu_riccati_exp import run
for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)), ("config3_shard", synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))),
                 ("config3_full", synthetic_batch(65536, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0)))):
    a = run(pb, 0, reps=7)
    for v in (1,):            # variant 1 = the other sweep (SIMT since the tensor-core sweep became the default)
        r = run(pb, v, reps=7)
        same = (a["st"] == 1) & (r["st"] == 1)
        print(json.dumps(dict(cfg=name, B=pb.B, simt_ms=a["ms"], tc_ms=r["ms"], speedup=a["ms"] / r["ms"], simt_solves_s=pb.B / a["ms"] * 1e3,
                              tc_solves_s=pb.B / r["ms"] * 1e3, simt_iters=float(a["it"].mean()), tc_iters=float(r["it"].mean()),
                              tc_solved=float((r["st"] == 1).mean()), simt_solved=float((a["st"] == 1).mean()),
                              iters_equal_frac=float((a["it"] == r["it"]).mean()), max_dX=float(np.abs(a["X"][same] - r["X"][same]).max()),
                              max_dU=float(np.abs(a["U"][same] - r["U"][same]).max()))), flush=True)
