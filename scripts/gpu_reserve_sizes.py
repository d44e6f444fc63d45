"""GPU experiment (round 2): reserved-SM rank assignment against the plain launch by batch size (env CMPC_NO_RESERVE,
CMPC_HARD_SM read once per process)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
from gpu_riccati_exp import run
tag = dict(no_reserve=os.environ.get("CMPC_NO_RESERVE"), hard_sm=os.environ.get("CMPC_HARD_SM"))
for gaits, mu, gname in ((("trot",), (1.0, 1.0), "trot"), (GAIT_NAMES, (0.3, 1.0), "mixed")):
    for B in (1024, 1536, 2048, 2560, 3072, 3584, 4096, 4440):
        ms = []
        for seed in (0, 1, 2):
            pb = synthetic_batch(B, N=10, gaits=gaits, seed=seed, mu=mu)
            ms.append(run(pb, 0, reps=5)["ms"])
        print(json.dumps(dict(gaits=gname, B=B, **tag, ms_mean_3_seeds=float(np.mean(ms)), ms=ms)), flush=True)
