"""GPU experiment (round 2): reserved-SM rank assignment of the dense N=10 kernel
(env CMPC_NO_RESERVE / CMPC_HARD_SM are read once per process: run one setting per process)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
from gpu_riccati_exp import run
tag = dict(no_reserve=os.environ.get("CMPC_NO_RESERVE"), hard_sm=os.environ.get("CMPC_HARD_SM"))
for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)), ("config3_shard", synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))),
                 ("config3_full", synthetic_batch(65536, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))), ("b2048", synthetic_batch(2048, N=10, seed=1))):
    r = run(pb, 0, reps=9)
    print(json.dumps(dict(cfg=name, B=pb.B, **tag, ms=r["ms"], solves_s=pb.B / r["ms"] * 1e3, iters=float(r["it"].mean()), solved=float((r["st"] == 1).mean()),
                          checksum=float(np.abs(r["U"]).sum()))), flush=True)
