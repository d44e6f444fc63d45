"""GPU experiment (round 2): shared-memory carve-out of the tensor-core dense kernel (env CMPC_CARVEOUT, percent of
228 KB, read once per process): cold solves of config 2 / 3 and the closed-loop rollout of config 5 (cached factor)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
from gpu_riccati_exp import run
tag = dict(carveout=os.environ.get("CMPC_CARVEOUT"), variant=int(os.environ.get("EXP_VARIANT", "0")))
for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)), ("config3_shard", synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))),
                 ("config3_full", synthetic_batch(65536, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))), ("b2048", synthetic_batch(2048, N=10, seed=1))):
    r = run(pb, tag["variant"], reps=9)
    print(json.dumps(dict(cfg=name, **tag, ms=r["ms"], solves_s=pb.B / r["ms"] * 1e3, iters=float(r["it"].mean()))), flush=True)
if tag["variant"] == 0:
    ro = pkg.ClosedLoopRollout(8192, N=10, gaits=("trot",), mu=(0.3, 1.0), seed=0, warm_mode=1)
    ro.run(21, use_graph=True, ticks_per_graph=20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); done = ro.run(1000, use_graph=True, ticks_per_graph=20); e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(cfg="config5_1000ticks", **tag, ms_per_tick=e0.elapsed_time(e1) / done, robot_ticks_s=8192 * done / e0.elapsed_time(e1) * 1e3)), flush=True)
