"""Kernel time of the default solve kernel on config 2 (4096 trot, N=10) and a config-3 shard (8192 mixed),
median of 30 L2-flushed launches (A/B of build flags such as -DCMPC_FFMA2)."""
import json, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_tail_exp import timeit
for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)),
                 ("config3_shard", synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))),
                 ("n20", synthetic_batch(4096, N=20, seed=0))):
    ms, it = timeit(pb, reps=30)
    print(json.dumps(dict(cfg=name, tag=os.environ.get("TAG", ""), kernel_ms=ms, solves_s=pb.B / ms * 1e3, mean_iters=float(it.mean()))), flush=True)
