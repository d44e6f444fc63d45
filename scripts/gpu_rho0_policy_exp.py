"""GPU experiment (round 2): would a per-problem initial rho seeded from the LPT conditioning score shorten the
hardest problems?  Emulated with the existing API: the batch is solved once per candidate rho0
and the per-problem iteration counts are combined by score threshold.  Score = mean diagonal of H over the
first- and last-stage stance variables (what score_kernel computes), here from the fp64 condensed H."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, DT, GAIT_NAMES
from oracle import srbd_qp
from gpu_riccati_exp import run
for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)), ("config3_shard", synthetic_batch(4096, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0)))):
    score = np.zeros(pb.B)
    for b in range(pb.B):
        x0, r, st, xd, mu = pb.problem(b)
        H, g, Sc, c0, idx = srbd_qp.condensed_qp(x0, r, st, xd, DT)
        d = np.diag(H)
        sel = [3 * s + c for s, (i, l) in enumerate(idx) if i in (0, pb.N - 1) for c in range(3)]
        score[b] = d[sel].mean() if sel else 0.0
    its = {}
    for rho0 in (0.5, 1.0, 2.0, 4.0, 8.0, 16.0):
        r_ = run(pb, 0, reps=1, rho=rho0, rho_min=0.05)
        its[rho0] = r_["it"].astype(float)
        assert (r_["st"] == 1).all()
    base = its[0.5]
    print(json.dumps(dict(cfg=name, policy="rho0=0.5 for all", mean=base.mean(), max=base.max(), top8=np.sort(base)[-8:].tolist(),
                          score_pct=np.percentile(score, [10, 50, 90, 97, 99, 100]).round(2).tolist(),
                          corr_iters_score=float(np.corrcoef(base, score)[0, 1]))), flush=True)
    for q in (90, 95, 97, 99):
        T = np.percentile(score, q)
        for rho_hi in (1.0, 2.0, 4.0, 8.0, 16.0):
            it = np.where(score > T, its[rho_hi], base)
            print(json.dumps(dict(cfg=name, policy=f"rho0={rho_hi} above the {q}th score percentile ({T:.1f})", mean=it.mean(), max=it.max(),
                                  top8=np.sort(it)[-8:].tolist())), flush=True)
