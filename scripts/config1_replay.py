"""BASELINE config 1: the reference's own case - one Lite3 MPC QP per tick at the default horizon
(N = 60, src/main.py:41) on the 1000 states of the reference's committed run
(src/simulation_log.pkl -> tests/golden/simulation_log_golden.npz).

Times the drop-in `MPC.solve(t, logger)` end to end (Python parameter assembly + cmpc_solve_host +
output extraction, warm-started exactly as the reference: previous x unshifted, y = 0) and, beside
it, the reference's CasADi->OSQP path restated in C (oracle/osqp_ref.c, the CPU baseline) on the
same ticks.  Prints one JSON line.  usage: python scripts/config1_replay.py [N] [ticks]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_b200 as pkg
from oracle.replay import ReplayMPC, params_from_golden, initial_from_golden
from oracle.cpu_baseline import OSQPRefC

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "simulation_log_golden.npz"))
gold = {k: gold[k] for k in gold.files}      # NpzFile decompresses on EVERY access: load once


class FakeLite3:
    def __init__(self): self.t = 0
    def retrieve_state(self):
        s, f = gold["state"][self.t], gold["feet"][self.t]
        d = {leg: {"pos": np.concatenate([np.zeros(3), f[l]]), "vel": np.zeros(6)} for l, leg in enumerate(pkg.LEGS)}
        d["TORSO"] = {"pos": s[0:3].copy(), "vel": s[6:9].copy()}
        d["com"] = {"pos": s[3:6].copy(), "vel": s[9:12].copy()}
        return d


class Logger:
    def log_tracking_data(self, a, d): pass
    def log_mpc_predictions(self, *a): pass


params = params_from_golden(gold, N=N)
initial = initial_from_golden(gold)
gp = pkg.GaitPlan.from_initial(initial, params)
lite3, logger = FakeLite3(), Logger()
mpc = pkg.MPC(lite3=lite3, initial=initial_from_golden(gold), footstep_planner=gp, params=params)
ms, its = [], []
for t in range(T):
    lite3.t = t
    t0 = time.perf_counter()
    mpc.solve(t, logger)
    ms.append(1e3 * (time.perf_counter() - t0))
    its.append(mpc.iters)
ms, its = np.array(ms[5:]), np.array(its)

# the reference's OSQP path in C on the same ticks (warm start and rho persist as in CasADi)
rep = ReplayMPC(initial_from_golden(gold), params)
c = OSQPRefC(N)
warm, cms, cit = None, [], []
for t in range(T):
    x0, r, stance, xd, v, om = rep.tick_problem(t, gold["state"][t], gold["feet"][t])
    t0 = time.perf_counter()
    sol, st, it, rho = c.solve(x0, r, (1 - stance).T.astype(float), xd, params["µ"], 0.01, params["g"], warm)
    cms.append(1e3 * (time.perf_counter() - t0))
    cit.append(it)
    rep.com_pos_start = rep.com_pos_start + v * 0.01
    rep.yaw_start = rep.yaw_start + om * 0.01
    warm = sol
cms = np.array(cms[5:])
print(json.dumps({
    "config": f"config 1: single Lite3 MPC QP per tick, N={N}, {T} logged ticks of the reference's run, warm-started",
    "dropin_ms_per_tick_p50": float(np.median(ms)), "dropin_ms_per_tick_mean": float(ms.mean()),
    "dropin_ms_per_tick_p99": float(np.percentile(ms, 99)), "dropin_iters_mean": float(its.mean()),
    "dropin_iters_max": int(its.max()), "dropin_hz_p50": float(1e3 / np.median(ms)),
    "c_osqp_port_ms_per_solve_p50": float(np.median(cms)), "c_osqp_port_ms_per_solve_mean": float(cms.mean()),
    "c_osqp_port_iters_mean": float(np.mean(cit)),
    "note": "drop-in time is the whole MPC.solve call (numpy parameter assembly + GPU solve + extraction); "
            "the C port time is the QP solve only (the reference adds ~5.5 ms of Python assembly at N=60 and "
            "reports 16.3 ms per tick overall, SURVEY.md section 6)"}))
