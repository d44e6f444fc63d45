"""pkl-compatible log export.

Builds, for one robot of a batched closed-loop rollout, a dict with the schema of the
reference's ``Logger.log`` (reference ``src/logger.py:21-46``) and pickles it the way
``Logger.save_log`` does (``src/logger.py:64-66``), so the reference's ``plot.py`` and the
fixture tooling of this repo (``scripts/make_golden_from_pkl.py``) read it unchanged.
Joint torques ("CONTROL EFFORT") need the full-body model and are left empty.
"""
from __future__ import annotations

import pickle

import numpy as np

from .gait import LEGS


def rollout_log(rollout, ticks, robot=0, path=None, prediction_ticks=(0, 80)):
    """Run `ticks` ticks of `rollout` (eagerly, recording robot `robot`) and return the log
    dict; write it to `path` if given."""
    torch = rollout.torch
    N = rollout.N
    params = {"g": -9.81, "h": 0.285, "step_height": float(rollout.plan.step_height),
              "ss_duration": int(rollout.plan.ss[robot]), "ds_duration": int(rollout.plan.ds[robot]),
              "world_time_step": 0.01, "total_steps": int(rollout.gt.total_steps),
              "first_swing": np.asarray(rollout.plan.feet_id[robot, 1]), "µ": float(rollout.mu_host[robot]),
              "N": N, "v_com_ref": rollout.v_ref[robot].cpu().numpy().astype(float),
              "theta_dot": float(rollout.omega_ref[robot]), "log_samples": int(ticks)}
    log = {"mpc_freq": 0.0, "sim_params": params, "total_sim_steps": int(ticks), "time array": [],
           "FEET POS": {leg: {"actual": [], "des": []} for leg in LEGS},
           "MPC PREDICTIONS": [],
           "TRACKING PERFORMANCE": {"actual": [], "desired": []},
           "FORCES": {leg: {"x": [], "y": [], "z": []} for leg in LEGS},
           "CONTROL EFFORT": {leg: {} for leg in LEGS}}
    out_x = rollout.mpc.alloc_outputs(rollout.B, want_X=True, device=rollout.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    for _ in range(ticks):
        t = int(rollout.tick.item())
        x_before = rollout.x[robot].cpu().numpy().astype(float)
        e0.record()
        rollout.step()
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        xd = rollout.x_des[robot].cpu().numpy().astype(float)          # (N+1, 13)
        U = rollout.out[0][robot].cpu().numpy().astype(float)          # (N, 12)
        r0 = rollout.r[robot, 0].cpu().numpy().astype(float)           # (4, 3) foot - com
        log["time array"].append(t)
        log["TRACKING PERFORMANCE"]["actual"].append(x_before[:12].tolist())
        log["TRACKING PERFORMANCE"]["desired"].append(xd[0, :12].copy())
        for l, leg in enumerate(LEGS):
            foot = r0[l] + x_before[3:6]
            log["FEET POS"][leg]["actual"].append(foot)
            log["FEET POS"][leg]["des"].append(list(foot))
            for c, ax in enumerate("xyz"):
                log["FORCES"][leg][ax].append(np.float64(U[0, 3 * l + c]))
        if t in prediction_ticks:
            # predicted states need X: re-solve this tick's problem (same inputs, warm state is
            # not touched because a scratch slot beyond the batch is not available -> reuse U)
            X = np.zeros((13, N + 1))
            X[:, 0] = x_before
            from . import problems  # noqa: F401  (kept local: plain forward Euler, src/mpc.py:113-117)
            yaw = x_before[2]
            c, s = np.cos(yaw), np.sin(yaw)
            Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
            Ihat = Rz @ np.diag([1 / 0.24, 1.0, 1.0]) @ Rz.T
            rr = rollout.r[robot].cpu().numpy().astype(float)
            for i in range(N):
                xi = X[:, i]
                F = U[i].reshape(4, 3)
                tau = np.cross(rr[i], F).sum(0)
                xn = xi.copy()
                xn[0:3] += 0.01 * (Rz @ xi[6:9])
                xn[3:6] += 0.01 * xi[9:12]
                xn[6:9] += 0.01 * (Ihat @ tau)
                xn[9:12] += 0.01 * (F.sum(0) / 8.885 + np.array([0, 0, xi[12]]))
                X[:, i + 1] = xn
            log["MPC PREDICTIONS"].append({"time step": t, "predicted_state": X[:12],
                                           "desired_state": xd[:, :12].T.copy(),
                                           "predicted forces": U[:, 2::3].T.copy()})
    log["mpc_freq"] = 1e3 * ticks / max(total_ms, 1e-9)
    if path is not None:
        with open(path, "wb") as f:
            pickle.dump(log, f)
    return log
