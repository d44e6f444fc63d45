"""pkl-compatible log export.

Builds, for one robot of a batched closed-loop rollout, a dict with the schema of the
reference's ``Logger.log`` (reference ``src/logger.py:21-46``) and pickles it the way
``Logger.save_log`` does (``src/logger.py:64-66``), so the reference's ``plot.py`` and the
fixture tooling of this repo (``scripts/make_golden_from_pkl.py``) read it unchanged.

Every field the reference logs is filled from device results:
* ``TRACKING PERFORMANCE``, ``FORCES``, ``MPC PREDICTIONS`` (predicted states = the solve kernel's X
  output) as ``MPC.solve`` / ``ground_controller`` log them (``src/mpc.py:295-301``, ``src/main.py:216-218``);
* ``FEET POS`` actual / desired as ``customPreStep`` does (``src/main.py:159-167``): the desired
  position is the plan's foothold for stance legs and the swing reference for swing legs
  (``cmpc_leg_torques``' ``p_des`` output);
* ``CONTROL EFFORT`` (``src/main.py:170-173``, keys ``<leg>_HipX/HipY/Knee`` of ``src/logger.py:39-42``):
  joint torques of ``cmpc_leg_kinematics`` + ``cmpc_leg_torques``.  The rollout's plant is a single
  rigid body (DART is unavailable), so the joint state is reconstructed kinematically: joint angles
  by inverse kinematics of the feet in the torso frame, joint velocities from stationary stance feet
  and swing feet that track their reference.
"""
from __future__ import annotations

import pickle

import numpy as np

from . import kinematics as kin
from .controllers import BatchedLegController
from .gait import LEGS

JOINTS = ("HipX", "HipY", "Knee")


def _joint_state(rollout, ctl, x, feet_world, stance_bits, v_des):
    """Kinematic reconstruction of (base_pos, theta, v, w, q, dq) device tensors for all robots."""
    torch = rollout.torch
    dev = rollout.dev
    B = x.shape[0]
    theta, com, w, v = x[:, 0:3], x[:, 3:6], x[:, 6:9], x[:, 9:12]
    Rb = np.stack([kin.rotvec_matrix(t) for t in theta])                    # (B,3,3)
    base = com - Rb @ kin.nominal_com_offset()
    foot_body = np.einsum("bji,blj->bli", Rb, feet_world - base[:, None, :])
    q = kin.leg_ik_batch(foot_body)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    args = [f32(base), f32(theta), f32(v), f32(w), f32(q)]
    k0 = ctl.kinematics(*args, torch.zeros((B, 4, 3), dtype=torch.float32, device=dev), with_dynamics=False)
    # foot velocity target: 0 for stance feet, the swing reference otherwise;  J dq = target - base part
    stance = ((stance_bits[:, None] >> np.arange(4)[None, :]) & 1).astype(bool)
    target = np.where(stance[..., None], 0.0, v_des)
    rhs = f32(target) - k0["foot_vel"]                                      # foot_vel at dq = 0 = base contribution
    dq = torch.linalg.solve(k0["J"], rhs.unsqueeze(-1)).squeeze(-1).contiguous()
    return args + [dq]


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
           "CONTROL EFFORT": {leg: {f"{leg[:2]}_{j}": [] for j in JOINTS} for leg in LEGS}}
    rollout.out = rollout.mpc.alloc_outputs(rollout.B, want_X=True, device=rollout.dev)   # X of every tick
    ctl = BatchedLegController(rollout.mpc, rollout.gt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tick_now = torch.zeros(1, dtype=torch.int32, device=rollout.dev)
    total_ms = 0.0
    for _ in range(ticks):
        t = int(rollout.tick.item())
        x_all = rollout.x.cpu().numpy().astype(float)
        x_before = x_all[robot]
        e0.record()
        rollout.step()
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        xd = rollout.x_des[robot].cpu().numpy().astype(float)          # (N+1, 13)
        U = rollout.out[0][robot].cpu().numpy().astype(float)          # (N, 12)
        X = rollout.out[1][robot].cpu().numpy().astype(float)          # (N+1, 13)
        feet = rollout.r[:, 0].cpu().numpy().astype(float) + x_all[:, None, 3:6]      # (B,4,3) foot = r0 + com
        # the controller side of the tick (src/main.py:152-173) for all robots, on the device
        tick_now.fill_(t)
        zero = torch.zeros((rollout.B, 4, 3), dtype=torch.float32, device=rollout.dev)
        z9 = torch.zeros((rollout.B, 4, 3, 3), dtype=torch.float32, device=rollout.dev)
        _, p_des, bits = ctl.torques(tick_now, rollout.out[0], z9, z9, z9, zero, zero, zero, zero)   # stance bits, p_des
        # swing velocity reference by differencing the swing position reference one tick ahead
        tick_now.fill_(t + 1)
        _, p_next, _ = ctl.torques(tick_now, rollout.out[0], z9, z9, z9, zero, zero, zero, zero)
        v_des = ((p_next - p_des) / 0.01).cpu().numpy().astype(float)
        tick_now.fill_(t)
        js = _joint_state(rollout, ctl, x_all, feet, bits.cpu().numpy().astype(np.int64), v_des)
        k = ctl.kinematics(*js)
        tau, _, _ = ctl.torques(tick_now, rollout.out[0], k["J"], k["Jdot"], k["Mleg"], k["cg"], js[5],
                                k["foot_pos"], k["foot_vel"])
        tau = tau[robot].cpu().numpy().astype(float)
        pd = p_des[robot].cpu().numpy().astype(float)
        log["time array"].append(t)
        log["TRACKING PERFORMANCE"]["actual"].append(x_before[:12].tolist())
        log["TRACKING PERFORMANCE"]["desired"].append(xd[0, :12].copy())
        for l, leg in enumerate(LEGS):
            log["FEET POS"][leg]["actual"].append(feet[robot, l].copy())
            log["FEET POS"][leg]["des"].append(pd[l].copy())
            for c, ax in enumerate("xyz"):
                log["FORCES"][leg][ax].append(np.float64(U[0, 3 * l + c]))
            for c, j in enumerate(JOINTS):
                log["CONTROL EFFORT"][leg][f"{leg[:2]}_{j}"].append(np.float64(tau[l, c]))
        if t in prediction_ticks:                                       # src/mpc.py:297-301
            log["MPC PREDICTIONS"].append({"time step": t, "predicted_state": X[:, :12].T.copy(),
                                           "desired_state": xd[:, :12].T.copy(),
                                           "predicted forces": U[:, 2::3].T.copy()})
    log["mpc_freq"] = 1e3 * ticks / max(total_ms, 1e-9)
    if path is not None:
        with open(path, "wb") as f:
            pickle.dump(log, f)
    return log
