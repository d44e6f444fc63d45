"""Per-tick QP parameter assembly on the host (fp64, numpy).

Mirrors the part of the reference's ``MPC.solve`` that turns (tick, measured state,
reference accumulators, gait plan) into the numeric QP parameters
(reference ``src/mpc.py:178-255``): the desired trajectory ``x_des``, the per-stage
lever arms ``r`` and the contact schedule.  The reference does this with N x 4 Python
calls per tick; here one vectorised evaluation of the :class:`GaitPlan` serves the whole
horizon, and the batched variant serves a whole batch of robots.
"""
from __future__ import annotations

import numpy as np

from .gait import GaitPlan, stance_bits


def desired_trajectory(N, delta, roll, pitch, yaw_start, com_start, v_ref, omega_ref, g):
    """x_des (13, N+1) as reference ``src/mpc.py:202-214``: constant roll/pitch, yaw and
    CoM integrated from the *reference accumulators* with sequential additions (kept
    sequential so the fp64 values match the reference's loop exactly)."""
    xd = np.zeros((13, N + 1))
    xd[0, :] = roll
    xd[1, :] = pitch
    xd[8, :] = omega_ref
    xd[9:12, :] = np.asarray(v_ref, dtype=float).reshape(3, 1)
    xd[12, :] = g
    inc_yaw = np.full(N + 1, omega_ref * delta)
    inc_yaw[0] = yaw_start
    xd[2, :] = np.cumsum(inc_yaw)
    inc_com = np.tile((np.asarray(v_ref, dtype=float) * delta).reshape(3, 1), (1, N + 1))
    inc_com[:, 0] = com_start
    xd[3:6, :] = np.cumsum(inc_com, axis=1)
    return xd


def assemble_tick(plan: GaitPlan, t: int, N: int, delta: float, x0, feet_now, x_des):
    """Lever arms and contact schedule of one tick.

    feet_now (4,3): measured foot positions; x0 (13,) measured state.
    Returns r (N,4,3) fp64 and stance (N,4) int (1 = stance).
    reference ``src/mpc.py:218-239`` (lever arms) and ``249-252`` (mask)."""
    x0 = np.asarray(x0, dtype=float).reshape(-1)
    r = np.empty((N, 4, 3))
    r[0] = np.asarray(feet_now, dtype=float) - x0[3:6]
    if N > 1:
        ticks = t + np.arange(1, N)
        r[1:] = plan.foot_position(ticks) - x_des[3:6, 1:N].T[:, None, :]
    stance = plan.stance_mask(t + np.arange(N))
    return r, stance


def reference_velocity(plan: GaitPlan, t: int, params):
    """v_ref, omega_ref of this tick: zeroed during the last planned step
    (reference ``src/mpc.py:178-183``)."""
    v = np.asarray(params["v_com_ref"], dtype=float)
    om = params["theta_dot"]
    if int(plan.step_index(t)) == params["total_steps"] - 1:
        return v * 0, om * 0
    return v, om


def pack_problem(x0, r, stance, x_des):
    """fp64 single-problem arrays -> the fp32 / uint8 records of the C ABI
    (x0[13], r[N,4,3], mask[N], x_des[N+1,13])."""
    return (np.asarray(x0, dtype=np.float32).reshape(13),
            np.asarray(r, dtype=np.float32),
            stance_bits(stance),
            np.ascontiguousarray(np.asarray(x_des, dtype=np.float32).T))
