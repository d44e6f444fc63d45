"""ORACLE (test infrastructure, not product code).

Replays the reference's ``MPC.solve`` tick by tick (reference ``src/mpc.py:176-303``)
on recorded inputs, with :class:`oracle.osqp_ref.OSQPRef` standing in for CasADi+OSQP.
State that persists across ticks is exactly the reference's: the primal warm start
(src/mpc.py:270-271), the reference accumulators (src/mpc.py:261-262) and OSQP's rho.
"""
from __future__ import annotations

import importlib
import numpy as np

from . import srbd_qp
from .osqp_ref import OSQPRef

_pkg = importlib.import_module("mpc-for-dynamic-locomotion-in-the-mit-cheetah-3_b200")


def params_from_golden(gold, N=None):
    return {
        "g": float(gold["g"]), "h": float(gold["h"]), "step_height": float(gold["step_height"]),
        "ss_duration": int(gold["ss_duration"]), "ds_duration": int(gold["ds_duration"]),
        "world_time_step": float(gold["world_time_step"]), "total_steps": int(gold["total_steps"]),
        "first_swing": np.asarray(gold["first_swing"]), "µ": float(gold["mu"]),
        "N": int(gold["N"]) if N is None else N,
        "v_com_ref": np.asarray(gold["v_com_ref"], dtype=float), "theta_dot": float(gold["theta_dot"]),
    }


def initial_from_golden(gold):
    s0 = gold["state"][0]
    ini = {leg: gold["feet"][0, l].copy() for l, leg in enumerate(_pkg.LEGS)}
    ini.update(roll=s0[0], pitch=s0[1], yaw=s0[2], com_position=s0[3:6].copy())
    return ini


class ReplayMPC:
    def __init__(self, initial, params, solver=None):
        self.params = params
        self.N = params["N"]
        self.delta = params["world_time_step"]
        self.mu = params["µ"]
        self.initial = initial
        self.plan = _pkg.GaitPlan.from_initial(initial, params)
        self.com_pos_start = np.array(initial["com_position"], dtype=float)
        self.com_pos_start[2] = params["h"]                       # src/mpc.py:36-37
        self.yaw_start = initial["yaw"]
        self.solver = solver if solver is not None else OSQPRef()
        self.warm = None
        self.last = {}

    def tick_problem(self, t, state12, feet):
        p = self.params
        v, om = _pkg.reference_velocity(self.plan, t, p)
        x0 = np.concatenate([np.asarray(state12, dtype=float), [p["g"]]])
        xd = _pkg.desired_trajectory(self.N, self.delta, self.initial["roll"], self.initial["pitch"],
                                     self.yaw_start, self.com_pos_start, v, om, p["g"])
        r, stance = _pkg.assemble_tick(self.plan, t, self.N, self.delta, x0, feet, xd)
        return x0, r, stance, xd, v, om

    def solve(self, t, state12, feet):
        x0, r, stance, xd, v, om = self.tick_problem(t, state12, feet)
        swing = (1 - stance).T.astype(float)                        # (4,N) swing_param
        qp, Pd, q, A, l, u = srbd_qp.build_sparse_qp(x0, r, swing, xd, self.mu, self.delta,
                                                     self.params["g"])
        sol, status = self.solver.solve(Pd, q, A, l, u, self.warm)
        if status != "solved":
            raise RuntimeError("OSQP restatement did not reach 'solved'")   # CasADi raises too
        self.com_pos_start = self.com_pos_start + v * self.delta    # src/mpc.py:261-262
        self.yaw_start = self.yaw_start + om * self.delta
        self.warm = sol                                             # unshifted primal warm start
        N = self.N
        U = sol[:12 * N].reshape(N, 12).T
        X = sol[12 * N:].reshape(N + 1, 13).T
        self.last = dict(U=U, X=X, x_des=xd, r=r, stance=stance, x0=x0,
                         iters=self.solver.info["iters"], rho=self.solver.info["rho"])
        return U[:, 0].copy()
