"""Batched leg controllers that consume the MPC forces.

The reference turns the MPC's ground-reaction forces into joint torques once per tick in
``Lite3Controller.customPreStep`` (``src/main.py:130-166``): legs the gait marks as stance get
``tau = J' (-f)`` (``ground_controller``, ``src/main.py:193-217``), swing legs follow the
polynomial swing reference of ``src/foot_trajectory_generator.py:27-96`` with a Cartesian PD +
feed-forward law (``swing_leg_controller``, ``src/main.py:219-282``).  The Jacobians, the
mass-matrix rows, the bias forces and the joint velocities come from the rigid-body simulator
(DART in the reference); everything downstream of them is one streaming CUDA kernel here
(``cmpc_leg_torques``), one thread per (robot, leg).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .solver import _ptr


class BatchedLegController:
    """tau[B,4,3] for B robots per call.  ``mpc`` is the :class:`BatchedMPC` whose forces are
    applied (its handle also fixes N and dt); ``gait_tables`` a ``_capi.GaitTables`` describing the
    footstep plans (see :class:`rollout.ClosedLoopRollout` for how one is filled)."""

    def __init__(self, mpc, gait_tables, kp=250.0, kd=15.0):      # src/main.py:48-49
        self.mpc = mpc
        self.gt = gait_tables
        self.kp = (C.c_float * 3)(*np.broadcast_to(np.asarray(kp, dtype=np.float32), (3,)))
        self.kd = (C.c_float * 3)(*np.broadcast_to(np.asarray(kd, dtype=np.float32), (3,)))

    def torques(self, tick, U, J, Jdot, Mleg, cg, dq, foot_pos, foot_vel, out=None, stream=None):
        """tick: int32 device tensor [1]; U [B,N,12]; J, Jdot, Mleg [B,4,3,3]; cg, dq, foot_pos,
        foot_vel [B,4,3] (fp32 CUDA tensors).  Returns (tau [B,4,3], p_des [B,4,3], stance [B] uint8)."""
        import torch
        B = U.shape[0]
        for name, t_, shape in (("U", U, (B, self.mpc.N, 12)), ("J", J, (B, 4, 3, 3)), ("Jdot", Jdot, (B, 4, 3, 3)),
                                ("Mleg", Mleg, (B, 4, 3, 3)), ("cg", cg, (B, 4, 3)), ("dq", dq, (B, 4, 3)),
                                ("foot_pos", foot_pos, (B, 4, 3)), ("foot_vel", foot_vel, (B, 4, 3))):
            if tuple(t_.shape) != shape or t_.dtype != torch.float32 or not t_.is_contiguous() or not t_.is_cuda:
                raise ValueError(f"{name}: expected a contiguous fp32 CUDA tensor of shape {shape}")
        if out is None:
            out = (torch.empty((B, 4, 3), dtype=torch.float32, device=U.device),
                   torch.empty((B, 4, 3), dtype=torch.float32, device=U.device),
                   torch.empty((B,), dtype=torch.uint8, device=U.device))
        tau, p_des, stance = out
        s = torch.cuda.current_stream(U.device).cuda_stream if stream is None else stream
        _capi.check(_capi.lib().cmpc_leg_torques(
            self.mpc._h, B, C.byref(self.gt), _ptr(tick), _ptr(U), _ptr(J), _ptr(Jdot), _ptr(Mleg),
            _ptr(cg), _ptr(dq), _ptr(foot_pos), _ptr(foot_vel), self.kp, self.kd, _ptr(tau),
            _ptr(p_des), _ptr(stance), C.c_void_p(s)))
        return tau, p_des, stance

    def kinematics(self, base_pos, theta, v_base, w_base, q, dq, gravity=-9.81, with_dynamics=True, stream=None):
        """Lite3 leg kinematics on the device (``cmpc_leg_kinematics``): base_pos, theta (torso rotation
        vector), v_base, w_base [B,3]; q, dq [B,4,3] fp32 CUDA tensors.  Returns a dict of device
        tensors foot_pos, foot_vel [B,4,3], J, Jdot [B,4,3,3] and, with ``with_dynamics``, Mleg [B,4,3,3]
        and cg [B,4,3] - exactly the inputs :meth:`torques` takes."""
        import torch
        B = q.shape[0]
        for name, t_, shape in (("base_pos", base_pos, (B, 3)), ("theta", theta, (B, 3)), ("v_base", v_base, (B, 3)),
                                ("w_base", w_base, (B, 3)), ("q", q, (B, 4, 3)), ("dq", dq, (B, 4, 3))):
            if tuple(t_.shape) != shape or t_.dtype != torch.float32 or not t_.is_contiguous() or not t_.is_cuda:
                raise ValueError(f"{name}: expected a contiguous fp32 CUDA tensor of shape {shape}")
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=q.device)
        out = dict(foot_pos=new(B, 4, 3), foot_vel=new(B, 4, 3), J=new(B, 4, 3, 3), Jdot=new(B, 4, 3, 3),
                   Mleg=new(B, 4, 3, 3) if with_dynamics else None, cg=new(B, 4, 3) if with_dynamics else None)
        s = torch.cuda.current_stream(q.device).cuda_stream if stream is None else stream
        _capi.check(_capi.lib().cmpc_leg_kinematics(
            self.mpc._h, B, _ptr(base_pos), _ptr(theta), _ptr(v_base), _ptr(w_base), _ptr(q), _ptr(dq),
            _ptr(out["foot_pos"]), _ptr(out["foot_vel"]), _ptr(out["J"]), _ptr(out["Jdot"]), _ptr(out["Mleg"]),
            _ptr(out["cg"]), C.c_float(gravity), C.c_void_p(s)))
        return out
