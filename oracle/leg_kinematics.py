"""ORACLE (test infrastructure, not product code).

Independent fp64 restatement of the Lite3 leg kinematics that the reference obtains from DART
(``src/main.py:203-214, 236-262, 286-350``), built straight from the joint tree of
``lite3_urdf/urdf/Lite3.urdf`` (:44-49, :72-77, :99-104, :121-125 and the three other legs) by
composing homogeneous transforms with scipy rotations; Jacobians and their time derivatives are
obtained NUMERICALLY (central differences of the forward kinematics / of the Jacobian along the
motion), i.e. by a different route than the closed forms of the product (kinematics.py, the CUDA
kernel).  Pinned on the reference's logged run: the feet and the centre of mass of tick 0
(``simulation_log.pkl``) from the initial configuration of ``src/main.py:67-81``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as Rot

# (joint origin xyz, axis) per leg, from the URDF; then the fixed ankle offset
LEGS = ("FL", "FR", "HL", "HR")
TREE = {
    "FL": [((0.1745, 0.062, 0.0), (-1, 0, 0)), ((0, 0.0985, 0), (0, -1, 0)), ((0, 0, -0.20), (0, -1, 0)), (0, 0, -0.21)],
    "FR": [((0.1745, -0.062, 0.0), (-1, 0, 0)), ((0, -0.0985, 0), (0, -1, 0)), ((0, 0, -0.20), (0, -1, 0)), (0, 0, -0.21)],
    "HL": [((-0.1745, 0.062, 0.0), (-1, 0, 0)), ((0, 0.0985, 0), (0, -1, 0)), ((0, 0, -0.20), (0, -1, 0)), (0, 0, -0.21)],
    "HR": [((-0.1745, -0.062, 0.0), (-1, 0, 0)), ((0, -0.0985, 0), (0, -1, 0)), ((0, 0, -0.20), (0, -1, 0)), (0, 0, -0.21)],
}
LINK_COM = {   # hip, thigh, shank (foot: origin), URDF inertial origins
    "FL": [(-0.0047, -0.0091, -0.0018), (-0.00523, -0.0216, -0.0273), (0.00585, -8.732e-07, -0.12)],
    "FR": [(-0.0047, 0.0091, -0.0018), (-0.00523, 0.0216, -0.0273), (0.00585, -8.732e-07, -0.12)],
    "HL": [(0.0047, -0.0091, -0.0018), (-0.00523, -0.0216, -0.0273), (0.00585, -8.732e-07, -0.12)],
    "HR": [(0.0047, 0.0091, -0.0018), (-0.00523, 0.0216, -0.0273), (0.00585, -8.732e-07, -0.12)],
}
LINK_MASS = (0.428, 0.61, 0.115, 0.01)
TORSO = ((4.130, (0.004098, -0.000663, -0.002069)), (1.0, (0.0, 0.0, 0.0)))   # INERTIA link; DART default for TORSO


def foot_and_coms(leg, base_pos, theta, q):
    """World position of the foot and of the 4 link centres of mass of one leg."""
    R = Rot.from_rotvec(theta).as_matrix()
    p = np.asarray(base_pos, dtype=float)
    coms = []
    for k in range(3):
        org, axis = TREE[leg][k]
        p = p + R @ np.asarray(org, dtype=float)
        R = R @ Rot.from_rotvec(np.asarray(axis, dtype=float) * q[k]).as_matrix()
        coms.append(p + R @ np.asarray(LINK_COM[leg][k]))
    foot = p + R @ np.asarray(TREE[leg][3], dtype=float)
    coms.append(foot)
    return foot, coms


def foot_position(base_pos, theta, q):
    return np.stack([foot_and_coms(leg, base_pos, theta, q[l])[0] for l, leg in enumerate(LEGS)])


def center_of_mass(base_pos, theta, q):
    R = Rot.from_rotvec(theta).as_matrix()
    m, num = 0.0, np.zeros(3)
    for mass, c in TORSO:
        m += mass
        num += mass * (np.asarray(base_pos) + R @ np.asarray(c))
    for l, leg in enumerate(LEGS):
        _, coms = foot_and_coms(leg, base_pos, theta, q[l])
        for mass, c in zip(LINK_MASS, coms):
            m += mass
            num += mass * c
    return num / m, m


def numeric_jacobian(base_pos, theta, q, h=1e-6):
    """(4,3,3): d foot_l / d q_l by central differences."""
    J = np.zeros((4, 3, 3))
    for l, leg in enumerate(LEGS):
        for k in range(3):
            qp, qm = np.array(q[l], dtype=float), np.array(q[l], dtype=float)
            qp[k] += h
            qm[k] -= h
            J[l][:, k] = (foot_and_coms(leg, base_pos, theta, qp)[0] - foot_and_coms(leg, base_pos, theta, qm)[0]) / (2 * h)
    return J


def numeric_mass_rows(base_pos, theta, q, h=1e-6):
    """(4,3,3): sum_i m_i d com_i / d q_l."""
    M = np.zeros((4, 3, 3))
    for l, leg in enumerate(LEGS):
        for k in range(3):
            qp, qm = np.array(q[l], dtype=float), np.array(q[l], dtype=float)
            qp[k] += h
            qm[k] -= h
            cp, cm = foot_and_coms(leg, base_pos, theta, qp)[1], foot_and_coms(leg, base_pos, theta, qm)[1]
            M[l][:, k] = sum(m * (a - b) for m, a, b in zip(LINK_MASS, cp, cm)) / (2 * h)
    return M


def advance(base_pos, theta, v_base, w_base, q, dq, h):
    """Configuration after a time h of constant velocities (world-frame angular velocity)."""
    R = Rot.from_rotvec(w_base * h) * Rot.from_rotvec(theta)
    return base_pos + v_base * h, R.as_rotvec(), q + dq * h


def numeric_rates(base_pos, theta, v_base, w_base, q, dq, h=1e-5):
    """foot velocity (4,3) and Jdot (4,3,3) by central differences along the motion."""
    a = advance(base_pos, theta, v_base, w_base, q, dq, +h)
    b = advance(base_pos, theta, v_base, w_base, q, dq, -h)
    vel = (foot_position(*a) - foot_position(*b)) / (2 * h)
    Jd = (numeric_jacobian(*a) - numeric_jacobian(*b)) / (2 * h)
    return vel, Jd
