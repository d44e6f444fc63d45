"""Lite3 leg kinematics (host side, numpy fp64), the model behind ``cmpc_leg_kinematics``.

The reference reads foot positions / velocities, the legs' linear Jacobians and their time
derivatives, mass-matrix rows and bias forces from DART every tick (reference
``src/main.py:203-214, 236-262, 286-350``).  DART is not available here; the same quantities
follow in closed form from the robot description the reference loads,
``lite3_urdf/urdf/Lite3.urdf``: hip offsets (:45, :143, :240, :337), hip-roll axis ``-x`` (:48),
thigh offset (:73), hip-pitch / knee axis ``-y`` (:76, :103), thigh length 0.20 (:100), shank length
0.21 to the foot frame (:122), link masses and centres of mass (:19-20, :32-33, :54-55, :82-83, :109).
DART gives the inertial-less ``TORSO`` link (:3-15) its default mass of 1 kg; with it the model
reproduces the reference's logged centre of mass of tick 0 (tests/test_leg_kinematics.py), and the
forward kinematics reproduces the logged feet of tick 0 from the initial joint angles of
``src/main.py:67-81``.

Conventions: legs FL, FR, HL, HR; joints (HipX, HipY, Knee) per leg; ``theta`` = rotation vector
of the torso (what ``retrieve_state()['TORSO']['pos']`` holds), world frame everywhere.
"""
from __future__ import annotations

import numpy as np

HIP_X, HIP_Y = 0.1745, 0.062            # Lite3.urdf:45
THIGH_OFF = 0.0985                      # Lite3.urdf:73
L_THIGH, L_SHANK = 0.20, 0.21           # Lite3.urdf:100, :122
SX = np.array([+1.0, +1.0, -1.0, -1.0])  # front / hind
SY = np.array([+1.0, -1.0, +1.0, -1.0])  # left / right
M_TORSO = np.array([4.130, 1.0])        # INERTIA link (:20) and DART's default mass of TORSO
C_TORSO = np.array([[0.004098, -0.000663, -0.002069], [0.0, 0.0, 0.0]])
M_LINKS = np.array([0.428, 0.61, 0.115, 0.01])     # hip, thigh, shank, foot (:33, :55, :83, :109)
TOTAL_MASS = float(M_TORSO.sum() + 4 * M_LINKS.sum())
GRAVITY = np.array([0.0, 0.0, -9.81])
Q_INIT = np.deg2rad([0.0, -60.0, 90.0])             # src/main.py:67-70
BASE_Z_INIT = 0.295 + 0.004                         # src/main.py:81


def link_coms(l):
    """Centres of mass of (hip, thigh, shank, foot) of leg l in their link frames."""
    return np.array([[-SX[l] * 0.0047, -SY[l] * 0.0091, -0.0018],     # :32 / :130 / :227 / :324
                     [-0.00523, -SY[l] * 0.0216, -0.0273],            # :54 / :152 / :249 / :346
                     [0.00585, -8.732e-07, -0.12],                    # :82
                     [0.0, 0.0, 0.0]])


def rotvec_matrix(theta):
    """exp([theta]x) (Rodrigues)."""
    theta = np.asarray(theta, dtype=float)
    a = np.linalg.norm(theta)
    K = np.array([[0, -theta[2], theta[1]], [theta[2], 0, -theta[0]], [-theta[1], theta[0], 0.0]])
    if a < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(a) / a * K + (1 - np.cos(a)) / a ** 2 * K @ K


def _rx(q):          # rotation about the joint axis -x by q
    c, s = np.cos(q), np.sin(q)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]], dtype=float)


def _ry(q):          # rotation about the joint axis -y by q
    c, s = np.cos(q), np.sin(q)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=float)


def leg_frames(l, q):
    """Body-frame joint origins o (3,3), joint axes a (3,3), link rotations R (4 of 3x3: hip, thigh,
    shank, foot) and link-frame origins (4,3) of leg l at joint angles q (3,)."""
    o1 = np.array([SX[l] * HIP_X, SY[l] * HIP_Y, 0.0])
    R1 = _rx(q[0])
    o2 = o1 + R1 @ np.array([0.0, SY[l] * THIGH_OFF, 0.0])
    R2 = R1 @ _ry(q[1])
    o3 = o2 + R2 @ np.array([0.0, 0.0, -L_THIGH])
    R3 = R2 @ _ry(q[2])
    foot = o3 + R3 @ np.array([0.0, 0.0, -L_SHANK])
    axes = np.stack([np.array([-1.0, 0, 0]), R1 @ np.array([0, -1.0, 0]), R2 @ np.array([0, -1.0, 0])])
    return np.stack([o1, o2, o3]), axes, [R1, R2, R3, R3], np.stack([o1, o2, o3, foot])


def leg_kinematics(base_pos, theta, v_base, w_base, q, dq):
    """All legs of one robot.  base_pos, theta (rotation vector), v_base, w_base (3,) world frame;
    q, dq (4,3).  Returns dict: foot_pos, foot_vel (4,3); J, Jdot (4,3,3) world-frame linear Jacobian
    of the foot w.r.t. the leg's joints and its time derivative; Mleg (4,3,3) = sum_i m_i J_com_i
    (the base-translation rows of the joint-space inertia matrix at the leg's columns);
    cg (4,3) gravity torques of the leg's joints (velocity-product terms are not modelled)."""
    Rb = rotvec_matrix(theta)
    base_pos, v_base, w_base = (np.asarray(a, dtype=float) for a in (base_pos, v_base, w_base))
    out = {k: np.zeros((4, 3)) for k in ("foot_pos", "foot_vel", "cg")}
    out.update({k: np.zeros((4, 3, 3)) for k in ("J", "Jdot", "Mleg")})
    for l in range(4):
        o_b, a_b, Rl, org_b = leg_frames(l, q[l])
        o = base_pos + o_b @ Rb.T                     # joint origins, world
        a = a_b @ Rb.T                                # joint axes, world
        p = base_pos + Rb @ org_b[3]
        # velocities of the joint origins / axes: link k-1 carries joint k
        w_link = [w_base]                             # angular velocity of torso, hip, thigh, shank
        for k in range(3):
            w_link.append(w_link[-1] + a[k] * dq[l, k])

        def point_vel(x, upto):                       # velocity of a point fixed in link `upto` (0 = torso)
            v = v_base + np.cross(w_base, x - base_pos)
            for k in range(upto):
                v = v + dq[l, k] * np.cross(a[k], x - o[k])
            return v
        pdot = point_vel(p, 3)
        J = np.stack([np.cross(a[k], p - o[k]) for k in range(3)], axis=1)
        Jd = np.zeros((3, 3))
        for k in range(3):
            adot = np.cross(w_link[k], a[k])
            odot = point_vel(o[k], k)
            Jd[:, k] = np.cross(adot, p - o[k]) + np.cross(a[k], pdot - odot)
        out["foot_pos"][l], out["foot_vel"][l], out["J"][l], out["Jdot"][l] = p, pdot, J, Jd
        coms = link_coms(l)
        Mrow = np.zeros((3, 3))
        for i in range(4):
            c = base_pos + Rb @ (org_b[i] + Rl[i] @ coms[i])
            for k in range(min(i + 1, 3)):            # link i moves with joints 0..min(i,2)
                Mrow[:, k] += M_LINKS[i] * np.cross(a[k], c - o[k])
        out["Mleg"][l] = Mrow
        out["cg"][l] = -Mrow.T @ GRAVITY
    return out


def center_of_mass(base_pos, theta, q):
    Rb = rotvec_matrix(theta)
    num = (M_TORSO[:, None] * C_TORSO).sum(0)
    for l in range(4):
        _, _, Rl, org = leg_frames(l, q[l])
        coms = link_coms(l)
        for i in range(4):
            num = num + M_LINKS[i] * (org[i] + Rl[i] @ coms[i])
    return np.asarray(base_pos, dtype=float) + Rb @ (num / TOTAL_MASS)


def leg_ik(l, foot_body):
    """Joint angles (3,) of leg l that put its foot at `foot_body` (torso frame), knee bent as on the
    robot (Knee > 0, Lite3.urdf:104) and the foot below the hip-pitch axis.  Raises ValueError outside
    the workspace."""
    f = np.asarray(foot_body, dtype=float) - np.array([SX[l] * HIP_X, SY[l] * HIP_Y, 0.0])
    d = SY[l] * THIGH_OFF
    h2 = f[1] ** 2 + f[2] ** 2 - d ** 2
    if h2 < 0:
        raise ValueError("foot inside the hip-roll circle")
    pz = -np.sqrt(h2)
    q1 = np.arctan2(pz, d) - np.arctan2(f[2], f[1])
    q1 = (q1 + np.pi) % (2 * np.pi) - np.pi
    px = f[0]
    c3 = (px ** 2 + pz ** 2 - L_THIGH ** 2 - L_SHANK ** 2) / (2 * L_THIGH * L_SHANK)
    if abs(c3) > 1:
        raise ValueError("foot out of reach")
    q3 = np.arccos(c3)
    q2 = np.arctan2(px, -pz) - np.arctan2(L_SHANK * np.sin(q3), L_THIGH + L_SHANK * np.cos(q3))
    return np.array([q1, q2, q3])


def leg_ik_batch(foot_body):
    """Vectorised :func:`leg_ik`: foot_body (...,4,3) torso-frame feet of FL, FR, HL, HR -> q (...,4,3).
    Unreachable targets are clamped to the workspace boundary."""
    f = np.asarray(foot_body, dtype=float) - np.stack([SX * HIP_X, SY * HIP_Y, np.zeros(4)], -1)
    d = SY * THIGH_OFF
    pz = -np.sqrt(np.maximum(f[..., 1] ** 2 + f[..., 2] ** 2 - d ** 2, 1e-12))
    q1 = np.arctan2(pz, d) - np.arctan2(f[..., 2], f[..., 1])
    q1 = (q1 + np.pi) % (2 * np.pi) - np.pi
    px = f[..., 0]
    c3 = np.clip((px ** 2 + pz ** 2 - L_THIGH ** 2 - L_SHANK ** 2) / (2 * L_THIGH * L_SHANK), -1.0, 1.0)
    q3 = np.arccos(c3)
    q2 = np.arctan2(px, -pz) - np.arctan2(L_SHANK * np.sin(q3), L_THIGH + L_SHANK * np.cos(q3))
    return np.stack([q1, q2, q3], -1)


def nominal_com_offset():
    """Centre of mass in the torso frame at the reference's initial joint angles (src/main.py:67-70)."""
    return center_of_mass(np.zeros(3), np.zeros(3), np.tile(Q_INIT, (4, 1)))
