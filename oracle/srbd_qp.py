"""ORACLE (test infrastructure, not product code).

fp64 CPU restatement of the QP that the reference's ``MPC.__init__`` builds with
CasADi ``Opti('conic')`` (reference ``src/mpc.py:49-173``) in the exact form it is
handed to OSQP, plus the closed-form condensed QP
 that the CUDA path solves.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.

Conventions (reference ``src/mpc.py:58-61``):
  U in R^{12 x N}: column i = stage-i forces, legs FL,FR,HL,HR, xyz each.
  X in R^{13 x (N+1)} = [Theta(3) p(3) omega(3) v(3) g].
  sparse variable vector  zeta = [vec(U); vec(X)]  (column-major):
     col(U[k,i]) = 12 i + k,  col(X[k,i]) = 12 N + 13 i + k.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# reference src/mpc.py:121-134 (state weights; force weight is 0.0)
W_STATE = np.array([1e4, 2.7e4, 1e4, 2.7e5, 2.7e5, 2.7e5,
                    1e4, 1e4, 1e4, 1.6e4, 1.6e4, 1.6e4, 0.0])
MASS = 8.885                                  # src/mpc.py:71
IBODY_INV = np.array([1 / 0.24, 1.0, 1.0])    # src/mpc.py:73-76
F_MIN, F_MAX = 3.0, 100.0                     # src/mpc.py:45-46
INF = np.inf


def skew(v):
    """reference src/utils.py:43-56"""
    return np.array([[0.0, -v[2], v[1]],
                     [v[2], 0.0, -v[0]],
                     [-v[1], v[0], 0.0]])


def rot_z(yaw):
    """reference src/mpc.py:64-69"""
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def continuous_A(yaw):
    """13x13 continuous-time A (reference src/mpc.py:86-96)."""
    A = np.zeros((13, 13))
    A[0:3, 6:9] = rot_z(yaw)
    A[3:6, 9:12] = np.eye(3)
    A[11, 12] = 1.0
    return A


def continuous_B(yaw, r_stage, mass=MASS, ibody_inv=IBODY_INV):
    """13x12 continuous-time B of one stage (reference src/mpc.py:98-107).
    r_stage: (4,3) lever arms foot - com."""
    Rz = rot_z(yaw)
    I_hat_inv = Rz @ np.diag(ibody_inv) @ Rz.T      # src/mpc.py:78
    B = np.zeros((13, 12))
    for l in range(4):
        B[6:9, 3 * l:3 * l + 3] = I_hat_inv @ skew(r_stage[l])
        B[9:12, 3 * l:3 * l + 3] = np.eye(3) / mass
    return B


class SparseQP:
    """min 1/2 z'Pz + q'z  s.t.  l <= A z <= u   in the order CasADi gives OSQP:
    A = [I_n ; A_g] (variable-bound identity block first, all infinite)."""

    def __init__(self, N):
        self.N = N
        self.nU = 12 * N
        self.nX = 13 * (N + 1)
        self.n = self.nU + self.nX
        self.mg = 13 + 66 * N
        self.m = self.n + self.mg

    # index helpers -----------------------------------------------------
    def cu(self, k, i):
        return 12 * i + k

    def cx(self, k, i):
        return self.nU + 13 * i + k


def build_sparse_qp(x0, r, swing, x_des, mu, delta, g,
                    w=W_STATE, mass=MASS, ibody_inv=IBODY_INV,
                    f_min=F_MIN, f_max=F_MAX):
    """Numeric QP of one tick.

    x0 (13,), r (N,4,3) lever arms, swing (4,N) in {0,1} (1 = swing; this is the
    reference's ``swing_param``), x_des (13,N+1), mu friction.
    Returns (qp, P_diag (n,), q (n,), A (m x n CSC, identity block first), l, u).
    """
    N = r.shape[0]
    qp = SparseQP(N)
    n, mg = qp.n, qp.mg
    yaw = x0[2]
    Ac = continuous_A(yaw)
    Ad = np.eye(13) + delta * Ac

    rows, cols, vals = [], [], []
    lb = np.zeros(mg)
    ub = np.zeros(mg)

    def add(rw, cl, v):
        rows.append(rw)
        cols.append(cl)
        vals.append(v)

    # initial state X[:,0] == x0                         (src/mpc.py:113)
    for k in range(13):
        add(k, qp.cx(k, 0), 1.0)
        lb[k] = ub[k] = x0[k]
    # dynamics  X_{i+1} - X_i - delta*(A X_i + B_i U_i) == 0  (src/mpc.py:116-117)
    for i in range(N):
        Bd = delta * continuous_B(yaw, r[i], mass, ibody_inv)
        base = 13 + 13 * i
        for k in range(13):
            add(base + k, qp.cx(k, i + 1), 1.0)
            for kk in range(13):
                if Ad[k, kk] != 0.0:
                    add(base + k, qp.cx(kk, i), -Ad[k, kk])
            for kk in range(12):
                if Bd[k, kk] != 0.0:
                    add(base + k, qp.cu(kk, i), -Bd[k, kk])
    # swing equality  swing[l,i]*U[3l:3l+3,i] == 0        (src/mpc.py:139-144)
    base = 13 + 13 * N
    for i in range(N):
        for l in range(4):
            for k in range(3):
                rw = base + 12 * i + 3 * l + k
                if swing[l, i] != 0.0:
                    add(rw, qp.cu(3 * l + k, i), float(swing[l, i]))
    # per-stage block of 41 rows                          (src/mpc.py:148-173)
    for i in range(N):
        base = 13 + 25 * N + 41 * i
        add(base, qp.cx(12, i), 1.0)                       # X[12,i] == g
        lb[base] = ub[base] = g
        for l in range(4):
            cond = 1.0 - swing[l, i]
            fz = qp.cu(3 * l + 2, i)
            r_lo, r_hi = base + 1 + 2 * l, base + 2 + 2 * l
            if cond != 0.0:
                add(r_lo, fz, cond)
                add(r_hi, fz, cond)
            lb[r_lo], ub[r_lo] = cond * f_min, INF         # cond*f_min <= cond*fz
            lb[r_hi], ub[r_hi] = -INF, cond * f_max        # cond*fz <= cond*f_max
        for blk, comp in ((9, 1), (25, 0)):                # fy rows then fx rows
            for l in range(4):
                ft = qp.cu(3 * l + comp, i)
                fz = qp.cu(3 * l + 2, i)
                rb = base + blk + 4 * l
                # source order (src/mpc.py:161-165 / 169-173), each as expr <= 0
                for j, sgn in enumerate((-1.0, 1.0, 1.0, -1.0)):
                    add(rb + j, ft, sgn)
                    add(rb + j, fz, -mu)
                    lb[rb + j], ub[rb + j] = -INF, 0.0
    Ag = sp.csc_matrix((vals, (rows, cols)), shape=(mg, n))
    A = sp.vstack([sp.identity(n, format="csc"), Ag], format="csc")
    l = np.concatenate([np.full(n, -INF), lb])
    u = np.concatenate([np.full(n, INF), ub])

    # cost  sum_k sum_j w_j (X[j,k]-x_des[j,k])^2           (src/mpc.py:120-136)
    P_diag = np.zeros(n)
    q = np.zeros(n)
    for i in range(N + 1):
        for k in range(13):
            P_diag[qp.cx(k, i)] = 2.0 * w[k]
            q[qp.cx(k, i)] = -2.0 * w[k] * x_des[k, i]
    return qp, P_diag, q, A, l, u


# ----------------------------------------------------------------------
# Condensed form.  Unknowns = stance leg-stage forces only.
# ----------------------------------------------------------------------

def stance_index(stance):
    """stance (N,4) in {0,1} -> list of (stage, leg) of the compact unknown order
    (stage-major, leg-minor) used by the CUDA path."""
    N = stance.shape[0]
    return [(i, l) for i in range(N) for l in range(4) if stance[i, l]]


def free_response(x0, N, delta):
    """c0 (13,N+1): trajectory with zero forces, by the plain recursion."""
    Ad = np.eye(13) + delta * continuous_A(x0[2])
    c0 = np.zeros((13, N + 1))
    c0[:, 0] = x0
    for k in range(N):
        c0[:, k + 1] = Ad @ c0[:, k]
    return c0


def prediction_matrix(x0, r, delta, mass=MASS, ibody_inv=IBODY_INV):
    """S (13(N+1) x 12N), X = c0 + S u by the plain recursion (generic, O(N^2) blocks)."""
    N = r.shape[0]
    yaw = x0[2]
    Ad = np.eye(13) + delta * continuous_A(yaw)
    S = np.zeros((13 * (N + 1), 12 * N))
    for j in range(N):
        blk = delta * continuous_B(yaw, r[j], mass, ibody_inv)
        for k in range(j + 1, N + 1):
            S[13 * k:13 * k + 13, 12 * j:12 * j + 12] = blk
            blk = Ad @ blk
    return S


def condensed_qp(x0, r, stance, x_des, delta, w=W_STATE, mass=MASS,
                 ibody_inv=IBODY_INV, r_weight=0.0):
    """Dense condensed QP over the stance unknowns:
        min 1/2 u'Hu + g'u ,  X = c0 + S_c u,
    H = 2 S_c' Qbar S_c + 2 r_weight I, gvec = 2 S_c' Qbar (c0 - x_des).
    Returns H (n,n), gvec (n,), S_c (13(N+1), n), c0 (13,N+1), idx list."""
    N = r.shape[0]
    idx = stance_index(stance)
    colsel = np.array([12 * i + 3 * l + k for (i, l) in idx for k in range(3)], dtype=int)
    S = prediction_matrix(x0, r, delta, mass, ibody_inv)
    Sc = S[:, colsel] if len(colsel) else np.zeros((13 * (N + 1), 0))
    c0 = free_response(x0, N, delta)
    Qbar = np.tile(w, N + 1)
    e0 = (c0 - x_des).T.reshape(-1)         # stage-major stacking matches S rows
    H = 2.0 * Sc.T @ (Qbar[:, None] * Sc) + 2.0 * r_weight * np.eye(len(colsel))
    gvec = 2.0 * Sc.T @ (Qbar * e0)
    return H, gvec, Sc, c0, idx


def constraint_rows(n_legs, mu):
    """Per stance leg 5 rows:  fz in [f_min,f_max];  +-fx - mu fz <= 0;  +-fy - mu fz <= 0."""
    A = np.zeros((5 * n_legs, 3 * n_legs))
    for s in range(n_legs):
        A[5 * s + 0, 3 * s + 2] = 1.0
        A[5 * s + 1, 3 * s + 0], A[5 * s + 1, 3 * s + 2] = 1.0, -mu
        A[5 * s + 2, 3 * s + 0], A[5 * s + 2, 3 * s + 2] = -1.0, -mu
        A[5 * s + 3, 3 * s + 1], A[5 * s + 3, 3 * s + 2] = 1.0, -mu
        A[5 * s + 4, 3 * s + 1], A[5 * s + 4, 3 * s + 2] = -1.0, -mu
    return A


def objective(X, x_des, w=W_STATE):
    """J = sum_k sum_j w_j (X-x_des)^2 over k=0..N (reference src/mpc.py:121-134)."""
    return float(np.sum(w[:, None] * (X - x_des) ** 2))


def stage_wrench(U, r):
    """Per-stage net wrench [sum f ; sum r x f]  (the unique part of the force solution).
    U (N,12) or (12,N)->use (N,12); r (N,4,3).  Returns (N,6)."""
    N = r.shape[0]
    F = U.reshape(N, 4, 3)
    return np.concatenate([F.sum(1), np.cross(r, F).sum(1)], axis=1)
