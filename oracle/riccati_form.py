"""ORACLE (test infrastructure, not product code).

Stage-wise (Riccati) form of the x-update of the CUDA path's ADMM, restated in numpy so that it can
be checked against the dense wrench-space form (oracle/wrench_form.py, itself checked against the
plain recursion of reference src/mpc.py:64-136).

The x-update solves  K dlt = b,  K = D^-1 + G' M G  (D = per-leg 1/(sigma+rho), 0 on swing legs).
M is the Gram matrix of six decoupled double integrators driven by the stage wrenches w_k = G_k u_k
(pos_{k+1} = pos_k + dt vel_k, vel_{k+1} = vel_k + dt w_k, cost sum_{k=1..N} w_pos pos_k^2 + w_vel vel_k^2),
so K dlt = b is an LQ problem with a 12-dim state and  dlt_k = D_k (b_k - G_k' q_k),  q = P^-1 (G D b)
is obtained by one backward and one forward sweep over the stages instead of a dense 6N x 6N solve:

  factor (backward, k = N-1 .. 0):  S_N = Q;  Shat = dt^2 S_vv;  T_k = G_k D_k G_k';
        Phi_k = (I + T_k Shat)^-1,  Gam_k = Phi_k T_k,  L_k = Gam_k B' S_{k+1} A,
        Y_k = A' S_{k+1} B Phi_k,   S_k = Q + A' S_{k+1} (A - B L_k)
  solve:  g_k = Y_k s_k (parallel);  p_k = F_k' p_{k+1} + g_k,  F_k = A - B L_k   (backward)
          w0_k = Phi_k s_k - dt Gam_k p^v_{k+1} (parallel);  xi_{k+1} = F_k xi_k + B w0_k   (forward)
          q_k = B' mu_{k+1},  mu_k = Q xi_k + A' mu_{k+1}   (adjoint of the tracking cost, diagonal)

Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np
from . import srbd_qp, wrench_form as wf


def stage_T(x0, r, stance, d):
    """T_k = G_k D_k G_k' (N,6,6) for per-leg d (N,4) (0 on swing legs)."""
    N = r.shape[0]
    Gh = wf.leg_maps(x0, r)
    T = np.zeros((N, 6, 6))
    for j in range(N):
        for l in range(4):
            if not stance[j, l]:
                continue
            Gp = np.vstack([Gh[j, l], np.eye(3) / srbd_qp.MASS])       # 6x3
            T[j] += d[j, l] * Gp @ Gp.T
    return T


def factor(T, dt, w=srbd_qp.W_STATE, dtype=np.float64):
    N = T.shape[0]
    f = dtype
    Q = np.diag(np.concatenate([2 * w[0:6], 2 * w[6:12]])).astype(f)       # (pos 6, vel 6)
    I6 = np.eye(6, dtype=f)
    A = np.block([[I6, f(dt) * I6], [np.zeros((6, 6), f), I6]]).astype(f)
    B = np.vstack([np.zeros((6, 6), f), f(dt) * I6]).astype(f)
    S = Q.copy()
    out = dict(Phi=np.zeros((N, 6, 6), f), Gam=np.zeros((N, 6, 6), f), L=np.zeros((N, 6, 12), f),
               Y=np.zeros((N, 12, 6), f), A=A, B=B, Q=Q)
    for k in range(N - 1, -1, -1):
        Tk = T[k].astype(f)
        Shat = (B.T @ S @ B).astype(f)
        Phi = np.linalg.inv((I6 + Tk @ Shat).astype(f)).astype(f) if dtype == np.float64 else _inv32(I6 + Tk @ Shat)
        Gam = (Phi @ Tk).astype(f)
        L = (Gam @ (B.T @ S @ A)).astype(f)
        out["Phi"][k], out["Gam"][k], out["L"][k] = Phi, Gam, L
        out["Y"][k] = (A.T @ S @ B @ Phi).astype(f)
        S = ((Q if k >= 1 else 0 * Q) + A.T @ S @ (A - B @ L)).astype(f)
        S = (0.5 * (S + S.T)).astype(f)
    return out


def _inv32(Mx):
    """6x6 Gauss-Jordan in fp32 (no pivoting needed: I + T Shat has a positive spectrum)."""
    n = Mx.shape[0]
    a = np.hstack([Mx.astype(np.float32), np.eye(n, dtype=np.float32)])
    for k in range(n):
        a[k] = a[k] / a[k, k]
        for i in range(n):
            if i != k:
                a[i] = a[i] - a[i, k] * a[k]
    return a[:, n:].astype(np.float32)


def solve(fac, s, dt, dtype=np.float64):
    """q (N,6) = P^-1 s for s (N,6)."""
    f = dtype
    N = s.shape[0]
    A, B, Q = fac["A"], fac["B"], fac["Q"]
    s = s.astype(f)
    g = np.einsum("kij,kj->ki", fac["Y"], s).astype(f)
    p = np.zeros((N + 1, 12), f)
    for k in range(N - 1, -1, -1):
        F = (A - B @ fac["L"][k]).astype(f)
        p[k] = (F.T @ p[k + 1] + g[k]).astype(f)
    w0 = (np.einsum("kij,kj->ki", fac["Phi"], s) - f(dt) * np.einsum("kij,kj->ki", fac["Gam"], p[1:, 6:])).astype(f)
    xi = np.zeros((N + 1, 12), f)
    for k in range(N):
        F = (A - B @ fac["L"][k]).astype(f)
        xi[k + 1] = (F @ xi[k] + B @ w0[k]).astype(f)
    mu = np.zeros((N + 2, 12), f)
    q = np.zeros((N, 6), f)
    for k in range(N, 0, -1):
        mu[k] = (np.diag(Q) * xi[k] + A.T @ mu[k + 1]).astype(f)
        q[k - 1] = (B.T @ mu[k]).astype(f)
    return q, xi
