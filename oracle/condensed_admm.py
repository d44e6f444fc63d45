"""ORACLE (test infrastructure, not product code).

``oracle_admm64`` / ``oracle_tight`` of SURVEY.md section 7: the algorithm of the CUDA
path - OSQP-style ADMM (Stellato et al. 2020, Algorithm 1) on the condensed QP over the
stance forces - restated in numpy so it can run in fp64 at any
tolerance.  The QP itself comes from :func:`oracle.srbd_qp.condensed_qp`, which follows
reference ``src/mpc.py:64-173`` (dynamics, cost, swing/bound/friction constraints).

    minimise 1/2 u'Hu + g'u   s.t.  l <= A u <= ub,
    A: 5 rows per stance leg  [fz ; fx-mu fz ; -fx-mu fz ; fy-mu fz ; -fy-mu fz]
    l = [f_min,-inf,...], ub = [f_max,0,0,0,0]

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.
"""
from __future__ import annotations

import numpy as np

from . import srbd_qp


def bounds(n_legs, f_min=srbd_qp.F_MIN, f_max=srbd_qp.F_MAX):
    lo = np.tile(np.array([f_min, -np.inf, -np.inf, -np.inf, -np.inf]), n_legs)
    hi = np.tile(np.array([f_max, 0.0, 0.0, 0.0, 0.0]), n_legs)
    return lo, hi


def admm(H, g, mu, rho=0.3, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
         max_iter=1000, check_every=5, x=None, y=None, f_min=srbd_qp.F_MIN,
         f_max=srbd_qp.F_MAX, dtype=np.float64, fixed_iters=None, adaptive_interval=0,
         adaptive_tolerance=5.0):
    """OSQP-style ADMM with a single rho, no scaling.  Returns dict(x, y, z, iters,
    pri_res, dua_res, status).  ``fixed_iters`` runs exactly that many iterations (for
    iterate-level parity with the CUDA kernel).  ``adaptive_interval`` > 0 applies OSQP's rho
    adaptation rule every that many iterations (refactorising K), as the CUDA kernel does."""
    n = H.shape[0]
    S = n // 3
    dt = dtype
    H = H.astype(dt)
    g = g.astype(dt)
    A = srbd_qp.constraint_rows(S, mu).astype(dt)
    lo, hi = bounds(S, f_min, f_max)
    lo, hi = lo.astype(dt), hi.astype(dt)
    x = np.zeros(n, dt) if x is None else x.astype(dt)
    y = np.zeros(5 * S, dt) if y is None else y.astype(dt)
    z = np.clip(A @ x, lo, hi)
    rho, sigma, alpha = dt(rho), dt(sigma), dt(alpha)
    AtA = A.T @ A

    def factor(rho_):
        K = H + sigma * np.eye(n, dtype=dt) + rho_ * AtA
        return np.linalg.cholesky(K.astype(np.float64)).astype(dt) if dt == np.float64 \
            else _chol32(K)
    if n:
        L = factor(rho)
    rho_updates = 0
    status, it = 0, 0
    pri = dua = dt(0)
    total = fixed_iters if fixed_iters is not None else max_iter
    for it in range(1, total + 1):
        rhs = sigma * x - g + A.T @ (rho * z - y)
        xt = _chol_solve(L, rhs) if n else rhs
        zt = A @ xt
        x = alpha * xt + (1 - alpha) * x
        zh = alpha * zt + (1 - alpha) * z
        zn = np.clip(zh + y / rho, lo, hi)
        y = y + rho * (zh - zn)
        z = zn
        adapt = adaptive_interval > 0 and it % adaptive_interval == 0
        if (fixed_iters is None and it % check_every == 0) or adapt:
            Ax = A @ x
            Hx = H @ x
            Aty = A.T @ y
            pri = np.max(np.abs(Ax - z)) if n else dt(0)
            dua = np.max(np.abs(Hx + g + Aty)) if n else dt(0)
            eps_p = eps_abs + eps_rel * max(_inf(Ax), _inf(z))
            eps_d = eps_abs + eps_rel * max(_inf(Hx), _inf(Aty), _inf(g))
            if fixed_iters is None and pri <= eps_p and dua <= eps_d:
                status = 1
                break
            if adapt and n:
                pr_n = pri / (max(_inf(Ax), _inf(z)) + 1e-10)
                du_n = dua / (max(_inf(Hx), _inf(Aty), _inf(g)) + 1e-10)
                rn = min(max(float(rho) * np.sqrt(pr_n / (du_n + 1e-10)), 1e-6), 1e6)
                if rn > float(rho) * adaptive_tolerance or rn * adaptive_tolerance < float(rho):
                    rho = dt(rn)
                    L = factor(rho)
                    rho_updates += 1
    if fixed_iters is not None or status == 0:
        Ax, Hx, Aty = A @ x, H @ x, A.T @ y
        pri = np.max(np.abs(Ax - z)) if n else dt(0)
        dua = np.max(np.abs(Hx + g + Aty)) if n else dt(0)
    return dict(x=x, y=y, z=z, iters=it, pri_res=float(pri), dua_res=float(dua), status=status,
                rho=float(rho), rho_updates=rho_updates)


def _inf(v):
    return float(np.max(np.abs(v))) if v.size else 0.0


def _chol32(K):
    """Cholesky carried out in fp32 arithmetic (column by column) - mimics the kernel."""
    n = K.shape[0]
    L = np.zeros_like(K)
    for j in range(n):
        d = K[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(d)
        if j + 1 < n:
            L[j + 1:, j] = (K[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def _chol_solve(L, b):
    import scipy.linalg as sla
    yv = sla.solve_triangular(L, b, lower=True, check_finite=False)
    return sla.solve_triangular(L.T, yv, lower=False, check_finite=False).astype(L.dtype)


def solve_problem(x0, r, stance, x_des, mu, delta, tight=False, r_weight=0.0, **kw):
    """Condense + ADMM for one problem.  Returns dict with U (N,12) (zeros on swing legs),
    X (13,N+1), J, wrench (N,6) and the ADMM info.  ``tight=True`` = oracle_tight
    (eps 1e-9, up to 200k iterations, fp64)."""
    N = r.shape[0]
    H, gvec, Sc, c0, idx = srbd_qp.condensed_qp(x0, r, stance, x_des, delta, r_weight=r_weight)
    if tight:
        kw = dict(dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=200000, check_every=25,
                       adaptive_interval=100), **kw)
    info = admm(H, gvec, mu, **kw)
    U = np.zeros((N, 12))
    for s, (i, l) in enumerate(idx):
        U[i, 3 * l:3 * l + 3] = info["x"][3 * s:3 * s + 3]
    X = c0 + (Sc @ info["x"]).reshape(N + 1, 13).T if len(idx) else c0
    info.update(U=U, X=X, J=srbd_qp.objective(X, x_des) + r_weight * float(np.sum(U * U)),
                wrench=srbd_qp.stage_wrench(U, r), H=H, g=gvec, idx=idx)
    return info
