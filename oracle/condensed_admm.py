"""ORACLE (test infrastructure, not product code).

``oracle_admm64`` / ``oracle_tight`` of SURVEY.md section 7: the algorithm of the CUDA
path restated in numpy so it can run in fp64 at any tolerance - OSQP's ADMM (Stellato et
al. 2020, Algorithm 1: x-update with K = H + (sigma+rho) I, relaxation alpha, dual update)
on the condensed QP over the stance forces, with the constraint
copy z kept in the feasible set of reference ``src/mpc.py:148-173`` by an exact projection:

    minimise 1/2 u'Hu + g'u   s.t.  u = z,  z_leg in C = { f : f_min <= fz <= f_max,
                                                          |fx| <= mu fz, |fy| <= mu fz }

(the same feasible set as the reference's 2 + 8 rows per leg; swing legs are eliminated).
The QP itself comes from :func:`oracle.srbd_qp.condensed_qp`, which follows reference
``src/mpc.py:64-136``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.
"""
from __future__ import annotations

import numpy as np

from . import srbd_qp


def project_frustum(w, mu, f_min=srbd_qp.F_MIN, f_max=srbd_qp.F_MAX):
    """Exact Euclidean projection of w (...,3) onto C.  For fixed fz the optimal fx, fy are
    clamps, which leaves a 1-D convex piecewise-quadratic in fz with three linear-derivative
    pieces; its root is then clamped to [f_min, f_max]."""
    w = np.asarray(w)
    ax, ay, wz = np.abs(w[..., 0]), np.abs(w[..., 1]), w[..., 2]
    big, small = np.maximum(ax, ay), np.minimum(ax, ay)
    f2 = (wz + mu * big) / (1 + mu * mu)
    f1 = (wz + mu * (ax + ay)) / (1 + 2 * mu * mu)
    fz = np.where(mu * wz >= big, wz, np.where(mu * f2 >= small, f2, f1))
    fz = np.clip(fz, f_min, f_max)
    out = np.empty_like(w)
    out[..., 0] = np.clip(w[..., 0], -mu * fz, mu * fz)
    out[..., 1] = np.clip(w[..., 1], -mu * fz, mu * fz)
    out[..., 2] = fz
    return out


def admm(H, g, mu, rho=0.3, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
         max_iter=1000, check_every=5, x=None, y=None, f_min=srbd_qp.F_MIN,
         f_max=srbd_qp.F_MAX, dtype=np.float64, fixed_iters=None, adaptive_interval=0,
         adaptive_tolerance=2.0, rho_lim=(0.03, 30.0), adaptive_floor=0.0):
    """Returns dict(x, y, z, iters, pri_res, dua_res, status, rho, rho_updates).
    ``fixed_iters`` runs exactly that many iterations (iterate-level parity with the CUDA
    kernel).  ``adaptive_interval`` > 0 applies OSQP's rho adaptation rule every that many
    iterations (refactorising K), as the CUDA kernel does.  Residuals are OSQP's with A = I:
    r_p = |x - z|_inf, r_d = |Hx + g + y|_inf.  ``adaptive_floor``: rho stops adapting once both
    normalised residuals are below it (the fp32 kernel uses 1e-6, its rounding-noise level; only
    reachable with eps < 1e-5)."""
    n = H.shape[0]
    dt = dtype
    H = H.astype(dt)
    g = g.astype(dt)
    x = np.zeros(n, dt) if x is None else x.astype(dt)
    y = np.zeros(n, dt) if y is None else y.astype(dt)
    proj = lambda v: project_frustum(v.reshape(-1, 3), dt(mu), dt(f_min), dt(f_max)).reshape(-1).astype(dt)
    z = proj(x)
    rho, sigma, alpha = dt(rho), dt(sigma), dt(alpha)

    def factor(rho_):
        K = H + (sigma + rho_) * np.eye(n, dtype=dt)
        return np.linalg.cholesky(K.astype(np.float64)).astype(dt) if dt == np.float64 \
            else _chol32(K)
    if n:
        L = factor(rho)
    rho_updates = 0
    status, it = 0, 0
    pri = dua = dt(0)
    total = fixed_iters if fixed_iters is not None else max_iter

    def residuals():
        Hx = H @ x
        pri_ = _inf(x - z)
        dua_ = _inf(Hx + g + y)
        return pri_, dua_, max(_inf(x), _inf(z)), max(_inf(Hx), _inf(y), _inf(g))

    if fixed_iters is None and n == 0:
        return dict(x=x, y=y, z=z, iters=0, pri_res=0.0, dua_res=0.0, status=1, rho=float(rho),
                    rho_updates=0)
    if fixed_iters is None:              # the kernel tests the initial iterate too (it = 0)
        pri, dua, nA, nD = residuals()
        if pri <= eps_abs + eps_rel * nA and dua <= eps_abs + eps_rel * nD:
            return dict(x=x, y=y, z=z, iters=0, pri_res=pri, dua_res=dua, status=1,
                        rho=float(rho), rho_updates=0)
    for it in range(1, total + 1):
        rhs = sigma * x - g + rho * z - y
        xt = _chol_solve(L, rhs) if n else rhs
        x = alpha * xt + (1 - alpha) * x
        zh = alpha * xt + (1 - alpha) * z
        zn = proj(zh + y / rho)
        y = y + rho * (zh - zn)
        z = zn
        adapt = adaptive_interval > 0 and it % adaptive_interval == 0
        if (fixed_iters is None and it % check_every == 0) or adapt:
            pri, dua, nA, nD = residuals()
            if fixed_iters is None and pri <= eps_abs + eps_rel * nA and dua <= eps_abs + eps_rel * nD:
                status = 1
                break
            if adapt and n and max(pri / (nA + 1e-10), dua / (nD + 1e-10)) > adaptive_floor:
                rn = float(rho) * np.sqrt((pri / (nA + 1e-10)) / (dua / (nD + 1e-10) + 1e-10))
                rn = min(max(rn, rho_lim[0]), rho_lim[1])
                if rn > float(rho) * adaptive_tolerance or rn * adaptive_tolerance < float(rho):
                    rho = dt(rn)
                    L = factor(rho)
                    rho_updates += 1
    if fixed_iters is not None or status == 0:
        pri, dua, _, _ = residuals()
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
                       adaptive_interval=100, rho_lim=(1e-4, 1e4)), **kw)
    info = admm(H, gvec, mu, **kw)
    U = np.zeros((N, 12))
    for s, (i, l) in enumerate(idx):
        U[i, 3 * l:3 * l + 3] = info["x"][3 * s:3 * s + 3]
    X = c0 + (Sc @ info["x"]).reshape(N + 1, 13).T if len(idx) else c0
    info.update(U=U, X=X, J=srbd_qp.objective(X, x_des) + r_weight * float(np.sum(U * U)),
                wrench=srbd_qp.stage_wrench(U, r), H=H, g=gvec, idx=idx)
    return info
