"""ORACLE (test infrastructure, not product code).

Vectorised-over-problems fp64 version of :func:`oracle.condensed_admm.admm` (same
OSQP-style ADMM on the condensed QP, reference src/mpc.py:64-173) for studying
convergence on whole batches.  All 12N forces are kept; swing legs are pinned to zero by
deleting their rows/columns (K^-1 is zero there).
"""
import numpy as np
from . import srbd_qp, wrench_form as wf


def build(pb, delta=0.01, r_weight=0.0):
    """H (B,12N,12N), g (B,12N) in fp64 through the wrench factorisation (checked against
    the plain recursion in tests/test_oracle_condensed.py)."""
    B, N = pb.B, pb.N
    M = wf.M_full(N, delta)
    H = np.zeros((B, 12 * N, 12 * N))
    g = np.zeros((B, 12 * N))
    for b in range(B):
        x0, r, st, xd, mu = pb.problem(b)
        Gf = wf.G_matrix(x0, r, np.ones_like(st))
        keep = np.repeat(st.reshape(-1).astype(bool), 3)
        Gf[:, ~keep] = 0.0
        H[b] = Gf.T @ M @ Gf + 2 * r_weight * np.diag(keep.astype(float))
        c0 = srbd_qp.free_response(x0, N, delta)
        S = srbd_qp.prediction_matrix(x0, r, delta)
        e0 = (c0 - xd).T.reshape(-1)
        g[b] = 2.0 * S.T @ (np.tile(srbd_qp.W_STATE, N + 1) * e0)
        g[b][~keep] = 0.0
    return H, g


def Aop(x, mu):
    f = x.reshape(x.shape[0], -1, 3)
    m = mu[:, None]
    return np.stack([f[..., 2], f[..., 0] - m * f[..., 2], -f[..., 0] - m * f[..., 2],
                     f[..., 1] - m * f[..., 2], -f[..., 1] - m * f[..., 2]], -1)


def ATop(y, mu):
    m = mu[:, None]
    return np.stack([y[..., 1] - y[..., 2], y[..., 3] - y[..., 4],
                     y[..., 0] - m * y[..., 1:].sum(-1)], -1).reshape(y.shape[0], -1)


def admm_batch(H, g, mu, keep, rho=0.3, sigma=1e-6, alpha=1.6, eps=1e-3, max_iter=1000,
               check_every=5, f_min=3.0, f_max=100.0, adaptive=None, tol=5.0, rho_lim=(1e-6, 1e6)):
    """keep (B,4N) bool stance per leg-stage.  rho may be scalar or (B,).  Returns x, y, iters."""
    B, n = g.shape
    L = n // 3
    rho = np.broadcast_to(np.asarray(rho, dtype=float), (B,)).copy()
    k3 = np.repeat(keep, 3, axis=1)
    lo = np.array([f_min, -np.inf, -np.inf, -np.inf, -np.inf])
    hi = np.array([f_max, 0.0, 0.0, 0.0, 0.0])

    def factor(idx):
        Kinv = np.zeros((len(idx), n, n))
        for ii, b in enumerate(idx):
            d = np.tile(np.array([2.0, 2.0, 1 + 4 * mu[b] ** 2]), L)
            K = H[b] + np.diag(sigma + rho[b] * d)
            sel = np.where(k3[b])[0]
            if len(sel):
                Kinv[ii][np.ix_(sel, sel)] = np.linalg.inv(K[np.ix_(sel, sel)])
        return Kinv
    Kinv = factor(np.arange(B))
    x = np.zeros((B, n)); y = np.zeros((B, L, 5))
    z = np.clip(Aop(x, mu), lo, hi) * keep[..., None]
    iters = np.full(B, max_iter); done = np.zeros(B, bool)
    nup = np.zeros(B, int)
    for it in range(max_iter + 1):
        if it % check_every == 0:
            Ax = Aop(x, mu) * keep[..., None]
            Hx = np.einsum('bij,bj->bi', H, x)
            Aty = ATop(y, mu) * k3
            pri = np.abs(Ax - z).reshape(B, -1).max(1)
            dua = np.abs(Hx + g + Aty).max(1)
            nA = np.maximum(np.abs(Ax).reshape(B, -1).max(1), np.abs(z).reshape(B, -1).max(1))
            nD = np.maximum(np.maximum(np.abs(Hx).max(1), np.abs(Aty).max(1)), np.abs(g).max(1))
            ok = (pri <= eps + eps * nA) & (dua <= eps + eps * nD) & ~done
            iters[ok] = it; done |= ok
            if done.all():
                break
            if adaptive and it > 0 and it % adaptive == 0:
                rn = rho * np.sqrt((pri / (nA + 1e-10)) / (dua / (nD + 1e-10) + 1e-10))
                rn = np.clip(rn, rho_lim[0], rho_lim[1])
                upd = ((rn > tol * rho) | (rn < rho / tol)) & ~done
                if upd.any():
                    rho[upd] = rn[upd]; nup[upd] += 1
                    Kinv[upd] = factor(np.where(upd)[0])
        rhs = sigma * x - g + ATop(rho[:, None, None] * z - y, mu) * k3
        xt = np.einsum('bij,bj->bi', Kinv, rhs)
        zt = Aop(xt, mu) * keep[..., None]
        xn = alpha * xt + (1 - alpha) * x
        zh = alpha * zt + (1 - alpha) * z
        zn = np.clip(zh + y / rho[:, None, None], lo, hi) * keep[..., None]
        yn = (y + rho[:, None, None] * (zh - zn)) * keep[..., None]
        act = ~done
        x[act], z[act], y[act] = xn[act], zn[act], yn[act]
    return dict(x=x, y=y, iters=iters, done=done, rho=rho, nup=nup)
