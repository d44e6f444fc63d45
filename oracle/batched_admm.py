"""ORACLE (test infrastructure, not product code).

Vectorised-over-problems fp64 version of :func:`oracle.condensed_admm.admm` (same
OSQP-style ADMM with frustum projection on the condensed QP, reference src/mpc.py:64-173) for studying
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


def admm_batch(H, g, mu, keep, rho=0.3, sigma=1e-6, alpha=1.6, eps=1e-3, max_iter=1000,
               check_every=5, f_min=3.0, f_max=100.0, adaptive=None, tol=2.0, rho_lim=(0.03, 30.0)):
    """Batched :func:`oracle.condensed_admm.admm` (projection form).  keep (B,4N) bool stance
    per leg-stage; rho scalar or (B,).  Returns dict(x, y, iters, done, rho, nup)."""
    from .condensed_admm import project_frustum
    B, n = g.shape
    L = n // 3
    rho = np.broadcast_to(np.asarray(rho, dtype=float), (B,)).copy()
    k3 = np.repeat(keep, 3, axis=1)
    mu3 = np.asarray(mu, dtype=float)[:, None]

    def factor(idx):
        Kinv = np.zeros((len(idx), n, n))
        for ii, b in enumerate(idx):
            K = H[b] + np.eye(n) * (sigma + rho[b])
            sel = np.where(k3[b])[0]
            if len(sel):
                Kinv[ii][np.ix_(sel, sel)] = np.linalg.inv(K[np.ix_(sel, sel)])
        return Kinv
    proj = lambda v: project_frustum(v.reshape(B, L, 3), mu3, f_min, f_max).reshape(B, n) * k3
    Kinv = factor(np.arange(B))
    x = np.zeros((B, n))
    y = np.zeros((B, n))
    z = proj(x)
    iters = np.full(B, max_iter)
    done = np.zeros(B, bool)
    nup = np.zeros(B, int)
    for it in range(max_iter + 1):
        if it % check_every == 0:
            Hx = np.einsum('bij,bj->bi', H, x)
            pri = np.abs(x - z).max(1)
            dua = np.abs(Hx + g + y).max(1)
            nA = np.maximum(np.abs(x).max(1), np.abs(z).max(1))
            nD = np.maximum(np.maximum(np.abs(Hx).max(1), np.abs(y).max(1)), np.abs(g).max(1))
            ok = (pri <= eps + eps * nA) & (dua <= eps + eps * nD) & ~done
            iters[ok] = it
            done |= ok
            if done.all():
                break
            if adaptive and it > 0 and it % adaptive == 0:
                rn = rho * np.sqrt((pri / (nA + 1e-10)) / (dua / (nD + 1e-10) + 1e-10))
                rn = np.clip(rn, rho_lim[0], rho_lim[1])
                upd = ((rn > tol * rho) | (rn * tol < rho)) & ~done
                if upd.any():
                    rho[upd] = rn[upd]
                    nup[upd] += 1
                    Kinv[upd] = factor(np.where(upd)[0])
        rhs = (sigma * x - g + rho[:, None] * z - y) * k3
        xt = np.einsum('bij,bj->bi', Kinv, rhs)
        xn = alpha * xt + (1 - alpha) * x
        zh = alpha * xt + (1 - alpha) * z
        zn = proj(zh + y / rho[:, None])
        yn = (y + rho[:, None] * (zh - zn)) * k3
        act = ~done
        x[act], z[act], y[act] = xn[act], zn[act], yn[act]
    return dict(x=x, y=y, iters=iters, done=done, rho=rho, nup=nup)
