"""ORACLE (test infrastructure, not product code).

``oracle_tight`` of SURVEY.md section 7-1b: an exact fp64 solve of the QP that the
reference's ``MPC.__init__`` declares (reference ``src/mpc.py:64-173``), used by the parity
tests as "the tight-tolerance solve" for EVERY problem of a sample (an ADMM needs a
problem-dependent number of iterations to get there; an interior-point method does not).

The problem is the condensed form of :func:`oracle.srbd_qp.condensed_qp` (which follows
``src/mpc.py:64-136`` by the plain recursion) with the reference's inequality rows on the
stance forces, one-sided as the reference writes them (``src/mpc.py:152-173``; the duplicate
friction rows are dropped, they do not change the feasible set):

    min 1/2 u'Hu + g'u   s.t.   f_min <= fz <= f_max,  +-fx - mu fz <= 0,  +-fy - mu fz <= 0

solved by a Mehrotra predictor-corrector primal-dual interior-point method.  H is only
positive SEMI-definite (zero force weight, ``src/mpc.py:121``): the force split over the legs
is not unique, but X, the objective and the per-stage net wrench are, and those are what the
tests compare.  Pinned to the sparse OSQP restatement (``oracle/osqp_ref.c``, itself pinned on
the reference's logged run) by ``tests/test_oracle_tight.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import srbd_qp


def leg_rows(mu, f_min=srbd_qp.F_MIN, f_max=srbd_qp.F_MAX):
    """C (6,3), d (6,) of one stance leg:  C f <= d."""
    C = np.array([[0.0, 0.0, -1.0],
                  [0.0, 0.0, 1.0],
                  [1.0, 0.0, -mu],
                  [-1.0, 0.0, -mu],
                  [0.0, 1.0, -mu],
                  [0.0, -1.0, -mu]])
    d = np.array([-f_min, f_max, 0.0, 0.0, 0.0, 0.0])
    return C, d


def solve_qp(H, g, mu, f_min=srbd_qp.F_MIN, f_max=srbd_qp.F_MAX, tol=1e-11, max_iter=100):
    """Returns dict(x, lam, iters, gap, res).  n = 3 * (number of stance leg-stages)."""
    n = H.shape[0]
    if n == 0:
        return dict(x=np.zeros(0), lam=np.zeros(0), iters=0, gap=0.0, res=0.0)
    nl = n // 3
    C1, d1 = leg_rows(mu, f_min, f_max)
    d = np.tile(d1, nl)

    def Cmul(x):                     # (6 nl,)
        return (x.reshape(nl, 3) @ C1.T).reshape(-1)

    def Ctmul(v):                    # (n,)
        return (v.reshape(nl, 6) @ C1).reshape(-1)

    # strictly feasible start: fz at the middle of its range, no tangential force
    x = np.tile(np.array([0.0, 0.0, 0.5 * (f_min + f_max)]), nl)
    s = d - Cmul(x)
    scale = max(1.0, float(np.abs(g).max()), float(np.abs(H).max()))
    lam = np.full(6 * nl, 1.0) * max(1.0, scale * 1e-3)
    it = 0
    for it in range(1, max_iter + 1):
        rd = H @ x + g + Ctmul(lam)                  # dual residual
        rp = Cmul(x) + s - d                         # primal residual (0 after a full step)
        mu_c = float(s @ lam) / (6 * nl)
        if max(np.abs(rd).max() / scale, np.abs(rp).max(), mu_c / scale) < tol:
            break
        w = lam / s
        # H + C' W C : the second term is block diagonal (3x3 per leg)
        blocks = np.einsum("ki,lk,kj->lij", C1, w.reshape(nl, 6), C1)
        Kmat = H.copy()
        for l in range(nl):
            Kmat[3 * l:3 * l + 3, 3 * l:3 * l + 3] += blocks[l]
        # H is singular and W -> 0 on inactive rows: a tiny proximal term keeps the Newton matrix
        # numerically positive definite (it perturbs the step, not the residuals, so the
        # fixed point is unchanged)
        reg = 1e-13 * scale
        while True:
            try:
                cf = sla.cho_factor(Kmat + reg * np.eye(n), lower=True, check_finite=False)
                break
            except np.linalg.LinAlgError:
                reg *= 100.0

        def newton(rc):
            # rows: H dx + C' dlam = -rd ; C dx + ds = -rp ; S dlam + L ds = -rc
            rhs = -rd - Ctmul((rc - lam * rp) / s * -1.0)
            dx = sla.cho_solve(cf, rhs, check_finite=False)
            ds = -rp - Cmul(dx)
            dl = -(rc + lam * ds) / s
            return dx, ds, dl

        dx, ds, dl = newton(s * lam)                 # affine (predictor)
        a_p = _step(s, ds)
        a_d = _step(lam, dl)
        mu_aff = float((s + a_p * ds) @ (lam + a_d * dl)) / (6 * nl)
        sig = (mu_aff / mu_c) ** 3 if mu_c > 0 else 0.0
        dx, ds, dl = newton(s * lam + ds * dl - sig * mu_c)   # corrector
        a_p = min(1.0, 0.995 * _step(s, ds))
        a_d = min(1.0, 0.995 * _step(lam, dl))
        x = x + a_p * dx
        s = s + a_p * ds
        lam = lam + a_d * dl
    return dict(x=x, lam=lam, iters=it, gap=float(s @ lam) / (6 * nl),
                res=float(np.abs(H @ x + g + Ctmul(lam)).max()))


def _step(v, dv):
    neg = dv < 0
    return 1.0 if not neg.any() else min(1.0, float(np.min(-v[neg] / dv[neg])))


def solve_problem(x0, r, stance, x_des, mu, delta, r_weight=0.0, **kw):
    """Condense + IPM for one problem: dict with U (N,12), X (13,N+1), J, wrench (N,6), H, g, idx."""
    N = r.shape[0]
    H, gvec, Sc, c0, idx = srbd_qp.condensed_qp(x0, r, stance, x_des, delta, r_weight=r_weight)
    info = solve_qp(H, gvec, mu, **kw)
    U = np.zeros((N, 12))
    for s_, (i, l) in enumerate(idx):
        U[i, 3 * l:3 * l + 3] = info["x"][3 * s_:3 * s_ + 3]
    X = c0 + (Sc @ info["x"]).reshape(N + 1, 13).T if len(idx) else c0
    info.update(U=U, X=X, J=srbd_qp.objective(X, x_des) + r_weight * float(np.sum(U * U)),
                wrench=srbd_qp.stage_wrench(U, r), H=H, g=gvec, idx=idx)
    return info
