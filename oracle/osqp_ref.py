"""ORACLE (test infrastructure, not product code).

fp64 restatement of what CasADi's OSQP conic plugin does with the reference's QP at
every ``self.opt.solve()`` (reference ``src/mpc.py:49-55, 242-258``).  CasADi and OSQP
are un-vendored, un-pinned third-party dependencies of the reference (README.md:138)
and are absent from this image; the algorithm below is OSQP 0.6-series' published
ADMM (Stellato et al., "OSQP: an operator splitting solver for quadratic programs",
2020) with the default settings and the CasADi call sequence listed in SURVEY.md
Appendix A.  It is pinned by ``tests/golden/simulation_log_golden.npz`` (all 1000x12
logged forces of the reference's committed run, see tests/test_oracle_golden.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_TOL = 1e-4
RHO_EQ_OVER_RHO_INEQ = 1e3
OSQP_INFTY = 1e30


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


class OSQPRef:
    """Persistent OSQP workspace as CasADi keeps it across ticks: rho survives,
    scaling/factorisation are redone at every solve (osqp_update_P_A)."""

    def __init__(self, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
                 max_iter=1000, scaling=10, check_termination=25,
                 adaptive_rho=True, adaptive_rho_interval=100,
                 adaptive_rho_tolerance=5.0):
        self.rho = rho
        self.sigma = sigma
        self.alpha = alpha
        self.eps_abs = eps_abs
        self.eps_rel = eps_rel
        self.max_iter = max_iter
        self.scaling = scaling
        self.check_termination = check_termination
        self.adaptive_rho = adaptive_rho
        self.adaptive_rho_interval = adaptive_rho_interval
        self.adaptive_rho_tolerance = adaptive_rho_tolerance
        self.info = {}

    # -- Ruiz equilibration (Appendix A-3) --------------------------------
    def _scale(self, P_diag, q, A, l, u):
        n, m = P_diag.size, A.shape[0]
        D = np.ones(n)
        E = np.ones(m)
        c = 1.0
        Pd = P_diag.copy()          # P is diagonal for this QP
        q = q.copy()
        A = A.tocsc(copy=True)
        for _ in range(self.scaling):
            absA = abs(A)
            colA = np.asarray(absA.max(axis=0).todense()).ravel()
            rowA = np.asarray(absA.max(axis=1).todense()).ravel()
            Dt = np.maximum(np.abs(Pd), colA)
            Et = rowA
            Dt = 1.0 / np.sqrt(_limit(Dt))
            Et = 1.0 / np.sqrt(_limit(Et))
            Pd = Dt * Pd * Dt
            A = sp.diags(Et) @ A @ sp.diags(Dt)
            q = Dt * q
            D *= Dt
            E *= Et
            # cost normalisation
            c_t = float(np.mean(np.abs(Pd)))
            qn = float(_limit(np.array([np.max(np.abs(q))]))[0])
            c_t = max(c_t, qn)
            c_t = float(_limit(np.array([c_t]))[0])
            c_t = 1.0 / c_t
            Pd = Pd * c_t
            q = q * c_t
            c *= c_t
        ls = np.where(np.isfinite(l), E * l, l)
        us = np.where(np.isfinite(u), E * u, u)
        return Pd, q, A.tocsc(), ls, us, D, E, c

    def _rho_vec(self, l, u):
        rv = np.full(l.size, self.rho)
        loose = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
        eq = (u - l) < RHO_TOL
        rv[loose] = RHO_MIN
        rv[eq & ~loose] = RHO_EQ_OVER_RHO_INEQ * self.rho
        return rv

    def _factor(self, Pd, A, rho_vec):
        n, m = Pd.size, A.shape[0]
        K = sp.bmat([[sp.diags(Pd + self.sigma), A.T],
                     [A, sp.diags(-1.0 / rho_vec)]], format="csc")
        return spla.splu(K)

    def solve(self, P_diag, q, A, l, u, x_warm=None):
        """One CasADi ``solve()``: rescale, refactor, warm start x (y=0, z=Ax), ADMM."""
        n, m = P_diag.size, A.shape[0]
        Pd, qs, As, ls, us, D, E, c = self._scale(P_diag, q, A, l, u)
        rho_vec = self._rho_vec(ls, us)
        lu = self._factor(Pd, As, rho_vec)
        x = np.zeros(n) if x_warm is None else x_warm / D
        z = As @ x
        y = np.zeros(m)
        Dinv, Einv, cinv = 1.0 / D, 1.0 / E, 1.0 / c
        status = "max_iter"
        it = 0
        rho_updates = 0
        for it in range(1, self.max_iter + 1):
            x_prev, z_prev = x, z
            rhs = np.concatenate([self.sigma * x_prev - qs, z_prev - y / rho_vec])
            sol = lu.solve(rhs)
            xt = sol[:n]
            zt = z_prev + (sol[n:] - y) / rho_vec
            x = self.alpha * xt + (1.0 - self.alpha) * x_prev
            zhat = self.alpha * zt + (1.0 - self.alpha) * z_prev
            z = np.clip(zhat + y / rho_vec, ls, us)
            y = y + rho_vec * (zhat - z)
            check = self.check_termination and it % self.check_termination == 0
            adapt = self.adaptive_rho and self.adaptive_rho_interval and \
                it % self.adaptive_rho_interval == 0
            if check or adapt:
                Ax = As @ x
                Px = Pd * x
                Aty = As.T @ y
            if check:
                pri = np.max(np.abs(Einv * (Ax - z)))
                dua = cinv * np.max(np.abs(Dinv * (Px + qs + Aty)))
                eps_p = self.eps_abs + self.eps_rel * max(np.max(np.abs(Einv * Ax)),
                                                           np.max(np.abs(Einv * z)))
                eps_d = self.eps_abs + self.eps_rel * cinv * max(
                    np.max(np.abs(Dinv * Px)), np.max(np.abs(Dinv * Aty)),
                    np.max(np.abs(Dinv * qs)))
                self.info.update(pri_res=pri, dua_res=dua, eps_pri=eps_p, eps_dua=eps_d)
                if pri < eps_p and dua < eps_d:
                    status = "solved"
                    break
            if adapt:
                pr = np.max(np.abs(Ax - z))
                dr = np.max(np.abs(Px + qs + Aty))
                pr /= max(np.max(np.abs(Ax)), np.max(np.abs(z))) + 1e-10
                dr /= max(np.max(np.abs(Px)), np.max(np.abs(Aty)), np.max(np.abs(qs))) + 1e-10
                rho_new = self.rho * np.sqrt(pr / (dr + 1e-10))
                rho_new = min(max(rho_new, RHO_MIN), RHO_MAX)
                if rho_new > self.rho * self.adaptive_rho_tolerance or \
                        rho_new < self.rho / self.adaptive_rho_tolerance:
                    self.rho = float(rho_new)
                    rho_vec = self._rho_vec(ls, us)
                    lu = self._factor(Pd, As, rho_vec)
                    rho_updates += 1
        self.info.update(iters=it, status=status, rho=self.rho, rho_updates=rho_updates)
        return D * x, status
