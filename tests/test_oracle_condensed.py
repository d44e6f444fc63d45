"""Internal consistency of the oracle's three formulations of the same QP
(reference src/mpc.py:64-173): sparse CasADi form, dense condensed form, wrench-space
factorisation (the form the CUDA kernels use)."""
import numpy as np

import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, DT
from oracle import condensed_admm as ca, srbd_qp, wrench_form as wf
from oracle.osqp_ref import OSQPRef
from oracle.replay import ReplayMPC, params_from_golden, initial_from_golden


def test_wrench_factorisation_equals_condensed_hessian():
    pb = synthetic_batch(6, N=10, gaits=pkg.problems.GAIT_NAMES, seed=4)
    M = wf.M_full(10, DT)
    for b in range(pb.B):
        x0, r, st, xd, mu = pb.problem(b)
        H, g, Sc, c0, idx = srbd_qp.condensed_qp(x0, r, st, xd, DT)
        G = wf.G_matrix(x0, r, st)
        assert np.abs(H - G.T @ M @ G).max() <= 1e-11 * max(np.abs(H).max(), 1.0)
        # closed-form free response vs the recursion
        k = np.arange(11)
        Rz = srbd_qp.rot_z(x0[2])
        assert np.allclose(c0[0:3], x0[0:3, None] + (Rz @ x0[6:9])[:, None] * k * DT, atol=1e-14)
        assert np.allclose(c0[5], x0[5] + k * DT * x0[11] + k * (k - 1) / 2 * DT ** 2 * x0[12])


def test_condensed_optimum_equals_sparse_optimum():
    """Tight condensed ADMM and the sparse OSQP restatement run to 1e-8 find the same unique
    quantities (objective, states)."""
    pb = synthetic_batch(3, N=5, seed=8)
    for b in range(pb.B):
        x0, r, st, xd, mu = pb.problem(b)
        tight = ca.solve_problem(x0, r, st, xd, mu, DT, tight=True)
        qp, Pd, q, A, l, u = srbd_qp.build_sparse_qp(x0, r, (1 - st).T.astype(float), xd, mu, DT, -9.81)
        solver = OSQPRef(eps_abs=1e-9, eps_rel=1e-9, max_iter=100000)
        sol, status = solver.solve(Pd, q, A, l, u)
        assert status == "solved"
        X = sol[60:].reshape(6, 13).T
        J = srbd_qp.objective(X, xd)
        assert abs(J / tight["J"] - 1) < 1e-6
        assert np.abs(X - tight["X"]).max() < 1e-5


def test_known_answer_tick0_of_the_golden_run(gold):
    """exact optimum of the golden run's tick 0 (N=60):
    J* = 20484.3999, stage-0 net force (118.3214, -0.0017, 127.5505) N."""
    p = params_from_golden(gold)
    mpc = ReplayMPC(initial_from_golden(gold), p)
    x0, r, stance, xd, v, om = mpc.tick_problem(0, gold["state"][0], gold["feet"][0])
    tight = ca.solve_problem(x0, r, stance, xd, 1.0, 0.01, tight=True, eps_abs=1e-7, eps_rel=1e-7,
                             max_iter=20000)
    assert tight["status"] == 1
    assert abs(tight["J"] - 20484.3999) < 0.05
    f0 = tight["U"][0].reshape(4, 3).sum(0)
    assert np.allclose(f0, [118.3214, -0.0017, 127.5505], atol=2e-3)
