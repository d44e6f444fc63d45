"""oracle_tight (oracle/tight_ipm.py, an interior-point solve of the reference's QP,
src/mpc.py:64-173) pinned to the PINNED oracle: the sparse OSQP restatement oracle/osqp_ref.c
(which reproduces the reference's logged run, tests/test_oracle_golden.py) run to eps 1e-10 on the
very same problems, and to the known answers of SURVEY.md section 8(c).  The GPU tight-parity
tests (tests/test_gpu_tight_parity.py) use the IPM for every sampled problem and repeat this
cross-check on a subset."""
import numpy as np
import pytest

import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, DT, GAIT_NAMES
from oracle import srbd_qp, tight_ipm as ipm, condensed_admm as ca
from oracle.replay import ReplayMPC, params_from_golden, initial_from_golden

ATOL, RTOL = 1e-2, 1e-3          # BASELINE.json north_star: 1e-2 N absolute / 1e-3 relative


def osqp_tight(x0, r, stance, xd, mu, eps=1e-10, max_iter=400000):
    """The pinned oracle at a tight tolerance: (X (13,N+1), wrench (N,6), J, iterations)."""
    from oracle import cpu_baseline as cb
    N = r.shape[0]
    o = cb.OSQPRefC(N)
    cb.lib().osqpref_set_tolerances(o.work, eps, eps, max_iter)
    sol, st, its, _ = o.solve(x0, r, (1 - stance).T.astype(float), xd, mu, DT, -9.81)
    o.close()     # (a problem may stop at max_iter just short of eps: what counts is the agreement below)
    X = sol[12 * N:].reshape(N + 1, 13).T
    U = sol[:12 * N].reshape(N, 12)
    return X, srbd_qp.stage_wrench(U, r), srbd_qp.objective(X, xd), its


def assert_same_optimum(a, X, W, J):
    assert abs(J / a["J"] - 1.0) <= 1e-5
    assert np.abs(a["X"] - X).max() <= 5e-6
    assert np.all(np.abs(a["wrench"] - W) <= ATOL + RTOL * np.abs(W)), np.abs(a["wrench"] - W).max()


@pytest.mark.parametrize("N,gaits,mu,nb", [(10, ("trot",), (1.0, 1.0), 4),
                                           (10, GAIT_NAMES, (0.3, 1.0), 6),
                                           (30, ("trot",), (1.0, 1.0), 2)])
def test_ipm_equals_pinned_osqp_oracle(N, gaits, mu, nb):
    pb = synthetic_batch(16, N=N, gaits=gaits, seed=33, mu=mu)
    for b in range(nb):
        x0, r, st, xd, m = pb.problem(b)
        a = ipm.solve_problem(x0, r, st, xd, m, DT)
        assert a["gap"] <= 1e-6 and a["iters"] < 60
        X, W, J, _ = osqp_tight(x0, r, st, xd, m)
        assert_same_optimum(a, X, W, J)
        # constraints of src/mpc.py:138-173 hold exactly for the interior-point solution
        F = a["U"].reshape(N, 4, 3)
        s = st.astype(bool)
        assert np.all(F[~s] == 0)
        assert np.all(F[s][:, 2] >= 3 - 1e-7) and np.all(F[s][:, 2] <= 100 + 1e-7)
        assert np.all(np.abs(F[s][:, :2]) <= m * F[s][:, 2:3] + 1e-7)


def test_ipm_equals_tight_condensed_admm():
    pb = synthetic_batch(8, N=10, gaits=GAIT_NAMES, seed=5, mu=(0.3, 1.0))
    for b in range(pb.B):
        x0, r, st, xd, m = pb.problem(b)
        a = ipm.solve_problem(x0, r, st, xd, m, DT)
        t = ca.solve_problem(x0, r, st, xd, m, DT, tight=True, rho=0.3, adaptive_interval=25,
                             adaptive_tolerance=3.0, rho_lim=(0.05, 300.0))
        assert t["status"] == 1
        assert abs(a["J"] / t["J"] - 1) < 1e-8 and np.abs(a["X"] - t["X"]).max() < 1e-6
        assert np.all(np.abs(a["wrench"] - t["wrench"]) <= 1e-3)


def test_ipm_flight_and_unique_force_weight():
    """Pronk flight (no stance leg in the whole horizon) and the strictly convex variant."""
    pb = synthetic_batch(4, N=10, gaits=("pronk",), seed=1)
    x0, r, st, xd, m = pb.problem(0)
    a = ipm.solve_problem(x0, r, np.zeros_like(st), xd, m, DT)
    assert np.all(a["U"] == 0) and a["iters"] == 0
    rw = 1e-2
    a = ipm.solve_problem(x0, r, np.ones_like(st), xd, m, DT, r_weight=rw)
    t = ca.solve_problem(x0, r, np.ones_like(st), xd, m, DT, tight=True, rho=0.3, r_weight=rw)
    assert np.abs(a["U"] - t["U"]).max() < 1e-4            # unique optimum: full forces agree


@pytest.mark.parametrize("tick,J_star,f0", [(0, 20484.3999, (118.3214, -0.0017, 127.5505)),
                                            (80, 17087.7415, (84.9185, -0.316, 84.9185))])
def test_known_answers_of_the_golden_run(gold, tick, J_star, f0):
    """exact optimum of the reference's own problems (N=60) at the two
    ticks whose predictions the reference logged."""
    p = params_from_golden(gold)
    mpc = ReplayMPC(initial_from_golden(gold), p)
    mpc.com_pos_start = gold["desired"][tick][3:6].copy()      # reference accumulators of that tick
    mpc.yaw_start = float(gold["desired"][tick][2])
    x0, r, stance, xd, v, om = mpc.tick_problem(tick, gold["state"][tick], gold["feet"][tick])
    a = ipm.solve_problem(x0, r, stance, xd, p["µ"], 0.01)
    assert abs(a["J"] - J_star) < 0.05
    assert np.allclose(a["U"][0].reshape(4, 3).sum(0), f0, atol=2e-3)
