"""The oracle against the reference's own golden vectors: the committed 1000-tick run
(reference src/simulation_log.pkl -> tests/golden/simulation_log_golden.npz, made by
scripts/make_golden_from_pkl.py).  Gate of SURVEY.md section 8(c): forces <= 1e-6 N,
predicted states <= 1e-8, desired states bit-exact."""
import numpy as np
import pytest

from oracle.replay import ReplayMPC, params_from_golden, initial_from_golden


def test_python_oracle_replays_first_100_ticks(gold):
    """numpy/scipy restatement (oracle/osqp_ref.py): ticks 0..99 incl. both logged horizon
    predictions (t=0, t=80), the 150-iteration first tick and its single rho update."""
    mpc = ReplayMPC(initial_from_golden(gold), params_from_golden(gold))
    for t in range(100):
        u0 = mpc.solve(t, gold["state"][t], gold["feet"][t])
        assert np.abs(u0 - gold["forces"][t]).max() <= 1e-6
        assert np.array_equal(mpc.last["x_des"][:12, 0], gold["desired"][t])
        if t == 0:
            assert mpc.last["iters"] == 150
            assert mpc.last["rho"] == pytest.approx(0.0017029794584659777, rel=1e-12)
        if t in (0, 80):
            k = list(gold["pred_t"]).index(t)
            assert np.abs(mpc.last["X"][:12] - gold["pred_state"][k]).max() <= 1e-8
            assert np.abs(mpc.last["U"][2::3] - gold["pred_fz"][k]).max() <= 1e-6
            assert np.array_equal(mpc.last["x_des"][:12], gold["pred_desired"][k])


def test_c_oracle_replays_all_1000_ticks(gold):
    """C restatement (oracle/osqp_ref.c, the CPU baseline of bench.py): every logged force of
    the reference's run, with the iteration histogram of SURVEY.md section 6."""
    from oracle.cpu_baseline import OSQPRefC
    p = params_from_golden(gold)
    mpc = ReplayMPC(initial_from_golden(gold), p)
    c = OSQPRefC(p["N"])
    warm, hist, worst = None, {}, 0.0
    for t in range(1000):
        x0, r, stance, xd, v, om = mpc.tick_problem(t, gold["state"][t], gold["feet"][t])
        sol, st, it, rho = c.solve(x0, r, (1 - stance).T.astype(float), xd, p["µ"], 0.01, p["g"], warm)
        assert st == 1
        mpc.com_pos_start = mpc.com_pos_start + v * 0.01
        mpc.yaw_start = mpc.yaw_start + om * 0.01
        warm = sol
        worst = max(worst, np.abs(sol[:12] - gold["forces"][t]).max())
        hist[it] = hist.get(it, 0) + 1
        assert np.array_equal(xd[:12, 0], gold["desired"][t])
    assert worst <= 1e-6, worst
    assert hist == {25: 848, 50: 117, 75: 23, 100: 6, 125: 5, 150: 1}
    assert rho == pytest.approx(0.0017029794584659777, rel=1e-10)


def test_c_and_python_oracle_agree_on_synthetic_batch():
    import mpc_b200 as pkg
    from oracle import cpu_baseline, srbd_qp
    from oracle.osqp_ref import OSQPRef
    pb = pkg.problems.synthetic_batch(6, N=10, gaits=pkg.problems.GAIT_NAMES, seed=2, mu=(0.4, 1.0))
    out = cpu_baseline.solve_batch(pb, threads=2)
    assert np.all(out["status"] == 1)
    for b in range(pb.B):
        x0, r, stance, xd, mu = pb.problem(b)
        qp, Pd, q, A, l, u = srbd_qp.build_sparse_qp(x0, r, (1 - stance).T.astype(float), xd, mu,
                                                     0.01, -9.81)
        solver = OSQPRef()
        sol, status = solver.solve(Pd, q, A, l, u)
        assert status == "solved" and solver.info["iters"] == out["iters"][b]
        assert np.abs(sol[:120].reshape(10, 12) - out["U"][b]).max() <= 1e-7
