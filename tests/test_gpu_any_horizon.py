"""Any horizon (reference src/main.py:41 accepts any params['N']): horizons without their own
compiled kernel run padded on the next compiled one (all-swing, zero-cost tail stages) and must give
the N-stage problem's own iterates and optimum."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mpc_b200 as pkg                                               # noqa: E402
from mpc_b200.problems import synthetic_batch, DT, GAIT_NAMES         # noqa: E402
from oracle import condensed_admm as ca                               # noqa: E402
from test_gpu_parity import gpu_solve, close, ATOL, RTOL              # noqa: E402
from test_gpu_tight_parity import check_sample                        # noqa: E402


@pytest.mark.parametrize("N", [1, 3, 7, 9, 13, 25, 33, 47])
def test_iterate_parity_on_padded_horizons(N):
    """Same ADMM, same iteration count as the fp64 oracle on the N-stage problem."""
    assert pkg._capi.kernel_horizon(N) > N
    B, K = 6, 50
    pb = synthetic_batch(B, N=N, gaits=GAIT_NAMES, seed=N)
    out = gpu_solve(pb, max_iter=K, check_every=100000, eps_abs=0.0, eps_rel=0.0, warm_mode=0,
                    adaptive_rho_interval=0)
    assert out["U"].shape == (B, N, 12) and out["X"].shape == (B, N + 1, 13)
    assert np.all(out["iters"] == K)
    for b in range(B):
        x0, r, stance, xd, mu = pb.problem(b)
        ref = ca.solve_problem(x0, r, stance, xd, mu, DT, fixed_iters=K, rho=float(out["mpc"].cfg.rho))
        err = np.abs(out["U"][b] - ref["U"])
        assert np.all(err <= 2 * ATOL + 3e-3 * np.abs(ref["U"]).max() + RTOL * np.abs(ref["U"])), (b, err.max())
        assert close(out["X"][b].T, ref["X"], atol=2e-4, rtol=1e-3)


@pytest.mark.parametrize("N", [7, 13, 25])
def test_tight_parity_on_padded_horizons(N):
    pb = synthetic_batch(32, N=N, gaits=GAIT_NAMES, seed=100 + N, mu=(0.3, 1.0))
    check_sample(pb, np.arange(10), K=40000, n_osqp=1)


def test_padded_horizon_warm_start_and_host_path():
    """Warm-start round trip ([B,N,12] views of the padded state) and the host path on N = 7."""
    N, B = 7, 40
    pb = synthetic_batch(B, N=N, seed=3)
    dev = torch.device("cuda", 0)
    args = [torch.from_numpy(a).to(dev) for a in pb.f32()]
    mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=1)
    U1, X1, s1 = mpc.solve(*args)
    xw, yw = mpc.get_warm(B)
    assert xw.shape == (B, N, 12) and torch.equal(xw, U1)
    U2, _, s2 = mpc.solve(*args)
    assert s2.iters.float().mean() < s1.iters.float().mean()
    m2 = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=1)
    m2.set_warm(U1)
    U3, _, s3 = m2.solve(*args)
    assert torch.equal(U3, U2) and torch.equal(s3.iters, s2.iters)
    m3 = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0)
    Uh, Xh, sh = m3.solve_host(*pb.f32())
    mpc.reset_warm()
    U4, X4, s4 = mpc.solve(*args)
    assert np.array_equal(Uh, U4.cpu().numpy()) and np.array_equal(Xh, X4.cpu().numpy())
    # stream-ordered masked reset: only the masked slots go cold again
    m = torch.zeros(B, dtype=torch.uint8, device=dev)
    m[::2] = 1
    mpc.reset_warm_async(B, 0, m)
    _, _, s5 = mpc.solve(*args)
    it5, it1, it2 = s5.iters.cpu().numpy(), s1.iters.cpu().numpy(), mpc.solve(*args)[2].iters.cpu().numpy()
    assert np.array_equal(it5[::2], it1[::2])


def test_dropin_mpc_with_an_uncompiled_horizon(gold):
    from oracle.replay import params_from_golden, initial_from_golden
    from test_gpu_parity import _FakeLite3, _Logger
    params = params_from_golden(gold, N=14)
    initial = initial_from_golden(gold)
    gp = pkg.GaitPlan.from_initial(initial, params)
    lite3, logger = _FakeLite3(gold), _Logger()
    mpc = pkg.MPC(lite3=lite3, initial=initial, footstep_planner=gp, params=params)
    for t in range(5):
        lite3.t = t
        f = mpc.solve(t, logger)
        assert mpc.x_log.shape == (12, 15) and mpc.u_plot.shape == (12, 14)
        assert 40.0 < sum(f[leg][2] for leg in pkg.LEGS) < 200.0
