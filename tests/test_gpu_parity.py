"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI
(ctypes -> libcmpc.so), against the fp64 oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): forces within 1e-3 relative / 1e-2 N absolute,
written as |a-b| <= 1e-2 + 1e-3 |b|.  Because the reference's QP has zero force weight
(src/mpc.py:121) its optimal force split over the legs is not unique,
so full force vectors are compared (a) iterate-for-iterate against the same ADMM in fp64
and (b) against the tight fp64 optimum when a force weight makes the optimum unique;
with the reference's r_weight = 0 the tight comparison is on the unique quantities
(state trajectory, objective, per-stage net wrench)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mpc_b200 as pkg                                   # noqa: E402
from mpc_b200.problems import synthetic_batch, DT, GAIT_NAMES   # noqa: E402
from oracle import condensed_admm as ca, srbd_qp, tight_ipm as ipm          # noqa: E402

ATOL, RTOL = 1e-2, 1e-3


def close(a, b, atol=ATOL, rtol=RTOL):
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


def dev_args(pb, dev="cuda:0"):
    return [torch.from_numpy(a).to(dev) for a in pb.f32()]


def gpu_solve(pb, want_X=True, **opts):
    mpc = pkg.BatchedMPC(N=pb.N, max_batch=pb.B, **opts)
    U, X, st = mpc.solve(*dev_args(pb), want_X=want_X)
    torch.cuda.synchronize()
    out = dict(U=U.cpu().numpy().astype(np.float64),
               X=None if X is None else X.cpu().numpy().astype(np.float64),
               iters=st.iters.cpu().numpy(), pri=st.pri_res.cpu().numpy(),
               dua=st.dua_res.cpu().numpy(), status=st.status.cpu().numpy(), mpc=mpc)
    return out


def solver_defaults(N):
    """cmpc_default_config's ADMM settings, for the fp64 oracle."""
    rho = 0.05 * N
    return dict(rho=rho, adaptive_interval=25, adaptive_tolerance=3.0, rho_lim=(0.1 * rho, 300.0))


def oracle_fixed(pb, b, K, mpc, **kw):
    x0, r, stance, xd, mu = pb.problem(b)
    return ca.solve_problem(x0, r, stance, xd, mu, DT, fixed_iters=K, rho=float(mpc.cfg.rho),
                            sigma=float(mpc.cfg.sigma), alpha=float(mpc.cfg.alpha), **kw)


# ------------------------------------------------------------------------------------------
def test_library_loaded_and_device():
    assert torch.cuda.is_available()
    assert pkg._capi.lib().cmpc_version() >= 1
    assert 10 in pkg._capi.supported_horizons()


@pytest.mark.parametrize("N,gaits", [(10, ("trot",)), (10, GAIT_NAMES), (5, ("trot",)),
                                     (20, ("pseudo_gallop",)), (30, ("trot",))])
def test_condense_matches_oracle(N, gaits):
    """H, g of cmpc_condense vs the plain recursion of reference src/mpc.py:64-136."""
    pb = synthetic_batch(6, N=N, gaits=gaits, seed=11)
    mpc = pkg.BatchedMPC(N=N, max_batch=pb.B)
    x0, r, mask, xd, mu = dev_args(pb)
    H, g = mpc.condense(x0, r, mask, xd)
    torch.cuda.synchronize()
    H, g = H.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64)
    for b in range(pb.B):
        px0, pr, pst, pxd, _ = pb.problem(b)
        # the kernel sees fp32 inputs: give the oracle the same rounded inputs
        px0, pr, pxd = (np.float32(px0).astype(np.float64), np.float32(pr).astype(np.float64),
                        np.float32(pxd).astype(np.float64))
        Ho, go, _, _, idx = srbd_qp.condensed_qp(px0, pr, pst, pxd, DT)
        sel = np.array([12 * i + 3 * l + k for (i, l) in idx for k in range(3)], dtype=int)
        scale = max(np.abs(Ho).max(), 1e-12)
        assert np.abs(H[b][np.ix_(sel, sel)] - Ho).max() <= 2e-5 * scale
        assert np.abs(g[b][sel] - go).max() <= 2e-4 * max(np.abs(go).max(), 1.0)
        off = np.setdiff1d(np.arange(12 * N), sel)
        assert np.all(H[b][off, :] == 0) and np.all(H[b][:, off] == 0) and np.all(g[b][off] == 0)


@pytest.mark.parametrize("N,gaits,B,K", [(10, ("trot",), 32, 60), (10, GAIT_NAMES, 32, 60),
                                         (5, ("trot",), 8, 60), (8, ("amble",), 8, 60),
                                         (12, ("trot",), 6, 60), (16, ("pronk",), 6, 60),
                                         (20, ("pseudo_gallop",), 6, 60), (30, ("trot",), 6, 60)])
def test_iterate_parity_fixed_iterations(N, gaits, B, K):
    """Same ADMM, same iteration count: fp32 CUDA vs fp64 oracle, full force vector.
    Once the fp64 iterate satisfies the reference's termination test (eps 1e-3) the forces
    must agree to 2e-2 N + 1e-3 |U| (iterate level: the null-space components of the force
    split drift at the fp32 floor; the north-star 1e-2 N is applied to the converged, unique
    quantities in the tight-parity tests below); while the iterate is still moving, the fp32
    Woodbury solve (relative error ~1e-3 of the step) allows 2e-3 max|U| in addition."""
    pb = synthetic_batch(B, N=N, gaits=gaits, seed=5)
    out = gpu_solve(pb, max_iter=K, check_every=100000, eps_abs=0.0, eps_rel=0.0, warm_mode=0,
                    adaptive_rho_interval=0)
    assert np.all(out["iters"] == K) and np.all(out["status"] == 0)
    worst, nconv = 0.0, 0
    for b in range(B):
        ref = oracle_fixed(pb, b, K, out["mpc"])
        x0, r, stance, xd, mu = pb.problem(b)
        Hx = ref["H"] @ ref["x"]
        conv = (ref["pri_res"] <= 1e-3 * (1 + max(np.abs(ref["x"]).max(initial=0), np.abs(ref["z"]).max(initial=0)))
                and ref["dua_res"] <= 1e-3 * (1 + max(np.abs(Hx).max(initial=0), np.abs(ref["y"]).max(initial=0),
                                                       np.abs(ref["g"]).max(initial=0))))
        err = np.abs(out["U"][b] - ref["U"])
        worst = max(worst, err.max())
        extra = 0.0 if conv else 2e-3 * np.abs(ref["U"]).max()
        nconv += bool(conv)
        assert np.all(err <= 2 * ATOL + extra + RTOL * np.abs(ref["U"])), (b, conv, err.max())
        assert close(out["X"][b].T, ref["X"], atol=1e-4, rtol=1e-3)
        # swing legs are exactly zero, bit for bit
        sw = np.repeat(pb.stance[b].reshape(-1) == 0, 3)
        assert np.all(out["U"][b].reshape(-1)[sw] == 0.0)
    assert nconv >= B // 2
    print(f"N={N} K={K} worst |dU| = {worst:.2e} N, {nconv}/{B} converged at the reference eps")


def test_converged_solve_meets_reference_eps():
    """Default settings (eps_abs = eps_rel = 1e-3 as in the reference's OSQP): every problem
    reports solved, and the residuals recomputed independently in fp64 from the returned
    primal/dual iterate are below the OSQP thresholds."""
    pb = synthetic_batch(64, N=10, seed=21)
    out = gpu_solve(pb, warm_mode=2)
    assert np.all(out["status"] == 1)
    assert out["iters"].max() < 1000
    mpc = out["mpc"]
    xw, yw = mpc.get_warm(pb.B)
    torch.cuda.synchronize()
    yw = yw.cpu().numpy().astype(np.float64)
    eps = 1e-3
    for b in range(64):
        x0, r, stance, xd, mu = pb.problem(b)
        x0, r, xd = (np.float32(x0).astype(np.float64), np.float32(r).astype(np.float64),
                     np.float32(xd).astype(np.float64))          # the data the kernel saw
        H, g, Sc, c0, idx = srbd_qp.condensed_qp(x0, r, stance, xd, DT)
        x = np.concatenate([out["U"][b][i, 3 * l:3 * l + 3] for (i, l) in idx])
        y = np.concatenate([yw[b][i, l] for (i, l) in idx])
        z = ca.project_frustum(x.reshape(-1, 3), mu).reshape(-1)       # distance to the feasible set
        pri = np.abs(x - z).max()
        dua = np.abs(H @ x + g + y).max()
        eps_p = eps + eps * max(np.abs(x).max(), np.abs(z).max())
        eps_d = eps + eps * max(np.abs(H @ x).max(), np.abs(y).max(), np.abs(g).max())
        # 1.0 x the reference's eps; 1e-4 relative slack for the fp32 evaluation of the same test
        assert pri <= eps_p * (1 + 1e-4), (b, pri, eps_p)
        assert dua <= eps_d * (1 + 1e-4), (b, dua, eps_d)
        # y is a valid multiplier: it lies in the normal cone of C at z (complementarity)
        assert abs(float(y @ (x - z))) <= 1e-2 * (1 + np.abs(y).max() * np.abs(x).max())
        # objective within 2 % of the tight optimum
        if b % 4 == 0:
            tight = ipm.solve_problem(x0, r, stance, xd, mu, DT)
            J = srbd_qp.objective(out["X"][b].T, xd)
            assert abs(J / tight["J"] - 1.0) < 2e-2


# Tight parity (north-star tolerance against the pinned oracle, every BASELINE config, no skipped
# problem) lives in tests/test_gpu_tight_parity.py.


def test_feasibility_and_masks_all_gaits():
    """Converged forces satisfy swing / bound / friction constraints of src/mpc.py:138-173
    (to the solver tolerance), for every gait incl. pronk flight phases and mu sweep."""
    pb = synthetic_batch(256, N=10, gaits=GAIT_NAMES, seed=3, mu=(0.3, 1.0))
    out = gpu_solve(pb)
    assert np.all(out["status"] == 1)
    F = out["U"].reshape(pb.B, 10, 4, 3)
    st = pb.stance.astype(bool)
    assert np.all(F[~st] == 0.0)
    mu = np.broadcast_to(pb.mu[:, None, None], st.shape)
    # OSQP primal tolerance of each problem: eps_abs + eps_rel * max(|x|, |z|)
    tol = (1e-3 + 1e-3 * np.abs(F).reshape(pb.B, -1).max(1))[:, None, None] * 1.05 + 1e-4
    fz = F[..., 2]
    assert np.all((fz >= 3.0 - tol)[st]) and np.all((fz <= 100.0 + tol)[st])
    # |x - z| <= eps with z in C  =>  |fx| <= mu fz + (1 + mu) eps
    assert np.all((np.abs(F[..., 0]) <= mu * fz + (1 + mu) * tol)[st])
    assert np.all((np.abs(F[..., 1]) <= mu * fz + (1 + mu) * tol)[st])
    # flight stages exist in the pronk problems and are all-zero
    assert (pb.stance.sum(-1) == 0).any()


def test_warm_start_semantics():
    """Reference semantics (src/mpc.py:270-271): previous primal solution, unshifted; zero
    duals.  A re-solve of the same problem from its own solution needs fewer iterations,
    reset_warm restores the cold behaviour, set_warm/get_warm round-trip."""
    pb = synthetic_batch(128, N=10, seed=17)
    mpc = pkg.BatchedMPC(N=10, max_batch=128, warm_mode=1)
    args = dev_args(pb)
    U1, _, s1 = mpc.solve(*args)
    it1 = s1.iters.cpu().numpy().copy()
    U1 = U1.cpu().numpy().copy()
    U2, _, s2 = mpc.solve(*args)
    it2 = s2.iters.cpu().numpy().copy()
    assert it2.mean() < it1.mean()
    xw, yw = mpc.get_warm(128)
    assert np.array_equal(xw.cpu().numpy(), U2.cpu().numpy())
    mpc.reset_warm()
    U3, _, s3 = mpc.solve(*args)
    assert np.array_equal(s3.iters.cpu().numpy(), it1)
    assert np.array_equal(U3.cpu().numpy(), U1)
    # warm_mode 2 carries the duals too: immediate convergence when re-solving
    mpc2 = pkg.BatchedMPC(N=10, max_batch=128, warm_mode=2)
    mpc2.solve(*args)
    _, _, s5 = mpc2.solve(*args)
    assert s5.iters.cpu().numpy().max() <= 5
    # set_warm on a fresh handle reproduces the warm-started solve of mpc
    mpc3 = pkg.BatchedMPC(N=10, max_batch=128, warm_mode=1)
    mpc3.set_warm(torch.from_numpy(U1).cuda())
    U4, _, s4 = mpc3.solve(*args)
    assert np.array_equal(s4.iters.cpu().numpy(), it2)
    assert np.array_equal(U4.cpu().numpy(), U2.cpu().numpy())


@pytest.mark.timeout(300)
def test_reserved_sm_rank_assignment_is_transparent():
    """Batches of 1.5-4 waves of CTAs take their launch rank when a CTA starts, the hardest ranks on
    reserved SMs (solve_kernel "work distribution"): results must be those of the plain one-CTA-per-rank
    launch - compared with the same problems solved in sub-batches of less than one wave - whatever the
    batch size, also when two such launches share the device (staged host path: two chunks, two streams)
    and when launches follow each other on one handle (the scheduling state is re-zeroed by the kernel)."""
    def in_sub_batches(q):
        parts = [gpu_solve(q.slice(a, min(a + 1000, q.B)), warm_mode=0) for a in range(0, q.B, 1000)]
        return np.concatenate([x["U"] for x in parts]), np.concatenate([x["iters"] for x in parts])
    big = synthetic_batch(5921, N=10, seed=0)
    big_U, big_it = in_sub_batches(big)
    pb = big.slice(0, 3000)                               # inside the range that uses the rank assignment
    ref_U, ref_it = big_U[:3000], big_it[:3000]
    mpc = pkg.BatchedMPC(N=10, max_batch=pb.B, warm_mode=0)
    args = dev_args(pb)
    for _ in range(3):                                   # repeated launches on the same scheduling state
        U, X, st = mpc.solve(*args)
        torch.cuda.synchronize()
        assert np.array_equal(U.cpu().numpy().astype(np.float64), ref_U)
        assert np.array_equal(st.iters.cpu().numpy(), ref_it)
    for B in (888, 889, 1332, 1333, 2048, 3552, 3553, 4736, 4737):   # around the 1.5- and 4-wave limits at 6 / 8 CTAs per SM
        d = gpu_solve(big.slice(0, B), warm_mode=0)
        assert np.array_equal(d["U"], big_U[:B]) and np.array_equal(d["iters"], big_it[:B]), B
        assert np.all(d["status"] == 1)
    # two launches of one handle on two streams (disjoint slot ranges), both in the range: they share the SMs, each
    # with its own reserved ones (the first version of the scheme could hang here)
    m2 = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a1, a2 = dev_args(big.slice(0, 2048)), dev_args(big.slice(2048, 4096))
    torch.cuda.synchronize()
    for _ in range(3):
        U1, _, st1 = m2.solve(*a1, slot0=0, stream=s1.cuda_stream)
        U2, _, st2 = m2.solve(*a2, slot0=2048, stream=s2.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(np.concatenate([U1.cpu().numpy(), U2.cpu().numpy()]).astype(np.float64), big_U[:4096])
        assert np.array_equal(np.concatenate([st1.iters.cpu().numpy(), st2.iters.cpu().numpy()]), big_it[:4096])
    # the chunked host path (pageable buffers, host_zero_copy = 0) launches its chunks plainly
    hin = [torch.from_numpy(a).pin_memory().numpy() for a in big.slice(0, 4096).f32()]
    m3 = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, host_zero_copy=0)
    Uh, Xh, sh = m3.solve_host(*hin)
    assert np.array_equal(Uh.astype(np.float64), big_U[:4096]) and np.array_equal(sh.iters, big_it[:4096])


def test_async_host_path_double_buffered():
    """cmpc_solve_host_async / cmpc_host_wait: two batches in flight on page-locked buffers give the
    results of the blocking call; pageable buffers are refused."""
    pbs = [synthetic_batch(1500, N=10, gaits=GAIT_NAMES, seed=s) for s in (21, 22, 23)]
    refs = [gpu_solve(pb, warm_mode=0) for pb in pbs]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    zeros = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
    mpc = pkg.BatchedMPC(N=10, max_batch=1500, warm_mode=0)
    ins = [[pin(a) for a in pb.f32()] for pb in pbs]
    outs = [(zeros((1500, 10, 12), torch.float32), zeros((1500, 11, 13), torch.float32), zeros((1500,), torch.int32),
             zeros((1500,), torch.float32), zeros((1500,), torch.float32), zeros((1500,), torch.int32)) for _ in pbs]
    tickets = [mpc.solve_host_async(*ins[k], out=outs[k]) for k in range(3)]       # three submissions in flight
    assert tickets == [0, 1, 2]
    for k in (2, 0, 1):                                                              # any waiting order
        mpc.host_wait(tickets[k])
        assert np.array_equal(outs[k][0].astype(np.float64), refs[k]["U"])
        assert np.array_equal(outs[k][1].astype(np.float64), refs[k]["X"])
        assert np.array_equal(outs[k][2], refs[k]["iters"]) and np.array_equal(outs[k][5], refs[k]["status"])
    for _ in range(12):                                                              # more submissions than kept events
        t = mpc.solve_host_async(*ins[0], out=outs[0])
    mpc.host_wait(3)
    mpc.host_wait(t)
    assert np.array_equal(outs[0][0].astype(np.float64), refs[0]["U"])
    with pytest.raises(pkg._capi.CmpcError):
        mpc.solve_host_async(*pbs[0].f32(), out=outs[0])                             # pageable inputs
    with pytest.raises(pkg._capi.CmpcError):
        mpc.host_wait(t + 1)


def test_host_path_matches_device_path_and_sharding():
    """cmpc_solve_host (pinned staging, chunked streams) returns bit-identical results to
    cmpc_solve, and solving two half batches equals solving the whole batch (problems are
    independent - the property the multi-GPU sharding relies on)."""
    pb = synthetic_batch(2500, N=10, gaits=GAIT_NAMES, seed=9)
    d = gpu_solve(pb, warm_mode=0)
    mpc = pkg.BatchedMPC(N=10, max_batch=pb.B, warm_mode=0)
    U, X, st = mpc.solve_host(*pb.f32())          # pageable buffers: staged through pinned arenas
    assert np.array_equal(U.astype(np.float64), d["U"])
    assert np.array_equal(X.astype(np.float64), d["X"])
    assert np.array_equal(st.iters, d["iters"]) and np.array_equal(st.status, d["status"])
    # page-locked caller buffers: (1) read / written in place by the kernel over PCIe (default),
    # (0) chunked cudaMemcpyAsync pipeline - both bit-identical to the device path
    hin = [torch.from_numpy(a).pin_memory().numpy() for a in pb.f32()]
    pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
    for zero_copy in (1, 0):
        m2 = pkg.BatchedMPC(N=10, max_batch=pb.B, warm_mode=0, host_zero_copy=zero_copy)
        hout = (pin((pb.B, 10, 12), torch.float32), pin((pb.B, 11, 13), torch.float32),
                pin((pb.B,), torch.int32), pin((pb.B,), torch.float32), pin((pb.B,), torch.float32),
                pin((pb.B,), torch.int32))
        Uh, Xh, sh = m2.solve_host(*hin, out=hout)
        assert np.array_equal(Uh.astype(np.float64), d["U"]), zero_copy
        assert np.array_equal(Xh.astype(np.float64), d["X"]), zero_copy
        assert np.array_equal(sh.iters, d["iters"]) and np.array_equal(sh.status, d["status"])
        # optional outputs may be NULL
        U3, X3, s3 = m2.solve_host(*hin, want_X=False,
                                   out=(hout[0], None, hout[2], hout[3], hout[4], hout[5]))
        assert X3 is None and np.array_equal(U3.astype(np.float64), d["U"])
    h = pb.B // 2
    a = gpu_solve(pb.slice(0, h), warm_mode=0)
    b = gpu_solve(pb.slice(h, pb.B), warm_mode=0)
    assert np.array_equal(np.concatenate([a["U"], b["U"]]), d["U"])
    assert np.array_equal(np.concatenate([a["iters"], b["iters"]]), d["iters"])


def test_edge_cases():
    mpc = pkg.BatchedMPC(N=10, max_batch=8)
    pb = synthetic_batch(8, N=10, seed=1)
    # empty batch
    U, X, st = mpc.solve(*[t[:0].contiguous() for t in dev_args(pb)])
    assert U.shape == (0, 10, 12)
    # single problem
    U1, _, s1 = mpc.solve(*[t[:1].contiguous() for t in dev_args(pb)])
    assert int(s1.status[0]) == 1
    # all legs in swing for the whole horizon -> zero forces, solved in 0 iterations
    args = dev_args(pb)
    args[2] = torch.zeros_like(args[2])
    mpc.reset_warm()
    U0, X0, s0 = mpc.solve(*args)
    assert torch.all(U0 == 0) and torch.all(s0.status == 1) and torch.all(s0.iters == 0)
    # free fall: v_z decreases by g*dt per stage (src/mpc.py:94, A[11,12] = 1)
    Xn = X0.cpu().numpy()
    np.testing.assert_allclose(Xn[:, 1:, 11] - Xn[:, :-1, 11], -9.81 * 0.01, rtol=1e-4)
    # NaN input -> status -1, other problems unaffected
    args = dev_args(pb)
    args[0][3, 4] = float("nan")
    mpc.reset_warm()
    _, _, sn = mpc.solve(*args)
    stat = sn.status.cpu().numpy()
    assert stat[3] == -1 and np.all(np.delete(stat, 3) == 1)
    # batch larger than the handle
    big = synthetic_batch(9, N=10, seed=1)
    with pytest.raises(pkg.CmpcError):
        mpc.solve(*dev_args(big))
    with pytest.raises(pkg.CmpcError):
        pkg.BatchedMPC(N=61, max_batch=4)         # beyond the largest compiled kernel (60)
    with pytest.raises(pkg.CmpcError):
        pkg.BatchedMPC(N=0, max_batch=4)
    with pytest.raises(pkg.CmpcError):
        pkg.BatchedMPC(N=10, max_batch=4, w=[1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 0])


@pytest.mark.parametrize("B,N,gaits,mu", [(4096, 10, ("trot",), (1.0, 1.0)),          # BASELINE config 2
                                          (8192, 10, GAIT_NAMES, (0.3, 1.0)),           # config 3 (one GPU's shard), mu sweep
                                          (16384, 30, ("trot",), (1.0, 1.0))])          # config 4
def test_full_size_properties(B, N, gaits, mu):
    """BASELINE.json configs at full per-GPU size: size-independent properties - all solved,
    forces inside the reference's feasible set (src/mpc.py:148-173: swing forces exactly zero,
    f_min <= fz <= f_max, |fx|,|fy| <= mu fz up to the solver tolerance), dynamics consistency of
    the returned X with the returned U (forward-Euler recursion of src/mpc.py:113-117 in fp64),
    permutation invariance (the property the hardest-first schedule and the sharding rely on)."""
    pb = synthetic_batch(B, N=N, gaits=gaits, seed=0, mu=mu)
    out = gpu_solve(pb, warm_mode=0)
    assert np.all(out["status"] == 1)
    U = out["U"].reshape(B, N, 4, 3)
    st = pb.stance.astype(bool)                                      # (B,N,4)
    assert np.all(U[~st] == 0.0)
    fz, fx, fy = U[..., 2][st], U[..., 0][st], U[..., 1][st]
    mu_b = np.broadcast_to(pb.mu[:, None, None], st.shape)[st]
    # OSQP primal tolerance of each problem, eps_abs + eps_rel * max(|x|, |z|), on |x - z| with z in C
    tol_b = (1e-3 + 1e-3 * np.abs(U).reshape(B, -1).max(1)) * (1 + 1e-4) + 1e-5
    tol = np.broadcast_to(tol_b[:, None, None], st.shape)[st]
    assert np.all(fz >= 3.0 - tol) and np.all(fz <= 100.0 + tol)
    assert np.all(np.abs(fx) <= mu_b * fz + (1 + mu_b) * tol) and np.all(np.abs(fy) <= mu_b * fz + (1 + mu_b) * tol)
    rng = np.random.default_rng(0)
    for b in rng.integers(0, B, 8):
        x0, r, stance, xd, mu_ = pb.problem(b)
        x0, r = np.float32(x0).astype(np.float64), np.float32(r).astype(np.float64)
        X = np.zeros((13, N + 1))
        X[:, 0] = x0
        Ad = np.eye(13) + DT * srbd_qp.continuous_A(x0[2])
        for i in range(N):
            X[:, i + 1] = Ad @ X[:, i] + DT * srbd_qp.continuous_B(x0[2], r[i]) @ out["U"][b][i]
        assert close(out["X"][b].T, X, atol=2e-5 * (N / 10) ** 2, rtol=1e-4)
    perm = rng.permutation(B)
    pb2 = pkg.problems.ProblemBatch(pb.x0[perm], pb.r[perm], pb.stance[perm], pb.x_des[perm],
                                    pb.mu[perm], pb.gait_id[perm], pb.tick[perm])
    out2 = gpu_solve(pb2, warm_mode=0)
    assert np.array_equal(out2["U"], out["U"][perm])


# ------------------------------------------------------------------------------------------
# Drop-in MPC class (reference src/mpc.py:8-318 interface) on the golden run's states
# ------------------------------------------------------------------------------------------
class _FakeLite3:
    """Stands in for the reference's Lite3Controller.retrieve_state (src/main.py:286-350)."""

    def __init__(self, gold):
        self.gold, self.t = gold, 0

    def retrieve_state(self):
        s, f = self.gold["state"][self.t], self.gold["feet"][self.t]
        d = {leg: {"pos": np.concatenate([np.zeros(3), f[l]]), "vel": np.zeros(6)}
             for l, leg in enumerate(pkg.LEGS)}
        d["TORSO"] = {"pos": s[0:3].copy(), "vel": s[6:9].copy()}
        d["com"] = {"pos": s[3:6].copy(), "vel": s[9:12].copy()}
        return d


class _RefLikePlanner:
    """Object with the reference FootstepPlanner's `.plan` layout (list of dicts)."""

    def __init__(self, gp):
        self.plan = [{"pos": {leg: gp.pos[s, l] for l, leg in enumerate(pkg.LEGS)},
                      "ang": gp.ang[s], "ss_duration": gp.ss, "ds_duration": gp.ds,
                      "feet_id": list(gp.feet_id[s])} for s in range(gp.n_steps)]


class _Logger:
    def __init__(self):
        self.track, self.pred = [], []

    def log_tracking_data(self, actual, des):
        self.track.append((actual, des))

    def log_mpc_predictions(self, x_log, x_des, forces_pred, t):
        self.pred.append((t, x_log.shape, x_des.shape, forces_pred.shape))


def test_mpc_dropin_on_golden_states(gold):
    from oracle.replay import params_from_golden, initial_from_golden
    params = params_from_golden(gold, N=10)
    initial = initial_from_golden(gold)
    gp = pkg.GaitPlan.from_initial(initial, params)
    lite3, logger = _FakeLite3(gold), _Logger()
    mpc = pkg.MPC(lite3=lite3, initial=initial, footstep_planner=_RefLikePlanner(gp), params=params)
    warm = None
    for t in list(range(0, 40)) + [80]:
        lite3.t = t
        if t == 80:      # jump: bring the reference accumulators where the real run had them
            mpc.com_pos_start[:] = gold["desired"][80][3:6]
            mpc.yaw_start = gold["desired"][80][2]
            warm = None
            mpc.solver.reset_warm()
        forces = mpc.solve(t, logger)
        assert set(forces) == set(pkg.LEGS) and all(f.shape == (3,) and f.dtype == np.float64
                                                    for f in forces.values())
        assert mpc.x.shape == (13, 1) and mpc.x_log.shape == (12, 11) and mpc.x_plot.shape == (3, 11)
        assert mpc.u.shape == (12,) and mpc.u_plot.shape == (12, 10)
        # desired state of this tick is the reference's, bit for bit (src/mpc.py:202-214,261-262)
        assert np.array_equal(logger.track[-1][1], gold["desired"][t])
        # same algorithm in fp64 with the reference's warm-start semantics
        x0 = np.concatenate([gold["state"][t], [params["g"]]])
        xd = logger.track[-1][1]
        xdes = pkg.desired_trajectory(10, 0.01, initial["roll"], initial["pitch"], xd[2], xd[3:6],
                                      *pkg.reference_velocity(gp, t, params), params["g"])
        r, stance = pkg.assemble_tick(gp, t, 10, 0.01, x0, gold["feet"][t], xdes)
        H, g, Sc, c0, idx = srbd_qp.condensed_qp(np.float32(x0).astype(float), np.float32(r).astype(float),
                                                 stance, np.float32(xdes).astype(float), 0.01)
        xw = None if warm is None else np.concatenate([warm[i, 3 * l:3 * l + 3] for (i, l) in idx])
        ref = ca.admm(H, g, 1.0, check_every=5, x=xw, **solver_defaults(10))
        assert ref["status"] == 1
        assert abs(mpc.iters - ref["iters"]) <= 10
        # EVERY tick: the same ADMM in fp64 stopped at the iteration count the kernel stopped at,
        # from the same warm start -> forces, states and per-stage wrench at iterate-level tolerance
        same = ca.admm(H, g, 1.0, x=xw, fixed_iters=mpc.iters, **solver_defaults(10)) if mpc.iters else ref
        U = np.zeros((10, 12))
        for s_, (i, l) in enumerate(idx):
            U[i, 3 * l:3 * l + 3] = same["x"][3 * s_:3 * s_ + 3]
        Ug = mpc.u_plot.T
        tolU = 2 * ATOL + 2e-3 * np.abs(U).max() + RTOL * np.abs(U)
        assert np.all(np.abs(Ug - U) <= tolU), (t, np.abs(Ug - U).max(), mpc.iters)
        X = c0 + (Sc @ same["x"]).reshape(11, 13).T
        assert close(mpc.x_log, X[:12], atol=1e-4, rtol=1e-3), t
        rr = np.float32(r).astype(float)
        Wg, Wr = srbd_qp.stage_wrench(Ug, rr), srbd_qp.stage_wrench(U, rr)
        assert np.all(np.abs(Wg - Wr) <= 4 * ATOL + 2e-3 * np.abs(Wr).max() + RTOL * np.abs(Wr)), (t, np.abs(Wg - Wr).max())
        Jg, Jr = srbd_qp.objective(np.vstack([mpc.x_log, np.full((1, 11), params["g"])]), xdes), \
            srbd_qp.objective(X, xdes)
        assert abs(Jg / Jr - 1) < 1e-3
        warm = mpc.u_plot.T.copy()
    assert len(logger.track) == 41 and [p[0] for p in logger.pred] == [0, 80]
    assert logger.pred[0][1:] == ((12, 11), (12, 11), (4, 10))
    # stance legs of tick 0 carry the robot: sum fz close to m*g
    lite3.t = 0


def test_mpc_dropin_raises_when_not_solved(gold):
    """CasADi raises RuntimeError when OSQP does not report 'solved' (SURVEY.md 8b)."""
    from oracle.replay import params_from_golden, initial_from_golden
    params = params_from_golden(gold, N=10)
    initial = initial_from_golden(gold)
    gp = pkg.GaitPlan.from_initial(initial, params)
    mpc = pkg.MPC(lite3=_FakeLite3(gold), initial=initial, footstep_planner=gp, params=params,
                  max_iter=3)
    with pytest.raises(RuntimeError):
        mpc.solve(0, _Logger())
