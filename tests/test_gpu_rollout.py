"""GPU tests of the on-device parameter assembly and the closed-loop rollout
engine (8f.2, BASELINE config 5) against their numpy/fp64 restatements."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mpc_b200 as pkg                                             # noqa: E402
from mpc_b200 import _capi                                         # noqa: E402
from mpc_b200.problems import GAIT_NAMES, DT, GRAVITY, synthetic_batch   # noqa: E402
from mpc_b200.solver import _ptr                                   # noqa: E402
from oracle import condensed_admm as ca, srbd_qp                    # noqa: E402


def host_assemble(ro, t, x, yaw_start, com_start):
    """numpy restatement of reference src/mpc.py:178-255 for all robots of a rollout."""
    B, N, plan = ro.B, ro.N, ro.plan
    tt = np.full(B, t)
    last = plan.step_index(tt) == 20 - 1
    v = np.where(last[:, None], 0.0, ro.v_ref.cpu().numpy().astype(np.float64))
    om = np.where(last, 0.0, ro.omega_ref.cpu().numpy().astype(np.float64))
    k = np.arange(N + 1)
    xd = np.zeros((B, N + 1, 13))
    xd[:, :, 2] = yaw_start[:, None] + om[:, None] * DT * k
    xd[:, :, 3:6] = com_start[:, None, :] + v[:, None, :] * DT * k[None, :, None]
    xd[:, :, 8] = om[:, None]
    xd[:, :, 9:12] = v[:, None, :]
    xd[:, :, 12] = GRAVITY
    ticks = t + np.arange(N)[None].repeat(B, 0)
    feet = plan.foot_position(ticks)
    r = np.empty((B, N, 4, 3))
    r[:, 0] = feet[:, 0] - x[:, None, 3:6]
    r[:, 1:] = feet[:, 1:] - xd[:, 1:N, None, 3:6]
    return xd, r, plan.stance_mask(ticks)


@pytest.mark.parametrize("N", [10, 5])
def test_assemble_matches_host(N):
    ro = pkg.ClosedLoopRollout(96, N=N, gaits=GAIT_NAMES, seed=4)
    rng = np.random.default_rng(1)
    L = _capi.lib()
    for t in (0, 7, 19, 20, 33, 95, 150, 379, 385, 400, 5000):
        x = np.zeros((ro.B, 13)); x[:, 3:6] = rng.normal(0, 0.3, (ro.B, 3)); x[:, 2] = rng.normal(0, 1, ro.B)
        ys = rng.normal(0, 0.5, ro.B); cs = rng.normal(0, 0.5, (ro.B, 3))
        ro.tick.fill_(t)
        ro.x.copy_(torch.from_numpy(x.astype(np.float32)))
        ro.yaw_start.copy_(torch.from_numpy(ys.astype(np.float32)))
        ro.com_start.copy_(torch.from_numpy(cs.astype(np.float32)))
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _capi.check(L.cmpc_assemble(ro.mpc._h, ro.B, C.byref(ro.gt), _ptr(ro.tick), _ptr(ro.x),
                                    _ptr(ro.yaw_start), _ptr(ro.com_start), _ptr(ro.x_des),
                                    _ptr(ro.r), _ptr(ro.mask), s))
        torch.cuda.synchronize()
        xd, r, stance = host_assemble(ro, t, x.astype(np.float32).astype(np.float64),
                                      ys.astype(np.float32).astype(np.float64),
                                      cs.astype(np.float32).astype(np.float64))
        assert np.array_equal(ro.mask.cpu().numpy(), pkg.stance_bits(stance)), t     # bit-exact
        np.testing.assert_allclose(ro.x_des.cpu().numpy(), xd, atol=2e-6, rtol=1e-6)
        np.testing.assert_allclose(ro.r.cpu().numpy(), r, atol=3e-6, rtol=1e-5)


def test_plant_step_matches_srbd_model():
    ro = pkg.ClosedLoopRollout(32, N=10, seed=2)
    rng = np.random.default_rng(3)
    x = rng.normal(0, 0.3, (32, 13)); x[:, 12] = GRAVITY
    U = rng.uniform(-20, 60, (32, 10, 12)); r = rng.normal(0, 0.2, (32, 10, 4, 3))
    ro.x.copy_(torch.from_numpy(x.astype(np.float32)))
    ro.r.copy_(torch.from_numpy(r.astype(np.float32)))
    ro.out[0].copy_(torch.from_numpy(U.astype(np.float32)))
    ro.x_des.zero_()
    ro.tick.fill_(40)
    L = _capi.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _capi.check(L.cmpc_plant_step(ro.mpc._h, ro.B, C.byref(ro.gt), _ptr(ro.tick), _ptr(ro.x), _ptr(ro.r),
                                  _ptr(ro.out[0]), _ptr(ro.x_des), _ptr(ro.yaw_start),
                                  _ptr(ro.com_start), _ptr(ro.track_err), s))
    torch.cuda.synchronize()
    assert int(ro.tick.item()) == 41
    xn = ro.x.cpu().numpy()
    xf, rf, Uf = x.astype(np.float32).astype(float), r.astype(np.float32).astype(float), U.astype(np.float32).astype(float)
    for b in range(32):                       # reference src/mpc.py:86-117
        ref = xf[b] + DT * (srbd_qp.continuous_A(xf[b, 2]) @ xf[b] + srbd_qp.continuous_B(xf[b, 2], rf[b, 0]) @ Uf[b, 0])
        np.testing.assert_allclose(xn[b], ref, atol=2e-5, rtol=1e-5)
    # reference accumulators advanced by v*dt (trot v_ref_x = 0.08)
    np.testing.assert_allclose(ro.com_start.cpu().numpy()[:, 0], pkg.problems.NOMINAL_COM[0] + 0.08 * DT, atol=1e-7)


def test_closed_loop_matches_fp64_loop():
    """4 robots x 15 ticks: the GPU loop (assemble -> solve -> plant) against the same loop in
    numpy fp64 with the oracle ADMM (same warm-start semantics).  Forces are compared through
    the state trajectory they produce (unique), which must agree closely."""
    B, T = 4, 15
    ro = pkg.ClosedLoopRollout(B, N=10, seed=5, mu=(0.5, 1.0))
    x = ro.x.cpu().numpy().astype(np.float64)
    ys = np.zeros(B); cs = ro.com_start.cpu().numpy().astype(np.float64)
    warm = [None] * B
    mu = ro.mu_host
    for t in range(T):
        xd, r, stance = host_assemble(ro, t, x, ys, cs)
        for b in range(B):
            H, g, Sc, c0, idx = srbd_qp.condensed_qp(x[b], r[b], stance[b], xd[b].T, DT)
            xw = None if warm[b] is None else np.concatenate([warm[b][i, 3 * l:3 * l + 3] for (i, l) in idx])
            res = ca.admm(H, g, mu[b], check_every=5, x=xw, rho=0.5, adaptive_interval=25,
                          adaptive_tolerance=3.0, rho_lim=(0.05, 300.0))
            U = np.zeros((10, 12))
            for s_, (i, l) in enumerate(idx):
                U[i, 3 * l:3 * l + 3] = res["x"][3 * s_:3 * s_ + 3]
            warm[b] = U
            x[b] = x[b] + DT * (srbd_qp.continuous_A(x[b, 2]) @ x[b] + srbd_qp.continuous_B(x[b, 2], r[b, 0]) @ U[0])
        cs = cs + ro.v_ref.cpu().numpy().astype(np.float64) * DT
        ro.step()
    torch.cuda.synchronize()
    xg = ro.x.cpu().numpy()
    # velocities respond to the forces directly: m dv = f dt, so 1e-3 m/s ~ 0.9 N integrated
    np.testing.assert_allclose(xg[:, 3:6], x[:, 3:6], atol=2e-4)
    np.testing.assert_allclose(xg[:, 9:12], x[:, 9:12], atol=3e-3)
    np.testing.assert_allclose(xg[:, 0:3], x[:, 0:3], atol=2e-3)


def test_rollout_graph_equals_eager_and_stays_upright():
    a = pkg.ClosedLoopRollout(256, N=10, gaits=GAIT_NAMES, seed=7)
    b = pkg.ClosedLoopRollout(256, N=10, gaits=GAIT_NAMES, seed=7)
    a.run(61, use_graph=True, ticks_per_graph=10)
    b.run(61, use_graph=False)
    sa, sb = a.summary(), b.summary()
    assert sa["ticks"] == sb["ticks"] == 61
    assert torch.equal(a.x, b.x) and torch.equal(a.out[0], b.out[0])
    assert sa["finite"] and sa["unsolved"] == 0
    assert 0.2 < sa["com_z_min"] and sa["com_z_max"] < 0.4          # nominal height 0.285
    assert sa["rms_pos_err"] < 0.05
    # warm start pays: fewer iterations per tick than a cold solve (~27-33 on mixed gaits)
    assert sa["mean_iters"] < 25


def test_log_export_has_the_reference_schema(tmp_path, gold):
    """The exported pickle has the keys / shapes of the reference's simulation_log.pkl
    (reference src/logger.py:21-46) - checked by running this repo's own fixture extractor's
    field accesses and the data accesses of the reference's plot.py (src/plot.py:22-83) on it."""
    import pickle
    from mpc_b200 import kinematics as kin
    ro = pkg.ClosedLoopRollout(8, N=10, seed=1)
    path = tmp_path / "simulation_log.pkl"
    T = 85
    log = pkg.logexport.rollout_log(ro, T, robot=3, path=str(path))
    log = pickle.load(open(path, "rb"))
    assert set(log) == {"mpc_freq", "sim_params", "total_sim_steps", "time array", "FEET POS",
                        "MPC PREDICTIONS", "TRACKING PERFORMANCE", "FORCES", "CONTROL EFFORT"}
    assert log["total_sim_steps"] == T and log["time array"] == list(range(T))
    state = np.array(log["TRACKING PERFORMANCE"]["actual"])
    desired = np.array(log["TRACKING PERFORMANCE"]["desired"])
    assert state.shape == desired.shape == (T, 12)
    forces = np.stack([np.stack([np.array(log["FORCES"][l][c]) for c in "xyz"], 1) for l in pkg.LEGS], 1)
    assert forces.shape == (T, 4, 3) and np.all(forces[:, :, 2] >= -1e-6)
    assert [p["time step"] for p in log["MPC PREDICTIONS"]] == [0, 80]
    p0 = log["MPC PREDICTIONS"][0]
    assert p0["predicted_state"].shape == (12, 11) and p0["desired_state"].shape == (12, 11)
    assert p0["predicted forces"].shape == (4, 10)
    assert np.allclose(p0["predicted_state"][:, 0], state[0], atol=1e-6)      # X_0 = x0 (src/mpc.py:113)
    assert set(log["sim_params"]) >= {"g", "h", "ss_duration", "ds_duration", "first_swing", "µ", "N",
                                      "v_com_ref", "theta_dot", "total_steps", "world_time_step"}
    assert log["mpc_freq"] > 1000.0          # solves per second of the single robot's tick
    # --- the reference's plot.py data accesses (src/plot.py:22-83) -------------------------------
    total_sim_steps, time_step = log["total_sim_steps"], log["sim_params"]["world_time_step"]
    for data in log["MPC PREDICTIONS"]:                                            # plot.py:27-35
        Np = data["predicted_state"].shape[1] - 1
        assert data["predicted_state"][3:6, :Np].shape == data["desired_state"][3:6, :Np].shape == (3, Np)
    com_position = np.array([elem[3:6] for elem in log["TRACKING PERFORMANCE"]["actual"]]).T    # plot.py:39-43
    com_desired = np.array([elem[3:6] for elem in log["TRACKING PERFORMANCE"]["desired"]]).T
    assert com_position.shape == com_desired.shape == (3, total_sim_steps) and time_step == 0.01
    for foot_name in pkg.LEGS:
        foot_traj, foot_des_traj = log["FEET POS"][foot_name]["actual"], log["FEET POS"][foot_name]["des"]
        foot_z = np.array([elem[2] for elem in foot_traj]).T                                    # plot.py:61-62
        foot_z_des = np.array([elem[2] for elem in foot_des_traj]).T
        assert foot_z.shape == foot_z_des.shape == (total_sim_steps,)
        effort = log["CONTROL EFFORT"][foot_name]                                               # plot.py:82-83
        assert list(effort) == [f"{foot_name[:2]}_{j}" for j in ("HipX", "HipY", "Knee")]       # src/logger.py:39-42
        assert all(len(v) == total_sim_steps and np.all(np.isfinite(v)) for v in effort.values())
        assert set(log["FORCES"][foot_name]) == {"x", "y", "z"}
    # desired feet are the controller's references (src/main.py:159-167), not a copy of the actual feet:
    # planned foothold for stance legs, swing polynomial (lifted by up to step_height) for swing legs
    fl_des = np.array(log["FEET POS"]["FL_FOOT"]["des"])
    assert fl_des[:, 2].max() > 0.05 and fl_des[:, 2].min() >= 0.0
    # joint torques of a standing tick: tau = J'(-f) (src/main.py:214) with the URDF leg geometry
    f0 = forces[0]                                                    # all-stance at tick 0
    base = state[0, 3:6] - kin.rotvec_matrix(state[0, 0:3]) @ kin.nominal_com_offset()
    feet0 = np.stack([log["FEET POS"][l]["actual"][0] for l in pkg.LEGS])
    q = kin.leg_ik_batch((feet0 - base) @ kin.rotvec_matrix(state[0, 0:3]))
    m = kin.leg_kinematics(base, state[0, 0:3], np.zeros(3), np.zeros(3), q, np.zeros((4, 3)))
    for l, leg in enumerate(pkg.LEGS):
        tau = np.array([log["CONTROL EFFORT"][leg][f"{leg[:2]}_{j}"][0] for j in ("HipX", "HipY", "Knee")])
        assert np.allclose(tau, m["J"][l].T @ -f0[l], atol=2e-3 + 1e-3 * np.abs(tau).max())
        assert np.abs(tau).max() > 0.5                                # the legs carry the robot


def test_factorisation_cache_reuses_on_static_data_and_keeps_the_answers():
    """cfg.cache_factorization: a second solve of the same problems reuses the cached -P^-1 (no
    sweep) and returns the same converged forces; moving a lever arm beyond the tolerance, flipping
    a contact bit or resetting the warm state forces a fresh factorisation."""
    pb = synthetic_batch(256, N=10, gaits=GAIT_NAMES, seed=21)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    ref = pkg.BatchedMPC(N=10, max_batch=256, warm_mode=0)
    U0, X0, s0 = ref.solve(*args)
    mpc = pkg.BatchedMPC(N=10, max_batch=256, warm_mode=0, cache_factorization=1, cache_max_iter=1000)
    U1, X1, s1 = mpc.solve(*args)
    m1 = mpc.cache_meta(256).cpu().numpy()
    assert np.all(m1[:, 3] == 0)                                    # nothing cached yet
    assert torch.equal(U1, U0) and torch.equal(s1.iters, s0.iters)  # a miss is the plain path, bit for bit
    U2, X2, s2 = mpc.solve(*args)
    m2 = mpc.cache_meta(256).cpu().numpy()
    ok = s1.status.cpu().numpy() == 1
    adapted = m1[:, 0] != np.float32(mpc.cfg.rho)                   # rho was adapted: cached factor is for another rho
    assert np.all(m2[ok & ~adapted, 3] == 1) and np.all(m2[adapted, 3] == 0)
    assert torch.equal(U2, U1) and torch.equal(s2.iters, s1.iters)  # same factor, same iterates
    # a lever arm moved by 1 mm stays inside the tolerance: reused, answers within the solver tolerance
    moved = [a.clone() for a in args]
    moved[1][:, 3, 1, 0] += 1e-3
    U3, X3, s3 = mpc.solve(*moved)
    m3 = mpc.cache_meta(256).cpu().numpy()
    assert np.all(m3[ok & ~adapted, 3] == 1)
    U3f, X3f, s3f = ref.solve(*moved)
    good = (s3.status == 1) & (s3f.status == 1)
    assert float(good.float().mean()) > 0.99
    J = lambda X, xd: ((X[:, :, :12] - xd[:, :, :12]) ** 2 * torch.tensor(
        [1e4, 2.7e4, 1e4, 2.7e5, 2.7e5, 2.7e5, 1e4, 1e4, 1e4, 1.6e4, 1.6e4, 1.6e4], device=X.device)).sum((1, 2))
    rel = (J(X3, moved[3]) / J(X3f, moved[3]) - 1).abs()[good]
    assert float(rel.max()) < 2e-2                                   # both within eps of the same optimum
    # 5 mm: outside the tolerance -> fresh factorisation, identical to the plain solver
    moved[1][:, 3, 1, 0] += 4e-3
    U4, _, s4 = mpc.solve(*moved)
    assert np.all(mpc.cache_meta(256).cpu().numpy()[:, 3] == 0)
    U4f, _, s4f = ref.solve(*moved)
    assert torch.equal(U4, U4f) and torch.equal(s4.iters, s4f.iters)
    # a flipped contact bit -> miss; reset_warm drops the cache
    flipped = [a.clone() for a in moved]
    flipped[2][:, 5] ^= 1
    mpc.solve(*flipped)
    assert np.all(mpc.cache_meta(256).cpu().numpy()[:, 3] == 0)
    mpc.solve(*flipped)
    assert np.mean(mpc.cache_meta(256).cpu().numpy()[:, 3]) > 0.8
    mpc.reset_warm()
    mpc.solve(*flipped)
    assert np.all(mpc.cache_meta(256).cpu().numpy()[:, 3] == 0)


def test_closed_loop_with_factorisation_cache_tracks_like_without():
    a = pkg.ClosedLoopRollout(256, N=10, gaits=("trot",), mu=(0.3, 1.0), seed=0, total_steps=4, cache_factorization=0)
    b = pkg.ClosedLoopRollout(256, N=10, gaits=("trot",), mu=(0.3, 1.0), seed=0, total_steps=4, cache_factorization=1)
    a.run(200, use_graph=False)
    b.run(200, use_graph=False)
    sa, sb = a.summary(), b.summary()
    assert sa["unsolved"] == 0 and sb["unsolved"] == 0 and sb["finite"]
    assert sb["cache_hit_frac"] > 0.3                    # the standing phase after the 4 planned steps
    assert abs(sb["rms_pos_err"] - sa["rms_pos_err"]) < 2e-4 and abs(sb["rms_ang_err"] - sa["rms_ang_err"]) < 5e-4
    xa, xb = a.x.cpu().numpy(), b.x.cpu().numpy()
    assert np.abs(xa[:, :12] - xb[:, :12]).max() < 2e-3


# ------------------------------------------------------------------------------------------
# SURVEY.md 8f.1 pinned on the reference: cmpc_assemble against fixtures recorded from the
# reference's own planner / swing generator (tests/golden/gait_golden.npz)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["trot", "pseudo_gallop", "pseudo_gallop_ds4", "amble", "pronk",
                                  "trot_turning", "trot_ss7", "stand"])
def test_assemble_matches_reference_planner_fixtures(gait_gold, name):
    """Contact masks bit-exact with reference get_phase_at_time (src/footstep_planner.py:226-246)
    and look-ahead feet (src/mpc.py:306-318 + src/foot_trajectory_generator.py:27-96) to fp32
    rounding, for every tick of the recorded schedule and past its end."""
    g = lambda k: gait_gold[f"{name}/{k}"]
    N = 10
    pos, fid = g("pos"), g("feet_id")                    # the reference planner's own plan
    S, T = pos.shape[0], g("mask").shape[0]
    dev = torch.device("cuda", 0)
    mpc = pkg.BatchedMPC(N=N, max_batch=1)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    plan_pos = f32(pos[None])
    feet_id = torch.from_numpy(pkg.stance_bits(fid)[None].copy()).to(dev)
    ss = torch.tensor([int(g("ss"))], dtype=torch.int32, device=dev)
    ds = torch.tensor([int(g("ds"))], dtype=torch.int32, device=dev)
    zero3, zero1, zero2 = f32(np.zeros((1, 3))), f32(np.zeros(1)), f32(np.zeros((1, 2)))
    gt = _capi.GaitTables(plan_pos=plan_pos.data_ptr(), feet_id=feet_id.data_ptr(), ss=ss.data_ptr(),
                          ds=ds.data_ptr(), v_ref=zero3.data_ptr(), omega_ref=zero1.data_ptr(),
                          rp0=zero2.data_ptr(), S=S, total_steps=int(g("total_steps")), step_height=0.08,
                          g=GRAVITY)
    x = f32(np.zeros((1, 13)))                           # com at the origin: r = foot position
    x_des = torch.empty((1, N + 1, 13), dtype=torch.float32, device=dev)
    r = torch.empty((1, N, 4, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((1, N), dtype=torch.uint8, device=dev)
    tick = torch.zeros(1, dtype=torch.int32, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    worst = 0.0
    for t in range(0, T - N):
        tick.fill_(t)
        _capi.check(_capi.lib().cmpc_assemble(mpc._h, 1, C.byref(gt), _ptr(tick), _ptr(x), _ptr(zero1),
                                              _ptr(zero3), _ptr(x_des), _ptr(r), _ptr(mask), s))
        torch.cuda.synchronize()
        assert np.array_equal(mask.cpu().numpy()[0], pkg.stance_bits(g("mask")[t:t + N])), t    # bit-exact
        foot = g("foot")[t:t + N]
        err = np.abs(r.cpu().numpy()[0].astype(np.float64) - foot)
        worst = max(worst, err.max())
        assert err.max() <= 2.5e-7, (t, err.max())        # ~1 ulp of fp32 at the feet's 0.5-2 m range
    print(f"{name}: worst |foot_gpu - foot_reference| = {worst:.2e} m over {T - N} ticks")
