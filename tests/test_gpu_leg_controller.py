"""cmpc_leg_torques (CUDA, through the C ABI) against the fp64 oracle of the reference's
ground / swing leg controllers (src/main.py:193-282) on the gaits of the golden fixtures."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mpc_b200 as pkg                                  # noqa: E402
from mpc_b200 import _capi                              # noqa: E402
from oracle import leg_controller as lc                 # noqa: E402

CASES = ["pseudo_gallop", "trot", "pronk", "trot_turning", "trot_ss7"]


def _tables(g, names, dev, step_height):
    """Stack the plans of several golden cases (same number of steps) into one batch of robots."""
    S = g[f"{names[0]}/pos"].shape[0]
    assert all(g[f"{n}/pos"].shape[0] == S for n in names)
    pos = np.stack([g[f"{n}/pos"] for n in names]).astype(np.float32)
    fid = np.stack([g[f"{n}/feet_id"] for n in names])
    bits = (fid * np.array([1, 2, 4, 8])).sum(-1).astype(np.uint8)
    B = len(names)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    keep = dict(pos=t(pos), bits=t(bits), ss=t(np.array([int(g[f"{n}/ss"]) for n in names], np.int32)),
                ds=t(np.array([int(g[f"{n}/ds"]) for n in names], np.int32)),
                v=torch.zeros((B, 3), device=dev), om=torch.zeros(B, device=dev), rp=torch.zeros((B, 2), device=dev))
    gt = _capi.GaitTables(plan_pos=keep["pos"].data_ptr(), feet_id=keep["bits"].data_ptr(), ss=keep["ss"].data_ptr(),
                          ds=keep["ds"].data_ptr(), v_ref=keep["v"].data_ptr(), omega_ref=keep["om"].data_ptr(),
                          rp0=keep["rp"].data_ptr(), S=S, total_steps=S, step_height=step_height, g=-9.81)
    return gt, keep


@pytest.mark.parametrize("group", [("pseudo_gallop", "trot", "pronk"), ("trot_turning",), ("trot_ss7",)])
def test_leg_torques_match_the_oracle(swing_gold, group):
    dev = torch.device("cuda", 0)
    names = list(group) * 5
    sh = float(swing_gold[f"{names[0]}/step_height"])
    B, N = len(names), 10
    gt, keep = _tables(swing_gold, names, dev, sh)
    mpc = pkg.BatchedMPC(N=N, max_batch=B)
    ctl = pkg.BatchedLegController(mpc, gt, kp=(250.0, 240.0, 260.0), kd=15.0)
    rng = np.random.default_rng(5)
    tick = torch.zeros(1, dtype=torch.int32, device=dev)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    for t in list(range(0, 64)) + [95, 170, 171, 399, 400, 5000]:
        U = rng.normal(0, 30, (B, N, 12)); J = rng.normal(0, 0.3, (B, 4, 3, 3)); Jd = rng.normal(0, 0.5, (B, 4, 3, 3))
        M = rng.normal(0, 0.2, (B, 4, 3, 3)); cg = rng.normal(0, 1, (B, 4, 3)); dq = rng.normal(0, 2, (B, 4, 3))
        fp = rng.normal(0, 0.2, (B, 4, 3)); fv = rng.normal(0, 0.5, (B, 4, 3))
        args = [f32(a) for a in (U, J, Jd, M, cg, dq, fp, fv)]
        tick.fill_(t)
        tau, p_des, stance = ctl.torques(tick, *args)
        torch.cuda.synchronize()
        tau, p_des, stance = tau.cpu().numpy(), p_des.cpu().numpy(), stance.cpu().numpy()
        h = [a.cpu().numpy().astype(np.float64) for a in args]       # the fp32-rounded inputs
        for b, n in enumerate(names):
            g = lambda k: swing_gold[f"{n}/{k}"]
            st, p, v, a = lc.controller_query(g("pos").astype(np.float32), g("feet_id"), int(g("ss")), int(g("ds")),
                                              sh, 0.01, t)
            assert int(stance[b]) == int((st * np.array([1, 2, 4, 8])).sum()), (n, t)      # bit-exact
            np.testing.assert_allclose(p_des[b], p, rtol=0, atol=2e-6, err_msg=f"{n} {t}")
            ref = lc.leg_torques(st, h[0][b, 0].reshape(4, 3), h[1][b], h[2][b], h[3][b], h[4][b], h[5][b],
                                 h[6][b], h[7][b], p, v, a, kp=(250.0, 240.0, 260.0), kd=15.0)
            scale = np.abs(ref).max() + 1.0
            assert np.abs(tau[b] - ref).max() <= 2e-5 * scale, (n, t, np.abs(tau[b] - ref).max(), scale)


def test_leg_torques_api_errors():
    dev = torch.device("cuda", 0)
    mpc = pkg.BatchedMPC(N=10, max_batch=4)
    z = torch.zeros((2, 4, 3), device=dev)
    gt = _capi.GaitTables()                       # null tables
    ctl = pkg.BatchedLegController(mpc, gt)
    with pytest.raises(pkg.CmpcError):
        ctl.torques(torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros((2, 10, 12), device=dev),
                    torch.zeros((2, 4, 3, 3), device=dev), torch.zeros((2, 4, 3, 3), device=dev),
                    torch.zeros((2, 4, 3, 3), device=dev), z, z, z, z)
    with pytest.raises(ValueError):
        ctl.torques(torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros((2, 10, 11), device=dev),
                    torch.zeros((2, 4, 3, 3), device=dev), torch.zeros((2, 4, 3, 3), device=dev),
                    torch.zeros((2, 4, 3, 3), device=dev), z, z, z, z)


def test_leg_kinematics_kernel_matches_numerical_restatement(gold):
    """cmpc_leg_kinematics against oracle/leg_kinematics.py (forward kinematics from the URDF joint
    tree, Jacobians by finite differences) and against the logged feet of the reference's tick 0."""
    from oracle import leg_kinematics as ok
    dev = torch.device("cuda", 0)
    B = 64
    rng = np.random.default_rng(2)
    base = rng.normal(0, 0.5, (B, 3)); theta = rng.normal(0, 0.4, (B, 3))
    v, w = rng.normal(0, 0.5, (B, 3)), rng.normal(0, 1.0, (B, 3))
    q = np.stack([rng.uniform(-0.4, 0.4, (B, 4)), rng.uniform(-2.0, 0.3, (B, 4)), rng.uniform(0.6, 2.7, (B, 4))], -1)
    dq = rng.normal(0, 2.0, (B, 4, 3))
    # robot 0 = the reference's initial configuration (src/main.py:67-81)
    base[0], theta[0], v[0], w[0] = (0, 0, 0.299), 0, 0, 0
    q[0], dq[0] = np.deg2rad([0.0, -60.0, 90.0]), 0
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    mpc = pkg.BatchedMPC(N=10, max_batch=B)
    ctl = pkg.BatchedLegController(mpc, _capi.GaitTables())
    args = [f32(a) for a in (base, theta, v, w, q, dq)]
    out = {k: t.cpu().numpy().astype(np.float64) for k, t in ctl.kinematics(*args).items()}
    assert np.abs(out["foot_pos"][0] - gold["feet"][0]).max() < 2e-7           # the reference's logged feet
    h = [a.cpu().numpy().astype(np.float64) for a in args]                       # fp32-rounded inputs
    for b in range(0, B, 4):
        fk = ok.foot_position(h[0][b], h[1][b], h[4][b])
        assert np.abs(out["foot_pos"][b] - fk).max() < 2e-6
        assert np.abs(out["J"][b] - ok.numeric_jacobian(h[0][b], h[1][b], h[4][b])).max() < 2e-6
        vel, Jd = ok.numeric_rates(h[0][b], h[1][b], h[2][b], h[3][b], h[4][b], h[5][b])
        assert np.abs(out["foot_vel"][b] - vel).max() < 2e-5
        assert np.abs(out["Jdot"][b] - Jd).max() < 1e-4
        Mr = ok.numeric_mass_rows(h[0][b], h[1][b], h[4][b])
        assert np.abs(out["Mleg"][b] - Mr).max() < 1e-6
        assert np.abs(out["cg"][b] - 9.81 * Mr[:, 2, :]).max() < 1e-5


def test_self_contained_joint_torques(swing_gold):
    """kinematics -> torques entirely on the device (no simulator inputs), against the fp64
    restatement of src/main.py:203-282 fed with the host kinematics model."""
    from mpc_b200 import kinematics as kin
    dev = torch.device("cuda", 0)
    names = ["pseudo_gallop", "trot", "pronk"] * 4
    sh = float(swing_gold[f"{names[0]}/step_height"])
    B, N = len(names), 10
    gt, keep = _tables(swing_gold, names, dev, sh)
    mpc = pkg.BatchedMPC(N=N, max_batch=B)
    ctl = pkg.BatchedLegController(mpc, gt)
    rng = np.random.default_rng(9)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    tick = torch.zeros(1, dtype=torch.int32, device=dev)
    for t in (0, 16, 18, 23, 40, 171):
        base = np.tile([0.0, 0.0, 0.29], (B, 1)) + rng.normal(0, 0.01, (B, 3))
        theta = rng.normal(0, 0.05, (B, 3)); v = rng.normal(0, 0.1, (B, 3)); w = rng.normal(0, 0.2, (B, 3))
        q = np.tile(kin.Q_INIT, (B, 4, 1)) + rng.normal(0, 0.1, (B, 4, 3)); dq = rng.normal(0, 1.0, (B, 4, 3))
        U = rng.normal(0, 30, (B, N, 12))
        a = [f32(x) for x in (base, theta, v, w, q, dq)]
        k = ctl.kinematics(*a)
        tick.fill_(t)
        Ud = f32(U)
        tau, p_des, stance = ctl.torques(tick, Ud, k["J"], k["Jdot"], k["Mleg"], k["cg"], a[5], k["foot_pos"], k["foot_vel"])
        torch.cuda.synchronize()
        tau = tau.cpu().numpy()
        h = [x.cpu().numpy().astype(np.float64) for x in a]
        for b, n in enumerate(names):
            g = lambda key: swing_gold[f"{n}/{key}"]
            st, p, vv, aa = lc.controller_query(g("pos").astype(np.float32), g("feet_id"), int(g("ss")), int(g("ds")),
                                                sh, 0.01, t)
            m = kin.leg_kinematics(h[0][b], h[1][b], h[2][b], h[3][b], h[4][b], h[5][b])
            ref = lc.leg_torques(st, Ud[b, 0].cpu().numpy().astype(np.float64).reshape(4, 3), m["J"], m["Jdot"],
                                 m["Mleg"], m["cg"], h[5][b], m["foot_pos"], m["foot_vel"], p, vv, aa)
            assert np.abs(tau[b] - ref).max() <= 5e-4 * (np.abs(ref).max() + 1.0), (n, t)
