"""Lite3 leg kinematics (reference src/main.py:203-214, 236-262, 286-350 reads
them from DART): the closed forms of the product (kinematics.py, mirrored by the CUDA kernel) against
(a) the reference's logged run - feet and centre of mass of tick 0 from the initial configuration of
src/main.py:67-81 - and (b) the independent numerical restatement oracle/leg_kinematics.py."""
import numpy as np

import mpc_b200 as pkg
from mpc_b200 import kinematics as kin
from oracle import leg_kinematics as ok


def test_tick0_of_the_logged_run(gold):
    q = np.tile(kin.Q_INIT, (4, 1))
    base = np.array([0.0, 0.0, kin.BASE_Z_INIT])
    out = kin.leg_kinematics(base, np.zeros(3), np.zeros(3), np.zeros(3), q, np.zeros((4, 3)))
    assert np.abs(out["foot_pos"] - gold["feet"][0]).max() < 5e-9          # logged to ~8 digits
    assert np.abs(ok.foot_position(base, np.zeros(3), q) - gold["feet"][0]).max() < 5e-9
    com = kin.center_of_mass(base, np.zeros(3), q)
    assert np.abs(com - gold["state"][0][3:6]).max() < 5e-9                # 'com' of retrieve_state
    com_o, m = ok.center_of_mass(base, np.zeros(3), q)
    assert np.abs(com_o - com).max() < 1e-14 and abs(m - kin.TOTAL_MASS) < 1e-12


def test_closed_forms_match_numerical_restatement():
    rng = np.random.default_rng(0)
    for _ in range(20):
        base = rng.normal(0, 0.5, 3)
        theta = rng.normal(0, 0.4, 3)
        v, w = rng.normal(0, 0.5, 3), rng.normal(0, 1.0, 3)
        q = np.stack([rng.uniform(-0.4, 0.4, 4), rng.uniform(-2.0, 0.3, 4), rng.uniform(0.6, 2.7, 4)], 1)
        dq = rng.normal(0, 2.0, (4, 3))
        out = kin.leg_kinematics(base, theta, v, w, q, dq)
        assert np.abs(out["foot_pos"] - ok.foot_position(base, theta, q)).max() < 1e-13
        assert np.abs(out["J"] - ok.numeric_jacobian(base, theta, q)).max() < 1e-8
        vel, Jd = ok.numeric_rates(base, theta, v, w, q, dq)
        assert np.abs(out["foot_vel"] - vel).max() < 1e-7
        assert np.abs(out["Jdot"] - Jd).max() < 2e-5
        assert np.abs(out["Mleg"] - ok.numeric_mass_rows(base, theta, q)).max() < 1e-8
        # gravity torque = -d(potential)/dq: m g z of the leg's links
        assert np.abs(out["cg"] + np.einsum("lik,i->lk", out["Mleg"], kin.GRAVITY)).max() < 1e-12


def test_inverse_kinematics_round_trip():
    rng = np.random.default_rng(1)
    for _ in range(50):
        l = int(rng.integers(0, 4))
        q = np.array([rng.uniform(-0.42, 0.42), rng.uniform(-1.5, 0.0), rng.uniform(0.6, 2.0)])   # foot below the hip
        _, _, _, org = kin.leg_frames(l, q)
        assert np.abs(kin.leg_ik(l, org[3]) - q).max() < 1e-9
