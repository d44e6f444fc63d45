"""Oracle of the leg controllers (oracle/leg_controller.py, reference src/main.py:130-282 and
src/foot_trajectory_generator.py:27-96) against fixtures recorded from the reference's own
generator (scripts/make_golden_swing.py) and against the desired foot positions the reference's
committed run logged (simulation_log.pkl 'FEET POS' desired -> simulation_log_golden.npz)."""
import numpy as np
import pytest

import mpc_b200 as pkg
from oracle import leg_controller as lc

CASES = ["pseudo_gallop", "trot", "pronk", "trot_turning", "trot_ss7"]


@pytest.mark.parametrize("name", CASES)
def test_controller_queries_match_the_reference_generator(swing_gold, name):
    g = lambda k: swing_gold[f"{name}/{k}"]
    T = g("p_des").shape[0]
    for t in range(T):
        st, p, v, a = lc.controller_query(g("pos"), g("feet_id"), int(g("ss")), int(g("ds")),
                                          float(g("step_height")), 0.01, t)
        assert np.array_equal(st, g("gait_ctrl")[t]), t            # bit-exact stance selection
        sw = st == 0
        np.testing.assert_allclose(p, g("p_des")[t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(v[sw], g("v_des")[t][sw], rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(a[sw], g("a_des")[t][sw], rtol=1e-13, atol=1e-10)


def test_desired_feet_of_the_committed_run(gold, swing_gold):
    """p_des of every leg at every tick of the reference's logged run (1000 ticks)."""
    params = {"ss_duration": int(gold["ss_duration"]), "ds_duration": int(gold["ds_duration"]),
              "v_com_ref": gold["v_com_ref"], "theta_dot": float(gold["theta_dot"]),
              "total_steps": int(gold["total_steps"]), "first_swing": gold["first_swing"],
              "world_time_step": 0.01, "step_height": float(gold["step_height"])}
    initial = {leg: gold["feet"][0][l].copy() for l, leg in enumerate(pkg.LEGS)}
    initial["yaw"] = 0.0
    plan = pkg.GaitPlan.from_initial(initial, params)
    for t in range(1000):
        _, p, _, _ = lc.controller_query(plan.pos, plan.feet_id, plan.ss, plan.ds, plan.step_height, 0.01, t)
        np.testing.assert_allclose(p, gold["feet_des"][t], rtol=0, atol=1e-12, err_msg=str(t))


def test_torque_laws_small_case():
    """Hand-checkable case: identity Jacobian -> stance tau = -f; swing tau = Kp dp + Kd dv + M_diag a + CG."""
    I = np.tile(np.eye(3), (4, 1, 1))
    f = np.arange(12.0).reshape(4, 3)
    M = np.tile(np.diag([2.0, 3.0, 4.0]) + 0.5, (4, 1, 1))          # off-diagonals drop out (elementwise with I)
    cg = np.full((4, 3), 0.1)
    z = np.zeros((4, 3))
    p_des, v_des, a_des = z + 0.01, z + 0.2, z + 1.0
    tau = lc.leg_torques(np.array([1, 0, 1, 0]), f, I, 0 * I, M, cg, z, z, z, p_des, v_des, a_des)
    np.testing.assert_allclose(tau[0], -f[0])
    np.testing.assert_allclose(tau[2], -f[2])
    np.testing.assert_allclose(tau[1], 250 * 0.01 + 15 * 0.2 + np.array([2.5, 3.5, 4.5]) + 0.1)
