"""Gait plan / contact schedule / swing look-ahead of the product host code against
fixtures recorded from the reference's own planner (scripts/make_golden_gait.py;
reference src/footstep_planner.py:29-256, src/foot_trajectory_generator.py:27-96).
Masks and step indices must be bit-exact; positions are compared bit-exact too."""
import numpy as np
import pytest

import mpc_b200 as pkg

CASES = ["trot", "pseudo_gallop", "pseudo_gallop_ds4", "amble", "pronk", "trot_turning",
         "trot_ss7", "stand"]


def _plan(gg, name):
    g = lambda k: gg[f"{name}/{k}"]
    params = {"ss_duration": int(g("ss")), "ds_duration": int(g("ds")), "v_com_ref": g("v"),
              "theta_dot": float(g("theta_dot")), "total_steps": int(g("total_steps")),
              "first_swing": g("first_swing"), "world_time_step": 0.01, "step_height": 0.08}
    initial = {leg: g("feet0")[l].copy() for l, leg in enumerate(pkg.LEGS)}
    initial["yaw"] = float(g("yaw"))
    return pkg.GaitPlan.from_initial(initial, params)


@pytest.mark.parametrize("name", CASES)
def test_plan_bit_exact(gait_gold, name):
    plan = _plan(gait_gold, name)
    assert plan.pos.shape == gait_gold[f"{name}/pos"].shape
    assert np.array_equal(plan.pos, gait_gold[f"{name}/pos"])
    assert np.array_equal(plan.feet_id, gait_gold[f"{name}/feet_id"])


@pytest.mark.parametrize("name", CASES)
def test_mask_and_step_bit_exact(gait_gold, name):
    plan = _plan(gait_gold, name)
    T = gait_gold[f"{name}/mask"].shape[0]
    t = np.arange(T)
    assert np.array_equal(plan.step_index(t), gait_gold[f"{name}/step"])
    assert np.array_equal(plan.stance_mask(t), gait_gold[f"{name}/mask"])
    # scalar queries agree with the vectorised ones
    for tt in (0, 9, 10, 19, 20, T - 1):
        assert np.array_equal(plan.stance_mask(tt), gait_gold[f"{name}/mask"][tt])


@pytest.mark.parametrize("name", CASES)
def test_lookahead_foot_positions(gait_gold, name):
    plan = _plan(gait_gold, name)
    T = gait_gold[f"{name}/foot"].shape[0]
    got = plan.foot_position(np.arange(T))
    assert np.array_equal(got, gait_gold[f"{name}/foot"])


def test_stance_bits_layout():
    assert pkg.stance_bits([1, 0, 0, 1]) == 0b1001
    assert pkg.stance_bits([0, 1, 1, 0]) == 0b0110
    m = pkg.stance_bits(np.array([[1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 1, 1]]))
    assert m.dtype == np.uint8 and list(m) == [15, 0, 12]


def test_golden_run_contact_schedule(gold):
    """In the reference's logged run a leg's logged force is ~0 (|f|<0.05 N, OSQP eps)
    exactly when the planner says swing (last table row)."""
    from oracle.replay import params_from_golden, initial_from_golden
    plan = pkg.GaitPlan.from_initial(initial_from_golden(gold), params_from_golden(gold))
    mask = plan.stance_mask(np.arange(1000))
    fz = gold["forces"].reshape(1000, 4, 3)[:, :, 2]
    assert np.all(np.abs(fz[mask == 0]) < 0.05)
    assert np.all(fz[mask == 1] > 1.0)
