"""the batched gait plan that feeds the device assembly
(problems.BatchedGaitPlan) against fixtures recorded from the reference's OWN planner and swing
generator (scripts/make_golden_gait.py -> tests/golden/gait_golden.npz; reference
src/footstep_planner.py:72-177, 226-256, src/foot_trajectory_generator.py:27-96).
Contact masks and step indices bit-exact; planned positions agree bit for bit in this image (the
assertion allows 2 ulp of fp64 for a BLAS with a different 2x2 kernel)."""
import numpy as np
import pytest

import mpc_b200 as pkg
from mpc_b200.problems import BatchedGaitPlan

CASES = ["trot", "pseudo_gallop", "pseudo_gallop_ds4", "amble", "pronk", "trot_turning", "trot_ss7"]


def batched_plan(gg, names):
    g = lambda n, k: gg[f"{n}/{k}"]
    steps = {int(g(n, "total_steps")) for n in names}
    assert len(steps) == 1
    return BatchedGaitPlan.build(
        np.stack([g(n, "feet0") for n in names]), np.array([float(g(n, "yaw")) for n in names]),
        np.stack([g(n, "first_swing") for n in names]), np.array([int(g(n, "ss")) for n in names]),
        np.array([int(g(n, "ds")) for n in names]), np.stack([g(n, "v") for n in names]),
        np.array([float(g(n, "theta_dot")) for n in names]), total_steps=steps.pop())


GROUPS = [["trot", "pseudo_gallop", "pseudo_gallop_ds4", "amble", "pronk"], ["trot_turning"], ["trot_ss7"]]


@pytest.mark.parametrize("names", GROUPS, ids=lambda g: "+".join(g))
def test_batched_plan_matches_reference_planner(gait_gold, names):
    """Heterogeneous robots in ONE batch (different first_swing / ss / ds / v / yaw rate)."""
    plan = batched_plan(gait_gold, names)
    for b, n in enumerate(names):
        ref_pos = gait_gold[f"{n}/pos"]
        assert plan.pos[b].shape == ref_pos.shape
        assert np.all(np.abs(plan.pos[b] - ref_pos) <= 2 * np.spacing(np.abs(ref_pos))), n      # bit-exact here
        assert np.array_equal(plan.feet_id[b], gait_gold[f"{n}/feet_id"]), n
        T = gait_gold[f"{n}/mask"].shape[0]
        t = np.arange(T)[None]
        sub = BatchedGaitPlan(plan.pos[b:b + 1], plan.feet_id[b:b + 1], plan.ss[b:b + 1], plan.ds[b:b + 1])
        assert np.array_equal(sub.step_index(t)[0], gait_gold[f"{n}/step"]), n          # bit-exact
        assert np.array_equal(sub.stance_mask(t)[0], gait_gold[f"{n}/mask"]), n         # bit-exact
        foot = sub.foot_position(t)[0]
        assert np.abs(foot - gait_gold[f"{n}/foot"]).max() <= 4e-16, n


def test_batched_plan_equals_single_robot_plan(gait_gold):
    """... and the single-robot GaitPlan (bit-exact against the same fixtures, tests/test_gait.py)."""
    from test_gait import _plan
    for n in CASES:
        one = _plan(gait_gold, n)
        bat = batched_plan(gait_gold, [n])
        assert np.abs(bat.pos[0] - one.pos).max() <= 4e-16
        assert np.array_equal(bat.feet_id[0], one.feet_id)
