"""Generate tests/golden/gait_golden.npz by importing the reference's OWN planner and
swing-trajectory generator (/root/reference/src/footstep_planner.py,
foot_trajectory_generator.py) with casadi/dartpy/matplotlib stubbed out, and recording,
for several gaits: the plan, the contact mask at every tick (get_phase_at_time), the
step index, and the foot position MPC.update_r_num would use (reference
src/mpc.py:306-318).  Run in the build container only; the npz travels.
"""
import sys
import types
import numpy as np

for name in ("casadi", "dartpy", "matplotlib", "matplotlib.pyplot"):
    m = types.ModuleType(name)
    sys.modules[name] = m
sys.modules["casadi"].MX = sys.modules["casadi"].DM = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference/src")
from footstep_planner import FootstepPlanner            # noqa: E402
from foot_trajectory_generator import FootTrajectoryGenerator   # noqa: E402

LEGS = ["FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT"]
gold = np.load("tests/golden/simulation_log_golden.npz")
feet0 = gold["feet"][0]

CASES = {
    # name: (first_swing, ss, ds, v_ref, theta_dot, total_steps, yaw0)
    "trot": ([1, 0, 0, 1], 10, 10, [0.08, 0.0, 0.0], 0.0, 20, 0.0),
    "pseudo_gallop": ([0, 0, 1, 1], 10, 5, [0.18, 0.0, 0.0], 0.0, 20, 0.0),
    "pseudo_gallop_ds4": ([0, 0, 1, 1], 10, 4, [0.18, 0.0, 0.0], 0.0, 20, 0.0),
    "amble": ([1, 0, 1, 0], 10, 10, [0.08, 0.0, 0.0], 0.0, 20, 0.0),
    "pronk": ([0, 0, 0, 0], 10, 8, [0.10, 0.0, 0.0], 0.0, 20, 0.0),
    "trot_turning": ([1, 0, 0, 1], 10, 10, [0.08, 0.03, 0.0], 0.3, 12, 0.2),
    "trot_ss7": ([1, 0, 0, 1], 7, 3, [0.2, -0.05, 0.0], -0.2, 9, -0.4),
    "stand": ([1, 0, 0, 1], 10, 10, [0.0, 0.0, 0.0], 0.0, 0, 0.0),
}
out = {}
for name, (fs, ss, ds, v, om, steps, yaw) in CASES.items():
    params = {"g": -9.81, "h": 0.285, "step_height": 0.08, "ss_duration": ss, "ds_duration": ds,
              "world_time_step": 0.01, "total_steps": steps, "first_swing": np.array(fs),
              "µ": 1, "N": 10, "v_com_ref": np.array(v), "theta_dot": om}
    initial = {leg: feet0[l].copy() for l, leg in enumerate(LEGS)}
    initial.update(yaw=yaw, roll=0.0, pitch=0.0, com_position=np.array([0., 0., 0.285]))
    planner = FootstepPlanner(initial_configuration=initial, params=params, show=False)
    gen = FootTrajectoryGenerator(footstep_planner=planner, params=params)
    S = len(planner.plan)
    T = (min(S, 24) + 2) * (ss + ds)
    pos = np.array([[np.asarray(s["pos"][leg], dtype=float) for leg in LEGS] for s in planner.plan])
    feet_id = np.array([np.asarray(s["feet_id"]) for s in planner.plan])
    mask = np.zeros((T, 4), dtype=np.int64)
    step = np.zeros(T, dtype=np.int64)
    foot = np.zeros((T, 4, 3))
    for t in range(T):
        phase = planner.get_phase_at_time(t)
        mask[t] = phase
        step[t] = planner.get_step_index_at_time(t)
        for l, leg in enumerate(LEGS):
            if planner.is_swing(leg, phase) == 1:
                foot[t, l] = gen.generate_feet_trajectories_at_time(t, leg)["pos"][3:]
            else:
                foot[t, l] = planner.plan[step[t]]["pos"][leg]
    for k, val in dict(first_swing=np.array(fs), ss=ss, ds=ds, v=np.array(v), theta_dot=om,
                       total_steps=steps, yaw=yaw, pos=pos, feet_id=feet_id, mask=mask,
                       step=step, foot=foot, feet0=feet0).items():
        out[f"{name}/{k}"] = val
np.savez_compressed("tests/golden/gait_golden.npz", **out)
print("wrote tests/golden/gait_golden.npz", len(out), "arrays")
