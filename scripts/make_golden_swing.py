"""Generate tests/golden/swing_golden.npz from the reference's OWN planner and swing-trajectory
generator (/root/reference/src/footstep_planner.py, foot_trajectory_generator.py; casadi /
dartpy / matplotlib stubbed), replaying the controller-side queries of the reference's
`customPreStep` tick by tick (src/main.py:130-166, 225-231):

    gait = plan[step_index]['feet_id']            (read once per tick, before the leg loop)
    stance leg: p_des = plan[step_index]['pos'][leg]
    swing  leg: generate_feet_trajectories_at_time(t, leg) -> p_des, v_des, a_des,
                z clamp `if p_des[2] < 0: p_des[2] = 0; v_des[2] = 0`

including the generator's side effect on plan[step]['feet_id'] (foot_trajectory_generator.py:54).
Run in the build container only; the npz travels.
"""
import sys
import types
import numpy as np

for name in ("casadi", "dartpy", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["casadi"].MX = sys.modules["casadi"].DM = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference/src")
from footstep_planner import FootstepPlanner            # noqa: E402
from foot_trajectory_generator import FootTrajectoryGenerator   # noqa: E402

LEGS = ["FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT"]
gold = np.load("tests/golden/simulation_log_golden.npz")
feet0 = gold["feet"][0]

CASES = {
    # name: (first_swing, ss, ds, v_ref, theta_dot, total_steps, yaw0, step_height)
    "pseudo_gallop": ([0, 0, 1, 1], 10, 5, [0.18, 0.0, 0.0], 0.0, 20, 0.0, 0.08),   # the committed run
    "trot": ([1, 0, 0, 1], 10, 10, [0.08, 0.0, 0.0], 0.0, 20, 0.0, 0.08),
    "pronk": ([0, 0, 0, 0], 10, 8, [0.10, 0.0, 0.0], 0.0, 20, 0.0, 0.08),
    "trot_turning": ([1, 0, 0, 1], 10, 10, [0.08, 0.03, 0.0], 0.3, 12, 0.2, 0.05),
    "trot_ss7": ([1, 0, 0, 1], 7, 3, [0.2, -0.05, 0.0], -0.2, 9, -0.4, 0.08),
}
out = {}
for name, (fs, ss, ds, v, om, steps, yaw, sh) in CASES.items():
    params = {"g": -9.81, "h": 0.285, "step_height": sh, "ss_duration": ss, "ds_duration": ds,
              "world_time_step": 0.01, "total_steps": steps, "first_swing": np.array(fs),
              "µ": 1, "N": 10, "v_com_ref": np.array(v), "theta_dot": om}
    initial = {leg: feet0[l].copy() for l, leg in enumerate(LEGS)}
    initial.update(yaw=yaw, roll=0.0, pitch=0.0, com_position=np.array([0., 0., 0.285]))
    planner = FootstepPlanner(initial_configuration=initial, params=params, show=False)
    gen = FootTrajectoryGenerator(footstep_planner=planner, params=params)
    S = len(planner.plan)
    T = (S + 2) * (ss + ds)
    pos = np.array([[np.asarray(s["pos"][leg], dtype=float) for leg in LEGS] for s in planner.plan])
    feet_id = np.array([np.asarray(s["feet_id"]) for s in planner.plan])     # before any side effect
    gait_ctrl = np.zeros((T, 4), dtype=np.int64)
    p_des = np.zeros((T, 4, 3)); v_des = np.zeros((T, 4, 3)); a_des = np.zeros((T, 4, 3))
    for t in range(T):
        step_index = planner.get_step_index_at_time(t)
        gait = planner.plan[step_index]["feet_id"]            # src/main.py:152-153
        gait_ctrl[t] = np.asarray(gait)
        for j, leg in enumerate(LEGS):
            if gait[j] == 1:
                p_des[t, j] = planner.plan[step_index]["pos"][leg]          # src/main.py:159
            else:
                sd = gen.generate_feet_trajectories_at_time(t, leg)         # src/main.py:225
                p, vv, a = sd["pos"][3:].copy(), sd["vel"][3:].copy(), sd["acc"][3:].copy()
                if p[2] < 0:                                                # src/main.py:229-231
                    p[2] = 0
                    vv[2] = 0
                p_des[t, j], v_des[t, j], a_des[t, j] = p, vv, a
    for k, val in dict(first_swing=np.array(fs), ss=ss, ds=ds, v=np.array(v), theta_dot=om,
                       total_steps=steps, yaw=yaw, step_height=sh, pos=pos, feet_id=feet_id,
                       gait_ctrl=gait_ctrl, p_des=p_des, v_des=v_des, a_des=a_des, feet0=feet0).items():
        out[f"{name}/{k}"] = val
np.savez_compressed("tests/golden/swing_golden.npz", **out)
print("wrote tests/golden/swing_golden.npz", len(out), "arrays")
