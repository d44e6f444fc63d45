"""Extract the golden vectors of the reference's committed run
(/root/reference/src/simulation_log.pkl, written by reference src/logger.py:64-66)
into tests/golden/simulation_log_golden.npz.  Run once in the build container
(the GPU box has no /root/reference).  Schema: SURVEY.md section 4.
"""
import pickle
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/simulation_log.pkl"
dst = sys.argv[2] if len(sys.argv) > 2 else "tests/golden/simulation_log_golden.npz"
log = pickle.load(open(src, "rb"))
legs = ["FL_FOOT", "FR_FOOT", "HL_FOOT", "HR_FOOT"]
state = np.array(log["TRACKING PERFORMANCE"]["actual"], dtype=np.float64)        # (1000,12)
desired = np.array(log["TRACKING PERFORMANCE"]["desired"], dtype=np.float64)     # (1000,12)
feet = np.stack([np.array(log["FEET POS"][l]["actual"], dtype=np.float64) for l in legs], 1)
feet_des = np.stack([np.array(log["FEET POS"][l]["des"], dtype=np.float64) for l in legs], 1)
forces = np.stack([np.stack([np.array(log["FORCES"][l][c], dtype=np.float64) for c in "xyz"], 1)
                   for l in legs], 1).reshape(len(state), 12)                    # (1000,12)
pred = log["MPC PREDICTIONS"]
sp = log["sim_params"]
np.savez_compressed(
    dst, state=state, desired=desired, feet=feet, feet_des=feet_des, forces=forces,
    pred_t=np.array([p["time step"] for p in pred]),
    pred_state=np.stack([p["predicted_state"] for p in pred]),
    pred_desired=np.stack([p["desired_state"] for p in pred]),
    pred_fz=np.stack([p["predicted forces"] for p in pred]),
    mpc_freq=np.float64(log["mpc_freq"]),
    g=sp["g"], h=sp["h"], step_height=sp["step_height"], ss_duration=sp["ss_duration"],
    ds_duration=sp["ds_duration"], world_time_step=sp["world_time_step"],
    total_steps=sp["total_steps"], first_swing=np.asarray(sp["first_swing"]),
    mu=float(sp["µ"]), N=sp["N"], v_com_ref=np.asarray(sp["v_com_ref"], dtype=np.float64),
    theta_dot=sp["theta_dot"])
print("wrote", dst, {k: v.shape for k, v in np.load(dst).items()})
