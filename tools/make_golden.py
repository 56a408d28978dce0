"""Generate tests/golden/*.npz from the REFERENCE's own NumPy twins (run in the build container only).

    python tools/make_golden.py [/root/reference] [--only reference_motion|rewards|rewards_standing|rewards_library]

Imports, unmodified, from the reference checkout:
  playground/common/rewards_numpy.py                     (twin of common/rewards.py)
  playground/open_duck_mini_v2/custom_rewards_numpy.py   (twin of open_duck_mini_v2/custom_rewards.py)
  playground/common/poly_reference_motion_numpy.py       (twin of common/poly_reference_motion.py) + the pickle
and records seeded inputs/outputs.  /root/reference does not exist on the GPU box, so the vectors are committed.
"""
import contextlib
import io
import os
import sys

import numpy as np

args = [a for a in sys.argv[1:] if not a.startswith("--")]
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None      # rewrite one file, leave the others untouched
if only:
    args = [a for a in args if a != only]
ref = args[0] if args else "/root/reference"
sys.path.insert(0, ref)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(out, exist_ok=True)
from playground.common import rewards_numpy as R  # noqa: E402
from playground.common.poly_reference_motion_numpy import PolyReferenceMotion  # noqa: E402
from playground.open_duck_mini_v2.custom_rewards_numpy import reward_imitation  # noqa: E402

rng = np.random.default_rng(20250925)
with contextlib.redirect_stdout(io.StringIO()):
    prm = PolyReferenceMotion(os.path.join(ref, "playground/open_duck_mini_v2/data/polynomial_coefficients.pkl"))

# ---- reference motion: grid points, in-between values, out-of-range values, every phase index
K = 160
dx = rng.uniform(-0.25, 0.3, K); dy = rng.uniform(-0.2, 0.2, K); dth = rng.uniform(-1.5, 1.5, K)
dx[:6], dy[:6], dth[:6] = [0.0, 0.1, -0.148, 0.222, 0.074, 0.0], [0.0, 0.0, -0.111, 0.111, 0.037, -0.05], [0.0, 0.3, -1.111, 1.222, 0.0, -0.1]
ii = rng.integers(0, 60, K); ii[:27] = np.arange(27)
vals = np.array([prm.get_reference_motion(dx[k], dy[k], dth[k], int(ii[k])) for k in range(K)], dtype=np.float64)
idx = np.array([prm.vel_to_index(dx[k], dy[k], dth[k]) for k in range(K)])
save = lambda name, **kw: np.savez_compressed(os.path.join(out, name + ".npz"), **kw) if only in (None, name) else None   # noqa: E731
save("reference_motion", dx=dx, dy=dy, dtheta=dth, i=ii, ref=vals, index=idx,
                    nb_steps_in_period=prm.nb_steps_in_period, dxs=prm.dxs, dys=prm.dys, dthetas=prm.dthetas)

# ---- rewards: the 6 library terms Joystick uses + imitation (joystick.py:634-667); default_pose = keyframe ctrl
default_pose = np.array([0.002, 0.053, -0.63, 1.368, -0.784, 0, 0, 0, 0, -0.003, -0.065, 0.635, 1.379, -0.796])
K = 256
cmd = rng.uniform(-1, 1, (K, 7)) * np.array([0.15, 0.2, 1.0, 1.1, 0.78, 1.5, 0.5])
cmd[::7] = 0.0                      # zero commands (stand_still / imitation gating)
cmd[1::11, :3] *= 0.02              # norms around the 0.01 threshold
linvel = rng.normal(0, 0.15, (K, 3)); gyro = rng.normal(0, 0.6, (K, 3))
force = rng.uniform(-3.23, 3.23, (K, 14)); act = rng.uniform(-1, 1, (K, 14)); last = rng.uniform(-1, 1, (K, 14))
base_qpos = np.concatenate([rng.normal(0, 0.1, (K, 3)), rng.normal(0, 1, (K, 4))], axis=1)
base_qvel = rng.normal(0, 0.3, (K, 6))
q = default_pose + rng.normal(0, 0.3, (K, 14)); qd = rng.normal(0, 2.0, (K, 14))
contact = rng.integers(0, 2, (K, 2)).astype(bool)
refm = np.array([prm.get_reference_motion(rng.uniform(-0.15, 0.22), rng.uniform(-0.11, 0.11), rng.uniform(-1, 1), int(rng.integers(0, 27))) for _ in range(K)])
sigma = 0.01
terms = np.zeros((K, 7))
for k in range(K):
    terms[k] = [R.reward_tracking_lin_vel(cmd[k], linvel[k], sigma), R.reward_tracking_ang_vel(cmd[k], gyro[k], sigma),
                R.cost_torques(force[k]), R.cost_action_rate(act[k], last[k]),
                R.cost_stand_still(cmd[k], q[k], qd[k], default_pose, ignore_head=False), R.reward_alive(),
                reward_imitation(base_qpos[k], base_qvel[k], q[k], qd[k], contact[k], refm[k], cmd[k], True)]
save("rewards", command=cmd, local_linvel=linvel, gyro=gyro, actuator_force=force, action=act,
                    last_act=last, base_qpos=base_qpos, base_qvel=base_qvel, q=q, qd=qd, contact=contact, ref=refm, terms=terms,
                    tracking_sigma=sigma, default_pose=default_pose)

# ---- Standing task (open_duck_mini_v2/standing.py:573-606): orientation, torques, action_rate, alive, stand_still(ignore_head=True), head_pos
K = 256
cmd_s = rng.uniform(-1, 1, (K, 7)) * np.array([0.15, 0.2, 1.0, 1.1, 0.78, 2.7, 0.5])
cmd_s[::2, :3] = 0.0                # the Standing env itself always commands zero velocity (standing.py:648-655)
cmd_s[1::11, :3] *= 0.02
up = rng.normal(0, 0.3, (K, 3)); up[:, 2] = 1.0; up /= np.linalg.norm(up, axis=1, keepdims=True)
force_s = rng.uniform(-3.23, 3.23, (K, 14)); act_s = rng.uniform(-1, 1, (K, 14)); last_s = rng.uniform(-1, 1, (K, 14))
q_s = default_pose + rng.normal(0, 0.3, (K, 14)); qd_s = rng.normal(0, 2.0, (K, 14))
terms_s = np.zeros((K, 6))
for k in range(K):
    terms_s[k] = [R.cost_orientation(up[k]), R.cost_torques(force_s[k]), R.cost_action_rate(act_s[k], last_s[k]), R.reward_alive(),
                  R.cost_stand_still(cmd_s[k], q_s[k], qd_s[k], default_pose, True), R.cost_head_pos(q_s[k], qd_s[k], cmd_s[k])]
save("rewards_standing", command=cmd_s, upvector=up, actuator_force=force_s, action=act_s, last_act=last_s,
                    q=q_s, qd=qd_s, terms=terms_s, default_pose=default_pose)

# ---- the rest of the reward library (common/rewards.py:37-90,120,152-241): terms no shipped env wires in.  Own seed, so adding
# cases here never changes the files above.  Input ranges follow the call sites the signatures document (sensor values, joint
# angles, feet sites); every gate (zero command, |cmd_y| > 0.1, contact flags, clip thresholds) is hit from both sides.
rl = np.random.default_rng(20251017)
K = 256
soft_lo = default_pose - rl.uniform(0.2, 0.6, 14); soft_hi = default_pose + rl.uniform(0.2, 0.6, 14)
hip_idx = np.array([1, 2, 10, 11]); knee_idx = np.array([3, 12])
weights = rl.uniform(0.01, 1.0, 14)
L = dict(
    global_linvel=rl.normal(0, 0.3, (K, 3)), global_angvel=rl.normal(0, 0.8, (K, 3)), base_height=rl.uniform(0.05, 0.25, K),
    base_height_target=np.full(K, 0.15), base_y_speed=rl.normal(0, 0.2, K), freq=rl.uniform(0.5, 3.0, K), amplitude=rl.uniform(0, 0.3, K),
    t=rl.uniform(0, 20, K), tracking_sigma=np.full(K, 0.01), qvel=rl.normal(0, 2.0, (K, 14)), qfrc_actuator=rl.uniform(-3.23, 3.23, (K, 14)),
    qpos=default_pose + rl.normal(0, 0.45, (K, 14)), done=rl.integers(0, 2, K).astype(np.float64),
    command=rl.uniform(-1, 1, (K, 7)) * np.array([0.15, 0.2, 1.0, 1.1, 0.78, 1.5, 0.5]),
    contact=rl.integers(0, 2, (K, 2)).astype(np.float64), feet_vel=rl.normal(0, 0.4, (K, 2, 3)),
    foot_pos=np.concatenate([rl.normal(0, 0.1, (K, 2, 2)), rl.uniform(0.0, 0.08, (K, 2, 1))], axis=2), max_foot_height=np.full(K, 0.03),
    swing_peak=rl.uniform(0, 0.06, (K, 2)), first_contact=rl.integers(0, 2, (K, 2)).astype(np.float64), air_time=rl.uniform(0, 0.9, (K, 2)),
    threshold_min=np.full(K, 0.1), threshold_max=np.full(K, 0.5), rz=rl.uniform(0, 0.05, (K, 2)))
L["command"][::7] = 0.0
L["command"][1::11, :3] *= 0.02
L["command"][2::5, 1] *= 0.3                         # |cmd_y| on both sides of the 0.1 hip gate
L["threshold_min"][::3] = 0.2                        # the "# 0.2" alternative of rewards.py:216
lib_terms = np.zeros((K, 15))
for k in range(K):
    g = {n: v[k] for n, v in L.items()}
    lib_terms[k] = [
        R.cost_lin_vel_z(g["global_linvel"]), R.cost_ang_vel_xy(g["global_angvel"]), R.cost_base_height(g["base_height"], g["base_height_target"]),
        R.reward_base_y_swing(g["base_y_speed"], g["freq"], g["amplitude"], g["t"], g["tracking_sigma"]),
        R.cost_energy(g["qvel"], g["qfrc_actuator"]), R.cost_joint_pos_limits(g["qpos"], soft_lo, soft_hi), R.cost_termination(g["done"]),
        R.cost_joint_deviation_hip(g["qpos"], g["command"], hip_idx, default_pose), R.cost_joint_deviation_knee(g["qpos"], knee_idx, default_pose),
        R.cost_pose(g["qpos"], default_pose, weights), R.cost_feet_slip(g["contact"], g["global_linvel"]),
        R.cost_feet_clearance(g["feet_vel"], g["foot_pos"], g["max_foot_height"]), R.cost_feet_height(g["swing_peak"], g["first_contact"], g["max_foot_height"]),
        R.reward_feet_air_time(g["air_time"], g["first_contact"], g["command"], g["threshold_min"], g["threshold_max"]),
        R.reward_feet_phase(g["foot_pos"], g["rz"])]
save("rewards_library", terms=lib_terms, soft_lowers=soft_lo, soft_uppers=soft_hi, hip_indices=hip_idx, knee_indices=knee_idx, weights=weights,
     default_pose=default_pose, **L)
print("golden written:", sorted(os.listdir(out)))
