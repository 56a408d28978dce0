"""Generate tests/golden/*.npz from the REFERENCE's own NumPy twins (run in the build container only).

    python tools/make_golden.py [/root/reference]

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

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
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
np.savez_compressed(os.path.join(out, "reference_motion.npz"), dx=dx, dy=dy, dtheta=dth, i=ii, ref=vals, index=idx,
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
np.savez_compressed(os.path.join(out, "rewards.npz"), command=cmd, local_linvel=linvel, gyro=gyro, actuator_force=force, action=act,
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
np.savez_compressed(os.path.join(out, "rewards_standing.npz"), command=cmd_s, upvector=up, actuator_force=force_s, action=act_s, last_act=last_s,
                    q=q_s, qd=qd_s, terms=terms_s, default_pose=default_pose)
print("golden written:", sorted(os.listdir(out)))
