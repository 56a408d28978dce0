"""Environment configuration: ``default_config()`` of the reference Joystick task.

Mirrors reference open_duck_mini_v2/joystick.py:49-102 (values) and the tables ``_post_init`` derives
(joystick.py:121-204).  ``ml_collections`` is not available in this image, so :class:`ConfigDict` is a
minimal attribute dictionary with the same access pattern (``cfg.noise_config.scales.gyro``) and
``config_overrides`` support (``{"noise_config.level": 0.0}``).
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Any, Dict, Optional

import numpy as np

from . import capi
from .mjcf import CompiledModel
from .poly_reference_motion import PolyTable

# module-level switches of the reference (joystick.py:45-46)
USE_IMITATION_REWARD = True
USE_MOTOR_SPEED_LIMITS = True

JOINTS_ORDER_NO_HEAD = [  # reference constants.py:65-76
    "left_hip_yaw", "left_hip_roll", "left_hip_pitch", "left_knee", "left_ankle",
    "right_hip_yaw", "right_hip_roll", "right_hip_pitch", "right_knee", "right_ankle",
]


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def update_from_flattened_dict(self, overrides: Dict[str, Any]) -> None:
        for key, val in overrides.items():
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(key)
            node[parts[-1]] = val


def _cd(**kw):
    return ConfigDict(kw)


def default_config() -> ConfigDict:
    return _cd(
        ctrl_dt=0.02,
        sim_dt=0.002,
        episode_length=1000,
        action_repeat=1,
        action_scale=0.25,
        dof_vel_scale=0.05,
        history_len=0,
        soft_joint_pos_limit_factor=0.95,
        max_motor_velocity=5.24,
        noise_config=_cd(
            level=1.0,
            action_min_delay=0,
            action_max_delay=3,
            imu_min_delay=0,
            imu_max_delay=3,
            scales=_cd(hip_pos=0.03, knee_pos=0.05, ankle_pos=0.08, joint_vel=2.5, gravity=0.1, linvel=0.1, gyro=0.1,
                       accelerometer=0.05),
        ),
        reward_config=_cd(
            scales=_cd(tracking_lin_vel=2.5, tracking_ang_vel=6.0, torques=-1.0e-3, action_rate=-0.5, stand_still=-0.2,
                       alive=20.0, imitation=1.0),
            tracking_sigma=0.01,
        ),
        push_config=_cd(enable=True, interval_range=[5.0, 10.0], magnitude_range=[0.1, 1.0]),
        lin_vel_x=[-0.15, 0.15],
        lin_vel_y=[-0.2, 0.2],
        ang_vel_yaw=[-1.0, 1.0],
        neck_pitch_range=[-0.34, 1.1],
        head_pitch_range=[-0.78, 0.78],
        head_yaw_range=[-1.5, 1.5],
        head_roll_range=[-0.5, 0.5],
        head_range_factor=1.0,
    )


def standing_default_config() -> ConfigDict:
    """``default_config()`` of the reference Standing task (open_duck_mini_v2/standing.py:44-102)."""
    return _cd(
        ctrl_dt=0.02,
        sim_dt=0.002,
        episode_length=1000,
        action_repeat=1,
        action_scale=0.25,
        dof_vel_scale=0.05,
        history_len=0,
        soft_joint_pos_limit_factor=0.95,
        noise_config=_cd(
            level=1.0,
            action_min_delay=0,
            action_max_delay=3,
            imu_min_delay=0,
            imu_max_delay=3,
            scales=_cd(hip_pos=0.03, knee_pos=0.05, ankle_pos=0.08, joint_vel=2.5, gravity=0.1, linvel=0.1, gyro=0.05,
                       accelerometer=0.005),
        ),
        reward_config=_cd(
            scales=_cd(orientation=-0.5, torques=-1.0e-3, action_rate=-0.375, stand_still=-0.3, alive=20.0, head_pos=-2.0),
            tracking_sigma=0.01,
        ),
        push_config=_cd(enable=True, interval_range=[5.0, 10.0], magnitude_range=[0.1, 1.0]),
        neck_pitch_range=[-0.34, 1.1],
        head_pitch_range=[-0.78, 0.78],
        head_yaw_range=[-2.7, 2.7],
        head_roll_range=[-0.5, 0.5],
        head_range_factor=1.0,
    )


def qpos_noise_scale(config: ConfigDict, nu: int) -> np.ndarray:
    """joystick.py:184-200: indices come from the 10-entry JOINTS_ORDER_NO_HEAD but index a 14-long
    per-actuator array, so the scales land on actuator slots 0..9."""
    out = np.zeros(nu)
    s = config.noise_config.scales
    for idx, j in enumerate(JOINTS_ORDER_NO_HEAD):
        if "_hip" in j:
            out[idx] = s.hip_pos
        if "_knee" in j:
            out[idx] = s.knee_pos
        if "_ankle" in j:
            out[idx] = s.ankle_pos
    return out


TASK_TERMS = {"tracking_lin_vel", "tracking_ang_vel", "torques", "action_rate", "stand_still", "alive", "imitation", "head_pos"}


def _fill_reward_library(lib, model: CompiledModel, rc: ConfigDict, standing: bool) -> None:
    """Any other key of ``reward_config.scales`` selects a term of the reward library (include/oduck.h OduckRewardLibrary;
    common/rewards.py:37-224), the way the reference selects terms by listing them in ``scales``.  Parameters come from
    ``reward_config`` with the defaults of the functions' signatures / of this robot: ``base_height_target`` (keyframe
    height), ``max_foot_height`` (0.03 m), ``air_time_threshold_min/max`` (0.1 / 0.5 s, rewards.py:216-217),
    ``soft_joint_pos_limit_factor`` (0.95 of the actuator ranges about their centres), ``pose_weights`` (ones),
    ``base_y_swing_freq`` / ``base_y_swing_amplitude`` (one sway per 0.54 s gait period, 0.05 m/s),
    ``hip_joints`` (hip yaw + roll) and ``knee_joints`` (actuator names)."""
    for k, v in rc.scales.items():
        if k in TASK_TERMS or (k == "orientation" and standing):
            continue
        if k not in capi.LIB_TERMS:
            raise ValueError(f"reward_config.scales: unknown term {k!r} (task terms {sorted(TASK_TERMS)}, library terms {capi.LIB_TERMS})")
        lib.scale[capi.LIB_TERMS.index(k)] = float(v)
    nu = model.nu
    lib.base_height_target = float(rc.get("base_height_target", model.key_qpos[2]))
    lib.max_foot_height = float(rc.get("max_foot_height", 0.03))
    lib.air_time_threshold_min = float(rc.get("air_time_threshold_min", 0.1))
    lib.air_time_threshold_max = float(rc.get("air_time_threshold_max", 0.5))
    # gait-clocked terms (include/oduck.h): lateral sway once per gait period unless told otherwise (0.54 s: poly_reference_motion.py:54-60)
    lib.base_y_swing_freq = float(rc.get("base_y_swing_freq", 1.0 / 0.54))
    lib.base_y_swing_amplitude = float(rc.get("base_y_swing_amplitude", 0.05))
    factor = float(rc.get("soft_joint_pos_limit_factor", 0.95))
    lo, hi = np.asarray(model.act_ctrlrange[:nu, 0], np.float64), np.asarray(model.act_ctrlrange[:nu, 1], np.float64)   # inheritrange: ctrlrange = joint range
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * factor
    w = np.asarray(rc.get("pose_weights", np.ones(nu)), np.float64)
    if w.shape != (nu,):
        raise ValueError(f"reward_config.pose_weights must have {nu} entries")
    for i in range(nu):
        lib.soft_lowers[i], lib.soft_uppers[i], lib.pose_weights[i] = float(mid[i] - half[i]), float(mid[i] + half[i]), float(w[i])
    names = list(model.actuator_names)
    hips = [names.index(j) for j in rc.get("hip_joints", [n for n in names if n.endswith(("hip_yaw", "hip_roll"))])]
    knees = [names.index(j) for j in rc.get("knee_joints", [n for n in names if n.endswith("knee")])]
    if len(hips) > 4 or len(knees) > 4:
        raise ValueError("at most 4 hip and 4 knee joints")
    lib.n_hip, lib.n_knee = len(hips), len(knees)
    for i, j in enumerate(hips):
        lib.hip_indices[i] = j
    for i, j in enumerate(knees):
        lib.knee_indices[i] = j


def build_env_config(model: CompiledModel, config: ConfigDict, poly: Optional[PolyTable], auto_reset: bool = True,
                     use_imitation_reward: bool = USE_IMITATION_REWARD,
                     use_motor_speed_limits: bool = USE_MOTOR_SPEED_LIMITS, task: int = capi.TASK_JOYSTICK):
    """Pack ``config`` into the C struct.  Returns (struct, keepalive) -- keepalive owns the coefficient array.
    ``task``: capi.TASK_JOYSTICK (joystick.py) or capi.TASK_STANDING (standing.py: no imitation reward, no motor speed limits)."""
    c = capi.OduckEnvConfig()
    c.task = int(task)
    standing = task == capi.TASK_STANDING
    if standing:
        use_imitation_reward, use_motor_speed_limits = False, False
    n_sub = int(round(config.ctrl_dt / config.sim_dt))
    c.n_substeps = n_sub
    c.episode_length = int(config.episode_length)
    c.use_imitation_reward = int(use_imitation_reward)
    c.use_motor_speed_limits = int(use_motor_speed_limits)
    c.push_enable = int(bool(config.push_config.enable))
    nc = config.noise_config
    c.action_min_delay, c.action_max_delay = int(nc.action_min_delay), int(nc.action_max_delay)
    c.imu_min_delay, c.imu_max_delay = int(nc.imu_min_delay), int(nc.imu_max_delay)
    c.auto_reset = int(auto_reset)
    c.ctrl_dt = float(config.ctrl_dt)
    c.action_scale = float(config.action_scale)
    c.dof_vel_scale = float(config.dof_vel_scale)
    c.max_motor_velocity = float(config.get("max_motor_velocity", 0.0))
    c.reset_base_qvel_noise = 0.5 if standing else 0.05          # standing.py:247 / joystick.py:253
    c.noise_level = float(nc.level)
    c.noise_gyro, c.noise_accelerometer = float(nc.scales.gyro), float(nc.scales.accelerometer)
    c.noise_gravity, c.noise_joint_vel = float(nc.scales.gravity), float(nc.scales.joint_vel)
    qn = qpos_noise_scale(config, model.nu)
    for i in range(model.nu):
        c.qpos_noise_scale[i] = qn[i]
    rs = config.reward_config.scales
    g = lambda k: float(rs.get(k, 0.0))
    c.scale_tracking_lin_vel, c.scale_tracking_ang_vel = g("tracking_lin_vel"), g("tracking_ang_vel")
    c.scale_torques, c.scale_action_rate = g("torques"), g("action_rate")
    c.scale_stand_still, c.scale_alive, c.scale_imitation = g("stand_still"), g("alive"), g("imitation")
    c.scale_orientation, c.scale_head_pos = (g("orientation") if standing else 0.0), g("head_pos")
    c.tracking_sigma = float(config.reward_config.tracking_sigma)
    _fill_reward_library(c.lib, model, config.reward_config, standing)
    for i in range(2):
        c.push_interval_range[i] = float(config.push_config.interval_range[i])
        c.push_magnitude_range[i] = float(config.push_config.magnitude_range[i])
    f = float(config.head_range_factor)
    zero = [0.0, 0.0]
    ranges = [config.get("lin_vel_x", zero), config.get("lin_vel_y", zero), config.get("ang_vel_yaw", zero),
              [config.neck_pitch_range[0] * f, config.neck_pitch_range[1] * f],
              [config.head_pitch_range[0] * f, config.head_pitch_range[1] * f],
              [config.head_yaw_range[0] * f, config.head_yaw_range[1] * f],
              [config.head_roll_range[0] * f, config.head_roll_range[1] * f]]
    for i, r in enumerate(ranges):
        c.cmd_range[i][0], c.cmd_range[i][1] = float(r[0]), float(r[1])
    keep = None
    if use_imitation_reward:
        if poly is None:
            raise ValueError("USE_IMITATION_REWARD needs the polynomial reference-motion table")
        if len(poly.dxs) > 8 or len(poly.dys) > 8 or len(poly.dthetas) > 16:
            raise ValueError("reference-motion grid exceeds the C struct limits")
        c.ndx, c.ndy, c.ndth = len(poly.dxs), len(poly.dys), len(poly.dthetas)
        c.nb_steps_in_period = int(poly.nb_steps_in_period)
        for i, v in enumerate(poly.dxs):
            c.dxs[i] = v
        for i, v in enumerate(poly.dys):
            c.dys[i] = v
        for i, v in enumerate(poly.dthetas):
            c.dthetas[i] = v
        for i in range(2):
            c.dx_range[i], c.dy_range[i], c.dtheta_range[i] = poly.dx_range[i], poly.dy_range[i], poly.dtheta_range[i]
        keep = np.ascontiguousarray(poly.coef, dtype=np.float64)
        c.poly_coef = keep.ctypes.data_as(C.POINTER(C.c_double))
    else:
        c.nb_steps_in_period = 1
    return c, keep
