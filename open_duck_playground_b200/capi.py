"""ctypes mirror of include/oduck.h.

One :class:`Library` wraps one shared object exporting the ``oduck_*`` C-ABI.  The product loads
``csrc/liboduck_cuda.so`` through :func:`load_cuda_library`; the CPU oracle (``oracle/``) exports the same
symbols and is loaded through this same class by the tests only -- nothing in this package references it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from .mjcf import CompiledModel

ABI_VERSION = 9
MAX_BODY, MAX_JNT, MAX_NQ, MAX_NV, MAX_NU, MAX_SITE, MAX_VERT, MAX_FACE, NFEET, NCMD = 20, 28, 36, 32, 16, 8, 32, 64, 2, 7
OBS_STATE, OBS_PRIV, NMETRIC, REF_DIM, POLY_DEG, MAX_CON = 101, 212, 8, 40, 16, 12

d, i32 = C.c_double, C.c_int32


class OduckModel(C.Structure):
    _fields_ = [
        ("abi_version", i32), ("nbody", i32), ("njnt", i32), ("nq", i32), ("nv", i32), ("nu", i32), ("nsite", i32),
        ("body_parentid", i32 * MAX_BODY), ("body_jntadr", i32 * MAX_BODY), ("body_jntnum", i32 * MAX_BODY),
        ("body_dofadr", i32 * MAX_BODY), ("body_dofnum", i32 * MAX_BODY),
        ("body_pos", d * 3 * MAX_BODY), ("body_quat", d * 4 * MAX_BODY), ("body_ipos", d * 3 * MAX_BODY),
        ("body_iquat", d * 4 * MAX_BODY), ("body_mass", d * MAX_BODY), ("body_inertia", d * 3 * MAX_BODY),
        ("body_invweight0", d * 2 * MAX_BODY),
        ("jnt_type", i32 * MAX_JNT), ("jnt_qposadr", i32 * MAX_JNT), ("jnt_dofadr", i32 * MAX_JNT),
        ("jnt_bodyid", i32 * MAX_JNT), ("jnt_limited", i32 * MAX_JNT),
        ("jnt_pos", d * 3 * MAX_JNT), ("jnt_axis", d * 3 * MAX_JNT), ("jnt_range", d * 2 * MAX_JNT),
        ("qpos0", d * MAX_NQ),
        ("dof_bodyid", i32 * MAX_NV), ("dof_jntid", i32 * MAX_NV), ("dof_parentid", i32 * MAX_NV),
        ("dof_armature", d * MAX_NV), ("dof_damping", d * MAX_NV), ("dof_frictionloss", d * MAX_NV),
        ("dof_invweight0", d * MAX_NV),
        ("act_jntid", i32 * MAX_NU), ("act_kp", d * MAX_NU), ("act_kv", d * MAX_NU),
        ("act_ctrlrange", d * 2 * MAX_NU), ("act_forcerange", d * 2 * MAX_NU),
        ("site_bodyid", i32 * MAX_SITE), ("site_pos", d * 3 * MAX_SITE), ("site_quat", d * 4 * MAX_SITE),
        ("imu_site", i32), ("foot_site", i32 * NFEET),
        ("floor_is_hfield", i32), ("floor_friction", d), ("foot_body", i32 * NFEET), ("foot_nvert", i32),
        ("foot_vert", d * 3 * MAX_VERT * NFEET), ("foot_nface", i32), ("foot_face", i32 * 3 * MAX_FACE),
        ("foot_nplane", i32), ("foot_plane_nvert", i32 * 32), ("foot_plane_vert", i32 * 8 * 32), ("foot_plane_normal", d * 3 * 32 * NFEET),
        ("foot_nedge", i32), ("foot_edge_vert", i32 * 2 * 48), ("foot_edge_plane", i32 * 2 * 48), ("foot_center", d * 3 * NFEET), ("foot_radius", d),
        ("foot_friction", d), ("enable_foot_foot", i32),
        ("timestep", d), ("gravity", d * 3), ("tolerance", d), ("ls_tolerance", d), ("impratio", d), ("meaninertia", d),
        ("iterations", i32), ("ls_iterations", i32), ("solref", d * 2), ("solimp", d * 5),
        ("key_qpos", d * MAX_NQ), ("key_ctrl", d * MAX_NU),
        ("hfield_nrow", i32), ("hfield_ncol", i32), ("hfield_size", d * 4), ("hfield_data", C.POINTER(C.c_float)),
    ]


LIB_TERMS = ["orientation", "lin_vel_z", "ang_vel_xy", "base_height", "energy", "joint_pos_limits", "termination", "joint_deviation_hip",
             "joint_deviation_knee", "pose", "feet_slip", "feet_clearance", "feet_height", "feet_air_time", "base_y_swing", "feet_phase"]   # enum OduckLibTerm


class OduckRewardLibrary(C.Structure):
    _fields_ = [
        ("scale", d * len(LIB_TERMS)),
        ("base_height_target", d), ("max_foot_height", d), ("air_time_threshold_min", d), ("air_time_threshold_max", d),
        ("base_y_swing_freq", d), ("base_y_swing_amplitude", d),
        ("soft_lowers", d * MAX_NU), ("soft_uppers", d * MAX_NU), ("pose_weights", d * MAX_NU),
        ("n_hip", i32), ("hip_indices", i32 * 4), ("n_knee", i32), ("knee_indices", i32 * 4),
    ]


class OduckEnvConfig(C.Structure):
    _fields_ = [
        ("task", i32), ("n_substeps", i32), ("episode_length", i32), ("use_imitation_reward", i32), ("use_motor_speed_limits", i32),
        ("push_enable", i32), ("action_min_delay", i32), ("action_max_delay", i32), ("imu_min_delay", i32),
        ("imu_max_delay", i32), ("auto_reset", i32),
        ("ctrl_dt", d), ("action_scale", d), ("dof_vel_scale", d), ("max_motor_velocity", d), ("noise_level", d),
        ("noise_gyro", d), ("noise_accelerometer", d), ("noise_gravity", d), ("noise_joint_vel", d),
        ("qpos_noise_scale", d * MAX_NU),
        ("scale_tracking_lin_vel", d), ("scale_tracking_ang_vel", d), ("scale_torques", d), ("scale_action_rate", d),
        ("scale_stand_still", d), ("scale_alive", d), ("scale_imitation", d),
        ("scale_orientation", d), ("scale_head_pos", d), ("reset_base_qvel_noise", d), ("tracking_sigma", d),
        ("push_interval_range", d * 2), ("push_magnitude_range", d * 2), ("cmd_range", d * 2 * NCMD),
        ("ndx", i32), ("ndy", i32), ("ndth", i32), ("nb_steps_in_period", i32),
        ("dxs", d * 8), ("dys", d * 8), ("dthetas", d * 16),
        ("dx_range", d * 2), ("dy_range", d * 2), ("dtheta_range", d * 2),
        ("poly_coef", C.POINTER(d)),
        ("lib", OduckRewardLibrary),
    ]


class OduckPolicyWeights(C.Structure):
    _fields_ = [
        ("obs_dim", i32), ("hidden", i32 * 3), ("out_dim", i32),
        ("obs_mean", C.c_void_p), ("obs_std", C.c_void_p), ("w", C.c_void_p * 4), ("b", C.c_void_p * 4), ("packed", C.c_void_p * 4),
    ]


class OduckRolloutSink(C.Structure):
    """include/oduck.h: caller-owned rollout buffers the kernels of oduck_rollout_step write into."""
    _fields_ = [
        ("unroll", i32), ("num_envs", i32), ("env_offset", i32), ("policy_dim", i32), ("value_dim", i32),
        ("obs_policy", C.c_void_p), ("obs_value", C.c_void_p), ("raw_action", C.c_void_p), ("log_prob", C.c_void_p),
        ("reward", C.c_void_p), ("done", C.c_void_p), ("truncation", C.c_void_p),
    ]


class OduckPpoConfig(C.Structure):
    """include/oduck_ppo.h"""
    _fields_ = [
        ("batch_envs", i32), ("unroll", i32), ("num_actions", i32), ("policy_dims", i32 * 5), ("value_dims", i32 * 5),
        ("normalize_advantage", i32),
        ("discounting", C.c_float), ("gae_lambda", C.c_float), ("clipping_epsilon", C.c_float), ("entropy_cost", C.c_float),
        ("reward_scaling", C.c_float), ("learning_rate", C.c_float), ("max_grad_norm", C.c_float),
        ("adam_b1", C.c_float), ("adam_b2", C.c_float), ("adam_eps", C.c_float), ("matmul_tf32", i32),
    ]


class OduckRollout(C.Structure):
    _fields_ = [
        ("num_envs", i32), ("unroll", i32),
        ("obs_policy", C.c_void_p), ("obs_value", C.c_void_p), ("raw_action", C.c_void_p), ("log_prob", C.c_void_p),
        ("reward", C.c_void_p), ("done", C.c_void_p), ("truncation", C.c_void_p),
        ("block_envs", i32), ("block_stride", C.c_int64), ("obs_policy_ld", i32),
    ]


class OduckNormalizer(C.Structure):
    _fields_ = [("policy_mean", C.c_void_p), ("policy_std", C.c_void_p), ("value_mean", C.c_void_p), ("value_std", C.c_void_p)]


PPO_STAGE_FORWARD, PPO_STAGE_LOSS, PPO_STAGE_BACKWARD, PPO_STAGE_ADAM, PPO_ALL, PPO_DEBUG_SIMT, PPO_NO_COOP = 1, 2, 4, 8, 15, 256, 512
PPO_BUF = {name: k for k, name in enumerate(["PARAMS", "GRADS", "ADAM_M", "ADAM_V", "LOGITS", "VALUES", "LOSSES", "ADV", "VS", "STEP"])}

BUF = {name: k for k, name in enumerate([
    "QPOS", "QVEL", "QACC_WARM", "QACC", "CTRL", "OBS_STATE", "OBS_PRIV", "REWARD", "DONE", "TRUNCATION", "METRICS",
    "EFC_FORCE", "CONTACT_DIST", "SENSORDATA", "ACTUATOR_FORCE", "SITE_XPOS_FEET", "INFO_RNG", "INFO_COMMAND",
    "INFO_STEP", "INFO_STEPS", "INFO_LAST_ACT", "INFO_MOTOR_TARGETS", "INFO_FEET_AIR_TIME", "INFO_LAST_CONTACT",
    "INFO_SWING_PEAK", "INFO_PUSH", "INFO_PUSH_STEP", "INFO_PUSH_INTERVAL", "INFO_ACTION_HISTORY", "INFO_IMU_HISTORY",
    "INFO_IMITATION_I", "INFO_REF_MOTION", "INFO_IMITATION_PHASE", "DR_PARAMS", "FIRST_QPOS", "FIRST_QVEL",
    "FIRST_OBS_STATE", "FIRST_OBS_PRIV"])}
DTYPE_NP = {0: np.float32, 1: np.int32, 2: np.uint32, 3: np.float64}
METRIC_NAMES = ["reward/tracking_lin_vel", "reward/tracking_ang_vel", "cost/torques", "cost/action_rate",
                "cost/stand_still", "reward/alive", "reward/imitation", "swing_peak"]
METRIC_NAMES_STANDING = ["cost/orientation", "cost/torques", "cost/action_rate", "cost/stand_still", "reward/alive", "cost/head_pos", "swing_peak"]
TASK_JOYSTICK, TASK_STANDING = 0, 1
OBS_DIMS = {TASK_JOYSTICK: (OBS_STATE, OBS_PRIV), TASK_STANDING: (85, 153)}


def model_to_struct(m: CompiledModel) -> OduckModel:
    s = OduckModel()
    s.abi_version = ABI_VERSION
    for k in ("nbody", "njnt", "nq", "nv", "nu", "nsite"):
        setattr(s, k, int(getattr(m, k)))
    for name, ctype in OduckModel._fields_:
        if name in ("abi_version", "nbody", "njnt", "nq", "nv", "nu", "nsite"):
            continue
        if name.startswith("hfield_"):                       # optional height-field asset (rough terrain scenes)
            if "hfield_data" in m.arrays and name == "hfield_data":
                data = np.ascontiguousarray(m.arrays["hfield_data"], dtype=np.float32)
                s._hfield_keep = data                        # the struct holds a pointer: keep the array alive with it
                s.hfield_data = data.ctypes.data_as(C.POINTER(C.c_float))
                s.hfield_nrow, s.hfield_ncol = int(data.shape[0]), int(data.shape[1])
            elif name == "hfield_size" and "hfield_size" in m.arrays:
                for k in range(4):
                    s.hfield_size[k] = float(m.arrays["hfield_size"][k])
            continue
        a = m.arrays[name]
        if isinstance(getattr(s, name), (int, float)):
            setattr(s, name, a.item())
        else:
            dst = np.ctypeslib.as_array(getattr(s, name))
            dst[...] = np.asarray(a).reshape(dst.shape)
    return s


class OduckError(RuntimeError):
    pass


class Library:
    """A loaded ``liboduck_*.so``.  ``is_device`` says whether pointers are CUDA device pointers."""

    def __init__(self, path: str, is_device: bool):
        if not os.path.exists(path):
            raise OduckError(f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.is_device = is_device
        self.lib = C.CDLL(path)
        L = self.lib
        vp, p = C.c_void_p, C.POINTER
        L.oduck_last_error.restype = C.c_char_p
        L.oduck_create.argtypes = [p(OduckModel), p(OduckEnvConfig), C.c_int, C.c_int, p(vp)]
        L.oduck_destroy.argtypes = [vp]
        L.oduck_num_envs.argtypes = [vp]
        L.oduck_randomize.argtypes = [vp, vp, vp]
        L.oduck_reset.argtypes = [vp, vp, vp, vp]
        L.oduck_step.argtypes = [vp, vp, vp]
        L.oduck_physics_substeps.argtypes = [vp, vp, C.c_int, vp]
        L.oduck_forward.argtypes = [vp, vp]
        L.oduck_set_state.argtypes = [vp, vp, vp, vp, vp]
        L.oduck_policy_forward.argtypes = [vp, p(OduckPolicyWeights), vp, vp, C.c_int, vp, vp, vp, vp]
        L.oduck_get_buffer.argtypes = [vp, C.c_int, p(vp), p(C.c_int64), p(C.c_int64), p(C.c_int)]
        L.oduck_launch_count.argtypes = [vp]
        L.oduck_launch_count.restype = C.c_int64
        L.oduck_policy_invalidate.argtypes = [vp]
        L.oduck_set_rollout_sink.argtypes = [vp, p(OduckRolloutSink)]
        L.oduck_rollout_step.argtypes = [vp, p(OduckPolicyWeights), vp, C.c_int, vp]
        self.has_ppo = hasattr(L, "oduck_ppo_create")            # the device learner (include/oduck_ppo.h) exists in the CUDA library only
        if self.has_ppo:
            L.oduck_ppo_create.argtypes = [p(OduckPpoConfig), C.c_int, p(vp)]
            L.oduck_ppo_destroy.argtypes = [vp]
            L.oduck_ppo_num_params.argtypes = [vp]
            L.oduck_ppo_num_params.restype = C.c_int64
            L.oduck_ppo_launch_count.argtypes = [vp]
            L.oduck_ppo_launch_count.restype = C.c_int64
            L.oduck_ppo_param_info.argtypes = [vp, C.c_int, C.c_int, C.c_int, p(C.c_int64), p(C.c_int64), p(C.c_int64)]
            L.oduck_ppo_set_params.argtypes = [vp, vp, C.c_int, vp]
            L.oduck_ppo_get_buffer.argtypes = [vp, C.c_int, p(vp), p(C.c_int64), p(C.c_int)]
            L.oduck_ppo_minibatch.argtypes = [vp, p(OduckRollout), p(OduckNormalizer), vp, vp, vp, C.c_int, vp]
            L.oduck_ppo_prefetch.argtypes = [vp, p(OduckRollout), p(OduckNormalizer), vp, vp]
            L.oduck_ppo_packed_weights.argtypes = [vp, C.c_int, C.c_int, p(vp)]
        if L.oduck_abi_version() != ABI_VERSION:
            raise OduckError(f"{path}: ABI version {L.oduck_abi_version()} != {ABI_VERSION}")
        if L.oduck_sizeof_model() != C.sizeof(OduckModel) or L.oduck_sizeof_env_config() != C.sizeof(OduckEnvConfig):
            raise OduckError(f"{path}: struct layout mismatch with capi.py")

    def check(self, rc: int) -> None:
        if rc != 0:
            raise OduckError(f"oduck error {rc}: {self.lib.oduck_last_error().decode()}")

    def create(self, model: OduckModel, cfg: OduckEnvConfig, num_envs: int, device: int = 0) -> "Handle":
        h = C.c_void_p()
        self.check(self.lib.oduck_create(C.byref(model), C.byref(cfg), num_envs, device, C.byref(h)))
        return Handle(self, h, num_envs)


class Handle:
    """Owns one ``OduckHandle*``.  All array arguments are raw addresses (``int``) in the library's memory space."""

    def __init__(self, lib: Library, h: C.c_void_p, n: int):
        self.L, self.h, self.n = lib, h, n

    def close(self):
        if self.h:
            self.L.lib.oduck_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def randomize(self, keys: int, stream: int = 0):
        self.L.check(self.L.lib.oduck_randomize(self.h, keys, stream))

    def reset(self, keys: int, mask: int = 0, stream: int = 0):
        self.L.check(self.L.lib.oduck_reset(self.h, keys, mask or None, stream))

    def step(self, action: int, stream: int = 0):
        self.L.check(self.L.lib.oduck_step(self.h, action, stream))

    def physics_substeps(self, ctrl: int, n: int, stream: int = 0):
        self.L.check(self.L.lib.oduck_physics_substeps(self.h, ctrl or None, n, stream))

    def forward(self, stream: int = 0):
        self.L.check(self.L.lib.oduck_forward(self.h, stream))

    def set_state(self, qpos: int = 0, qvel: int = 0, qacc_warm: int = 0, stream: int = 0):
        self.L.check(self.L.lib.oduck_set_state(self.h, qpos or None, qvel or None, qacc_warm or None, stream))

    def policy_forward(self, w: OduckPolicyWeights, obs: int, keys: int, deterministic: bool, action: int,
                       raw_action: int, log_prob: int, stream: int = 0):
        self.L.check(self.L.lib.oduck_policy_forward(self.h, C.byref(w), obs or None, keys or None, int(deterministic),
                                                     action or None, raw_action or None, log_prob or None, stream))

    def set_rollout_sink(self, sink: Optional[OduckRolloutSink]) -> None:
        self.L.check(self.L.lib.oduck_set_rollout_sink(self.h, C.byref(sink) if sink is not None else None))

    def rollout_step(self, w: OduckPolicyWeights, keys: int, t: int, stream: int = 0) -> None:
        self.L.check(self.L.lib.oduck_rollout_step(self.h, C.byref(w), keys, int(t), stream))

    def launch_count(self) -> int:
        return int(self.L.lib.oduck_launch_count(self.h))

    def policy_invalidate(self) -> None:
        self.L.check(self.L.lib.oduck_policy_invalidate(self.h))

    def buffer_info(self, name: str) -> Tuple[int, Tuple[int, ...], Tuple[int, ...], type]:
        ptr, shape, strides, dt = C.c_void_p(), (C.c_int64 * 4)(), (C.c_int64 * 4)(), C.c_int()
        self.L.check(self.L.lib.oduck_get_buffer(self.h, BUF[name], C.byref(ptr), shape, strides, C.byref(dt)))
        nd = 1 + sum(1 for k in (1, 2, 3) if shape[k] > 0)
        return ptr.value, tuple(shape[:nd]), tuple(strides[:nd]), DTYPE_NP[dt.value]

    def buffer_numpy(self, name: str) -> np.ndarray:
        """Zero-copy numpy view (host libraries only)."""
        if self.L.is_device:
            raise OduckError("buffer_numpy on a device library; use Joystick.buffer()")
        ptr, shape, strides, dt = self.buffer_info(name)
        item = np.dtype(dt).itemsize
        n_items = 1 + sum((s - 1) * st for s, st in zip(shape, strides))
        raw = (C.c_char * (n_items * item)).from_address(ptr)
        base = np.frombuffer(raw, dtype=dt)
        return np.lib.stride_tricks.as_strided(base, shape=shape, strides=tuple(st * item for st in strides))


class PpoHandle:
    """Owns one ``OduckPpo*`` (include/oduck_ppo.h): the on-device PPO learner step."""

    def __init__(self, lib: Library, cfg: OduckPpoConfig, device: int = 0):
        if not lib.has_ppo:
            raise OduckError(f"{lib.path} does not export the oduck_ppo_* learner (CUDA library only)")
        self.L, self.cfg = lib, cfg
        h = C.c_void_p()
        lib.check(lib.lib.oduck_ppo_create(C.byref(cfg), device, C.byref(h)))
        self.h = h
        self.num_params = int(lib.lib.oduck_ppo_num_params(h))

    def close(self):
        if self.h:
            self.L.lib.oduck_ppo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def param_info(self, net: int, layer: int, which: int) -> Tuple[int, int, int]:
        off, r, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.L.check(self.L.lib.oduck_ppo_param_info(self.h, net, layer, which, C.byref(off), C.byref(r), C.byref(c)))
        return off.value, r.value, c.value

    def packed_weights(self, net: int, layer: int) -> int:
        ptr = C.c_void_p()
        self.L.check(self.L.lib.oduck_ppo_packed_weights(self.h, net, layer, C.byref(ptr)))
        return ptr.value

    def set_params(self, flat: int, reset_opt: bool, stream: int = 0):
        self.L.check(self.L.lib.oduck_ppo_set_params(self.h, flat, int(reset_opt), stream))

    def buffer_info(self, name: str) -> Tuple[int, int, type]:
        ptr, cnt, dt = C.c_void_p(), C.c_int64(), C.c_int()
        self.L.check(self.L.lib.oduck_ppo_get_buffer(self.h, PPO_BUF[name], C.byref(ptr), C.byref(cnt), C.byref(dt)))
        return ptr.value, cnt.value, DTYPE_NP[dt.value]

    def minibatch(self, rollout: OduckRollout, norm: OduckNormalizer, env_idx: int, noise: int, key: int, stages: int, stream: int = 0):
        self.L.check(self.L.lib.oduck_ppo_minibatch(self.h, C.byref(rollout), C.byref(norm), env_idx, noise or None, key or None, stages, stream))

    def prefetch(self, rollout: OduckRollout, norm: OduckNormalizer, next_env_idx: int, stream: int = 0):
        self.L.check(self.L.lib.oduck_ppo_prefetch(self.h, C.byref(rollout), C.byref(norm), next_env_idx, stream))

    def launch_count(self) -> int:
        return int(self.L.lib.oduck_ppo_launch_count(self.h))


_cuda_lib: Optional[Library] = None


def cuda_library_path() -> str:
    """In-tree product library; ODUCK_CUDA_LIB points at another build of the same sources (kernel tuning, tools/variants.py)."""
    override = os.environ.get("ODUCK_CUDA_LIB")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "liboduck_cuda.so")


def load_cuda_library() -> Library:
    """Load the product library.  Fails loudly when it has not been built: there is no CPU fallback."""
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = Library(cuda_library_path(), is_device=True)
    return _cuda_lib
