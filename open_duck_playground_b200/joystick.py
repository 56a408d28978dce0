"""Joystick task for Open Duck Mini V2 -- drop-in surface over the B200 step library.

Mirrors the reference env API (open_duck_mini_v2/joystick.py:105-725, base.py:41-291):

    env = Joystick(task="flat_terrain_backlash")
    state = env.reset(rng)            # rng: uint32 [N, 2] jax.random key data (or one key -> N = 1)
    state = env.step(state, action)   # action: float32 [N, 14] on the env's device

``State`` carries ``data / obs / reward / done / metrics / info`` like ``mjx_env.State``; ``obs`` is the dict
``{"state": [N,101], "privileged_state": [N,212]}`` (joystick.py:617-620).  Differences forced by the design:
the env is batched natively (the reference is vmapped by Brax's wrapper), and ``step`` already includes the
Episode/AutoReset wrapper semantics of ``wrapper.wrap_for_brax_training`` (common/runner.py:117), fused in the kernel.
All tensors in a ``State`` are zero-copy views of library-owned device buffers: they are overwritten by the next
``reset``/``step`` (call ``state.clone()`` to keep a snapshot).  There is no CPU fallback: without
``csrc/liboduck_cuda.so`` construction fails.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Union

import numpy as np
import torch

from . import capi, config as config_mod, constants
from .config import ConfigDict, default_config  # noqa: F401  (re-export, reference: joystick.default_config)
from .mjcf import CompiledModel, compile_mjcf
from .poly_reference_motion import PolyTable

_TORCH_DTYPE = {np.float32: torch.float32, np.int32: torch.int32, np.uint32: torch.int32, np.float64: torch.float64}
_TYPESTR = {np.float32: "<f4", np.int32: "<i4", np.uint32: "<i4", np.float64: "<f8"}


class _CudaView:
    def __init__(self, ptr, shape, strides, np_dtype):
        item = np.dtype(np_dtype).itemsize
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": _TYPESTR[np_dtype], "data": (int(ptr), False), "version": 2,
            "strides": tuple(int(s) * item for s in strides),
        }


@dataclass
class Data:
    """The slice of ``mjx.Data`` the task reads (views)."""
    qpos: torch.Tensor
    qvel: torch.Tensor
    qacc_warmstart: torch.Tensor
    qacc: torch.Tensor
    ctrl: torch.Tensor
    sensordata: torch.Tensor
    actuator_force: torch.Tensor
    efc_force: torch.Tensor
    contact_dist: torch.Tensor
    site_xpos_feet: torch.Tensor


@dataclass
class State:
    data: Data
    obs: Dict[str, torch.Tensor]
    reward: torch.Tensor
    done: torch.Tensor
    metrics: Dict[str, torch.Tensor]
    info: Dict[str, Any] = field(default_factory=dict)

    def replace(self, **kw) -> "State":
        return State(**{**self.__dict__, **kw})

    def clone(self) -> "State":
        c = lambda t: t.clone() if isinstance(t, torch.Tensor) else copy.copy(t)
        return State(Data(**{k: c(v) for k, v in self.data.__dict__.items()}), {k: c(v) for k, v in self.obs.items()},
                     c(self.reward), c(self.done), {k: c(v) for k, v in self.metrics.items()}, {k: c(v) for k, v in self.info.items()})


_INFO_BUFFERS = {
    "rng": "INFO_RNG", "step": "INFO_STEP", "steps": "INFO_STEPS", "command": "INFO_COMMAND", "motor_targets": "INFO_MOTOR_TARGETS",
    "feet_air_time": "INFO_FEET_AIR_TIME", "last_contact": "INFO_LAST_CONTACT", "swing_peak": "INFO_SWING_PEAK", "push": "INFO_PUSH",
    "push_step": "INFO_PUSH_STEP", "push_interval_steps": "INFO_PUSH_INTERVAL", "action_history": "INFO_ACTION_HISTORY",
    "imu_history": "INFO_IMU_HISTORY", "imitation_i": "INFO_IMITATION_I", "current_reference_motion": "INFO_REF_MOTION",
    "imitation_phase": "INFO_IMITATION_PHASE", "truncation": "TRUNCATION",
}


class Joystick:
    """Track a joystick command (batched; reference class: joystick.py:105)."""

    TASK = capi.TASK_JOYSTICK                      # which env class of the reference the library steps (ODUCK_TASK_*)
    METRICS = capi.METRIC_NAMES
    DEFAULT_CONFIG = staticmethod(default_config)

    def __init__(self, task: str = "flat_terrain", config: Optional[ConfigDict] = None,
                 config_overrides: Optional[Dict[str, Union[str, int, list]]] = None, *, num_envs: Optional[int] = None,
                 device: Union[str, int, torch.device] = "cuda:0", xml_path: Optional[str] = None,
                 library: Optional[capi.Library] = None, auto_reset: bool = True):
        self._config = copy.deepcopy(config) if config is not None else self.DEFAULT_CONFIG()
        if config_overrides:
            self._config.update_from_flattened_dict(config_overrides)
        self._task = task
        if xml_path is not None:                                  # user-provided MJCF (reference checkout)
            self._mj_model = compile_mjcf(xml_path, timestep=self._config.sim_dt)
            self._xml_path = xml_path
        else:
            self._xml_path = constants.task_to_xml(task)          # KeyError on unknown task, like the reference
            self._mj_model = CompiledModel.load(constants.task_to_blob(task))
            self._mj_model.arrays["timestep"] = np.array(float(self._config.sim_dt))   # base.py:56
        self._use_imitation = config_mod.USE_IMITATION_REWARD and self.TASK == capi.TASK_JOYSTICK       # standing.py:42: False
        self.PRM = PolyTable.load(constants.POLY_BLOB) if self._use_imitation else None
        self._lib = library if library is not None else capi.load_cuda_library()
        self._device = torch.device(device) if self._lib.is_device else torch.device("cpu")
        self._auto_reset = auto_reset
        self._handle: Optional[capi.Handle] = None
        self._views: Dict[str, torch.Tensor] = {}
        self._init_q = torch.tensor(self._mj_model.key_qpos[: self._mj_model.nq], dtype=torch.float32)
        self._default_actuator = torch.tensor(self._mj_model.key_ctrl[: self._mj_model.nu], dtype=torch.float32)
        if num_envs is not None:
            self._create(int(num_envs))

    def spawn(self, num_envs: Optional[int] = None) -> "Joystick":
        """A second env of the same class over the SAME compiled model, config, scene path, library, device and auto-reset
        setting (its own library handle and state).  The evaluator's envs and the rollout's sub-batch envs are made this way, so
        an env built from a user MJCF (``xml_path=``) is evaluated and trained on that model, not on the shipped blob."""
        e = type(self).__new__(type(self))
        e.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_handle", "_views", "_keep", "_live")})
        e._config = copy.deepcopy(self._config)
        e._handle, e._views = None, {}
        if num_envs is not None:
            e._create(int(num_envs))
        return e

    # ------------------------------------------------------------------ reference accessors (base.py:277-291 + MjxEnv)
    @property
    def xml_path(self) -> str:
        return str(self._xml_path)

    @property
    def action_size(self) -> int:
        return self._mj_model.nu

    @property
    def mj_model(self) -> CompiledModel:
        return self._mj_model

    @property
    def mjx_model(self) -> CompiledModel:
        return self._mj_model

    @property
    def dt(self) -> float:
        return float(self._config.ctrl_dt)

    @property
    def sim_dt(self) -> float:
        return float(self._config.sim_dt)

    @property
    def n_substeps(self) -> int:
        return int(round(self.dt / self.sim_dt))

    # obs["privileged_state"][:, :101] IS obs["state"] (joystick.py:596-615: privileged_state = hstack([state, ...]); the kernels write
    # the same values into both records, csrc/oduck_env.cuh write_obs): a learner may read its policy rows out of the value rows
    privileged_obs_has_state_prefix = True

    @property
    def observation_size(self) -> Dict[str, tuple]:
        ds, dp = capi.OBS_DIMS[self.TASK]
        return {"state": (ds,), "privileged_state": (dp,)}

    @property
    def unwrapped(self) -> "Joystick":
        return self

    @property
    def num_envs(self) -> Optional[int]:
        return self._handle.n if self._handle else None

    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def handle(self) -> capi.Handle:
        if self._handle is None:
            raise RuntimeError("env has no envs yet: call reset(rng) or pass num_envs")
        return self._handle

    # ------------------------------------------------------------------ plumbing
    def _create(self, n: int) -> None:
        ms = capi.model_to_struct(self._mj_model)
        cs, self._keep = config_mod.build_env_config(self._mj_model, self._config, self.PRM, auto_reset=self._auto_reset,
                                                     use_imitation_reward=self._use_imitation,
                                                     use_motor_speed_limits=config_mod.USE_MOTOR_SPEED_LIMITS, task=self.TASK)
        dev = self._device.index or 0 if self._lib.is_device else 0
        self._handle = self._lib.create(ms, cs, n, dev)
        self._views = {}

    def buffer(self, name: str) -> torch.Tensor:
        """Zero-copy torch view of a library buffer (names: capi.BUF)."""
        if name not in self._views:
            ptr, shape, strides, dt = self.handle.buffer_info(name)
            if self._lib.is_device:
                with torch.cuda.device(self._device):
                    self._views[name] = torch.as_tensor(_CudaView(ptr, shape, strides, dt), device=self._device)
            else:
                arr = self.handle.buffer_numpy(name)
                self._views[name] = torch.from_numpy(arr.view(np.int32) if dt == np.uint32 else arr)
        return self._views[name]

    def _ptr(self, t: Optional[torch.Tensor], dtype, shape=None) -> int:
        if t is None:
            return 0
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        if t.dtype in (torch.uint32, torch.int64, torch.uint64) and dtype == torch.int32:
            t = torch.from_numpy(t.cpu().numpy().astype(np.uint32).view(np.int32)) if t.dtype != torch.uint32 else t.view(torch.int32)
        t = t.to(device=self._device, dtype=dtype).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        self._live = getattr(self, "_live", [])[-7:] + [t]      # keep async arguments alive
        return t.data_ptr()

    def _stream(self) -> int:
        return torch.cuda.current_stream(self._device).cuda_stream if self._lib.is_device else 0

    def _keys(self, rng) -> int:
        rng = np.asarray(rng.cpu() if isinstance(rng, torch.Tensor) else rng)
        if rng.ndim == 1:
            rng = rng[None]
        if self._handle is None:
            self._create(rng.shape[0])
        if rng.shape != (self._handle.n, 2):
            raise ValueError(f"rng must be key data of shape ({self._handle.n}, 2), got {rng.shape}")
        k = torch.from_numpy(np.ascontiguousarray(rng.astype(np.uint32)).view(np.int32))
        return self._ptr(k, torch.int32)

    # ------------------------------------------------------------------ env API
    def randomize(self, rng) -> None:
        """domain_randomize (common/randomize.py:26): per-env friction/frictionloss/armature/COM/mass/qpos0/kp."""
        keys = self._keys(rng)   # creates the handle on first use
        self.handle.randomize(keys, self._stream())

    def reset(self, rng, mask: Optional[torch.Tensor] = None) -> State:
        keys = self._keys(rng)
        mp = self._ptr(mask, torch.uint8, (self._handle.n,)) if mask is not None else 0
        self._handle.reset(keys, mp, self._stream())
        return self._state()

    def step(self, state: State, action: torch.Tensor) -> State:
        del state  # the library owns the state; the argument keeps the reference's functional signature
        self._handle.step(self._ptr(action, torch.float32, (self._handle.n, self.action_size)), self._stream())
        return self._state()

    def physics_substeps(self, ctrl: Optional[torch.Tensor], n: int) -> Data:
        """mjx_env.step(model, data, ctrl, n) alone (BASELINE config 2)."""
        self.handle.physics_substeps(self._ptr(ctrl, torch.float32, (self._handle.n, self.action_size)) if ctrl is not None else 0, n, self._stream())
        return self._data()

    def forward(self) -> Data:
        self.handle.forward(self._stream())
        return self._data()

    def set_state(self, qpos=None, qvel=None, qacc_warmstart=None) -> None:
        m = self._mj_model
        n = self.handle.n
        self._handle.set_state(self._ptr(qpos, torch.float32, (n, m.nq)), self._ptr(qvel, torch.float32, (n, m.nv)),
                               self._ptr(qacc_warmstart, torch.float32, (n, m.nv)), self._stream())

    def _data(self) -> Data:
        b = self.buffer
        return Data(b("QPOS"), b("QVEL"), b("QACC_WARM"), b("QACC"), b("CTRL"), b("SENSORDATA"), b("ACTUATOR_FORCE"),
                    b("EFC_FORCE"), b("CONTACT_DIST"), b("SITE_XPOS_FEET"))

    def _state(self) -> State:
        b = self.buffer
        met = b("METRICS")
        metrics = {name: met[:, i] for i, name in enumerate(self.METRICS)}
        info = {k: b(v) for k, v in _INFO_BUFFERS.items()}
        la = b("INFO_LAST_ACT")
        info.update(last_act=la[:, 0], last_last_act=la[:, 1], last_last_last_act=la[:, 2],
                    first_obs={"state": b("FIRST_OBS_STATE"), "privileged_state": b("FIRST_OBS_PRIV")})
        return State(self._data(), {"state": b("OBS_STATE"), "privileged_state": b("OBS_PRIV")}, b("REWARD"), b("DONE"), metrics, info)

    # reference helper names kept for callers that used them (base.py:166-222)
    def get_actuator_joints_qpos(self, qpos: torch.Tensor) -> torch.Tensor:
        m = self._mj_model
        return qpos[..., [int(m.jnt_qposadr[m.act_jntid[u]]) for u in range(m.nu)]]

    def get_actuator_joints_qvel(self, qvel: torch.Tensor) -> torch.Tensor:
        m = self._mj_model
        return qvel[..., [int(m.jnt_dofadr[m.act_jntid[u]]) for u in range(m.nu)]]

    def get_gravity(self, data: Data) -> torch.Tensor:      # "upvector" sensor (base.py:234)
        return data.sensordata[:, 9:12]

    def get_gyro(self, data: Data) -> torch.Tensor:
        return data.sensordata[:, 0:3]

    def get_local_linvel(self, data: Data) -> torch.Tensor:
        return data.sensordata[:, 3:6]

    def get_accelerometer(self, data: Data) -> torch.Tensor:
        return data.sensordata[:, 6:9]

    def get_global_angvel(self, data: Data) -> torch.Tensor:
        return data.sensordata[:, 12:15]
