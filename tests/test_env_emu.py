"""The env side of the CUDA library run on the CPU.  tests/emu/oduck_emu.cpp compiles csrc/oduck_cuda.cu itself (oduck_create,
k_randomize, k_reset, k_step, k_physics and the buffer views) for the host: every kernel launch runs block by block as 32
threads, one per lane, every warp intrinsic an exchange between two barriers.  The result is a host library with the C-ABI of
include/oduck.h, so the reference-facing classes drive it like the oracle and the GPU parity checks run here on a few envs:
key streams and counters bit-exact, physics / obs / rewards within the fp32-vs-fp64 tolerances.  This exercises the LOGIC of
the device code without a GPU; the kernels themselves are compared with the oracle on the B200 box (-m gpu)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from open_duck_playground_b200 import capi, rng as jr
from open_duck_playground_b200.joystick import Joystick
from open_duck_playground_b200.standing import Standing

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
CSRC = os.path.join(os.path.dirname(EMU), "..", "open_duck_playground_b200", "csrc")


def build_emu_library():
    out = os.path.join(EMU, "_build", "liboduck_emu.so")
    srcs = [os.path.join(EMU, f) for f in ("oduck_emu.cpp", "cuda_runtime.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", f"-I{EMU}", "-DWPB=1", "-x", "c++", os.path.join(EMU, "oduck_emu.cpp"), "-o", out])
    return capi.Library(out, is_device=False)


@pytest.fixture(scope="module")
def emu_lib():
    return build_emu_library()


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _rows_ok(a, b, atol, rtol=0.0):
    a, b = _np(a).reshape(len(a), -1), _np(b).reshape(len(b), -1)
    return np.abs(a - b).max(axis=1) <= atol + rtol * np.abs(b).max(axis=1)


def _pair(cls, task, emu_lib, oracle, n):
    emu, ref = cls(task, library=emu_lib), cls(task, library=oracle)
    for e in (emu, ref):
        e.randomize(jr.split(jr.PRNGKey(11), n))
    keys = jr.split(jr.PRNGKey(0), n)
    return emu, ref, emu.reset(keys), ref.reset(keys)


def _sync(emu, ref):
    f = lambda name: torch.from_numpy(ref.buffer(name).numpy().astype(np.float32))   # noqa: E731
    emu.set_state(f("QPOS"), f("QVEL"), f("QACC_WARM"))


def _step_and_compare(emu, ref, sg, sr, steps, seed, min_ok=1.0):
    rs = np.random.default_rng(seed)
    n = emu.handle.n
    for t in range(steps):
        act = torch.from_numpy(rs.uniform(-1, 1, (n, emu.action_size)).astype(np.float32))
        _sync(emu, ref)
        sg, sr = emu.step(sg, act), ref.step(sr, act)
        for name in ("INFO_RNG", "INFO_STEP", "INFO_STEPS", "INFO_PUSH_STEP", "INFO_IMITATION_I"):
            assert np.array_equal(emu.buffer(name).numpy(), ref.buffer(name).numpy()), (t, name)
        ok = (_rows_ok(sg.data.qpos, sr.data.qpos, 1e-4) & _rows_ok(sg.data.qvel, sr.data.qvel, 2e-3, 1e-3) & _rows_ok(sg.reward[:, None], sr.reward[:, None], 2e-4) &
              _rows_ok(emu.buffer("METRICS"), ref.buffer("METRICS"), 1e-3, 2e-3) & _rows_ok(sg.obs["state"], sr.obs["state"], 2e-3, 2e-3) &
              _rows_ok(sg.obs["privileged_state"], sr.obs["privileged_state"], 2e-3, 2e-3))
        assert ok.mean() >= min_ok, (t, ok)
        for k in ("command", "motor_targets", "action_history", "feet_air_time", "last_contact", "push"):
            assert _rows_ok(sg.info[k].reshape(n, -1), sr.info[k].reshape(n, -1), 1e-5)[ok].all(), (t, k)
        assert np.array_equal(_np(sg.done)[ok], _np(sr.done)[ok])
    return sg, sr


def test_emulated_randomize_reset_step_joystick(emu_lib, oracle):
    n = 8
    emu, ref, sg, sr = _pair(Joystick, "flat_terrain_backlash", emu_lib, oracle, n)
    m = emu.mj_model
    assert np.array_equal(emu.buffer("INFO_RNG").numpy(), ref.buffer("INFO_RNG").numpy())
    assert np.array_equal(emu.buffer("INFO_PUSH_INTERVAL").numpy(), ref.buffer("INFO_PUSH_INTERVAL").numpy())
    assert _rows_ok(emu.buffer("DR_PARAMS")[:, : m.nbody], ref.buffer("DR_PARAMS")[:, 1:1 + m.nbody], 1e-6).all()      # randomised masses
    assert _rows_ok(sg.data.qpos, sr.data.qpos, 1e-6).all() and _rows_ok(sg.data.qvel, sr.data.qvel, 1e-7).all()
    assert _rows_ok(sg.info["command"], sr.info["command"], 1e-6).all()
    assert _rows_ok(sg.info["current_reference_motion"], sr.info["current_reference_motion"], 3e-4).all()
    assert _rows_ok(sg.obs["state"], sr.obs["state"], 2e-3, 2e-3).mean() >= 0.85                                        # accelerometer follows qacc of the deep reset contact
    _step_and_compare(emu, ref, sg, sr, steps=2, seed=2)


def test_emulated_standing_step(emu_lib, oracle):
    emu, ref, sg, sr = _pair(Standing, "flat_terrain_backlash", emu_lib, oracle, 6)
    assert sg.obs["state"].shape == (6, 85)
    _step_and_compare(emu, ref, sg, sr, steps=1, seed=3)


def test_emulated_height_field_step(emu_lib, oracle):
    emu, ref, sg, sr = _pair(Joystick, "rough_terrain_backlash", emu_lib, oracle, 8)
    _step_and_compare(emu, ref, sg, sr, steps=1, seed=4, min_ok=0.75)       # manifold branch flips, see tests/test_hfield.py


@pytest.mark.parametrize("task", ["flat_terrain_backlash", "flat_terrain"])
def test_domain_randomize_every_column_matches_oracle(emu_lib, oracle, task):
    """A14 column by column (common/randomize.py:43-95): friction, frictionloss, armature, torso COM, masses, qpos0, kp -- the
    device code's k_randomize against the oracle, the same comparison tests/test_parity_gpu.py makes on the GPU."""
    from test_parity_gpu import _dr_columns
    emu, ref, _, _ = _pair(Joystick, task, emu_lib, oracle, 6)
    cols = _dr_columns(emu, ref)
    assert set(cols) == {"geom_friction0", "body_mass", "body_ipos[1]", "dof_frictionloss", "dof_armature", "qpos0", "actuator kp"}
    m = emu.mj_model
    for name, (g, r) in cols.items():
        g, r = _np(g), _np(r)
        assert (np.abs(g - r) <= 1e-6 + 3e-7 * np.abs(r)).all(), name
        assert np.abs(r).max() > 0 and (name in ("dof_frictionloss",) or np.std(r, axis=0).max() > 0), f"{name}: column is not randomised"
    # and the columns really are the randomised ones: mass within +-10 % (+-0.1 kg on the torso) of the model's
    mass = _np(cols["body_mass"][1])
    mm = np.asarray(m.body_mass[:m.nbody])
    heavy = np.flatnonzero(mm > 0)[1:]                                   # all but the torso (body 1), which also gets +-0.1 kg
    assert (np.abs(mass[:, heavy] / mm[heavy] - 1.0) <= 0.1 + 1e-9).all()
    kp = _np(cols["actuator kp"][1])
    assert (np.abs(kp / np.asarray(m.act_kp[:m.nu]) - 1.0) <= 0.1 + 1e-9).all()


def test_step_kernel_writes_the_transition_into_the_rollout_sink(emu_lib):
    """A17, device code on CPU threads: k_step with a sink attached (oduck_step_into_sink, the env half of oduck_rollout_step)
    stores reward / done / truncation of slot t and the new observations of slot t + 1 at the handle's env offset -- through an
    episode end, where the stored observation is the auto-reset one -- and fills slot 0 from the current observations at t = 0."""
    import ctypes as C
    n, T, off, wide = 3, 3, 2, 6
    env = Joystick("flat_terrain_backlash", library=emu_lib, config_overrides={"episode_length": 2})
    st = env.reset(jr.split(jr.PRNGKey(3), n))
    buf = {"obs_p": torch.zeros(T + 1, wide, 101), "obs_v": torch.zeros(T + 1, wide, 212), "raw": torch.zeros(T, wide, 14), "logp": torch.zeros(T, wide),
           "reward": torch.full((T, wide), -1.0), "done": torch.full((T, wide), -1.0), "trunc": torch.full((T, wide), -1.0)}
    from open_duck_playground_b200 import ppo
    ppo.attach_rollout_sink(env, buf, env_offset=off)
    L = emu_lib.lib
    L.oduck_step_into_sink.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    obs0 = st.obs["state"].clone()
    rs = np.random.default_rng(0)
    for t in range(T):
        act = torch.from_numpy(rs.uniform(-1, 1, (n, 14)).astype(np.float32))
        emu_lib.check(L.oduck_step_into_sink(env.handle.h, act.data_ptr(), t, None))
        s = env._state()
        assert torch.equal(buf["obs_p"][t + 1, off:off + n], s.obs["state"]) and torch.equal(buf["obs_v"][t + 1, off:off + n], s.obs["privileged_state"])
        assert torch.equal(buf["reward"][t, off:off + n], s.reward) and torch.equal(buf["done"][t, off:off + n], s.done)
        assert torch.equal(buf["trunc"][t, off:off + n], s.info["truncation"])
    assert torch.equal(buf["obs_p"][0, off:off + n], obs0)
    assert (buf["trunc"][1, off:off + n] == 1).all() and torch.equal(buf["obs_p"][2, off:off + n], env.buffer("FIRST_OBS_STATE"))   # episode_length 2
    assert (buf["reward"][:, :off] == -1).all() and (buf["reward"][:, off + n:] == -1).all() and (buf["obs_p"][:, :off] == 0).all()
