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
