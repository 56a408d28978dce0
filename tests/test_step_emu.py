"""The physics substep of the CUDA library run on the CPU: tests/emu compiles forward_euler (csrc/oduck_physics.cuh, the device
code inside k_physics / k_step) for the host and runs one warp as 32 threads.  fp32 device logic vs the fp64 oracle on shared
states, without a GPU -- flat floor, foot-foot rare path, height-field instantiation.  The kernels themselves are compared
with the oracle on the B200 box (tests/test_parity_gpu.py, tests/test_hfield.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import make_handle
from open_duck_playground_b200 import capi, constants
from open_duck_playground_b200.mjcf import CompiledModel

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
CSRC = os.path.join(os.path.dirname(EMU), "..", "open_duck_playground_b200", "csrc")


def build_emu(name, source, flags=()):
    out = os.path.join(EMU, "_build", f"lib{name}.so")
    srcs = [os.path.join(EMU, source), os.path.join(EMU, "cuda_runtime.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", f"-I{EMU}", *flags, os.path.join(EMU, source), "-o", out])
    return C.CDLL(out)


def load_emu(name="step_emu", flags=()):
    lib = build_emu(name, "step_emu.cpp", flags)
    lib.emu_physics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    st = (C.c_int * 16)()
    lib.emu_strides(st)
    lib.S = dict(zip(["PHYS", "DR", "OUT", "QVEL", "QACCW", "CTRL", "O_QACC", "O_SENS", "O_EFC", "O_CDIST", "O_AFRC"], list(st)))
    return lib


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def _run_pair(emu, oracle, model, poly, n, nsub, seed, warm_steps=3, mutate=None):
    """Bring n oracle envs into contact-rich states, then advance the SAME states by nsub substeps in the oracle and in the
    emulated device code.  Returns dicts of arrays (emulated, oracle)."""
    rs = np.random.default_rng(seed)
    h = make_handle(oracle, model, poly, n)
    keys = np.stack([np.zeros(n, np.uint32), np.arange(n, dtype=np.uint32) + seed], 1)
    h.reset(keys.ctypes.data)
    for _ in range(warm_steps):
        act = (0.3 * rs.uniform(-1, 1, (n, model.nu))).astype(np.float32)
        h.step(act.ctypes.data)
    if mutate is not None:
        mutate(h)
    S = emu.S
    q0, v0, w0 = (h.buffer_numpy(k).astype(np.float32).copy() for k in ("QPOS", "QVEL", "QACC_WARM"))
    ctrl = (model.key_ctrl[: model.nu] + 0.1 * rs.uniform(-1, 1, (n, model.nu))).astype(np.float32)
    h.set_state(q0.ctypes.data, v0.ctypes.data, w0.ctypes.data)          # both sides start from the fp32-rounded state
    h.physics_substeps(ctrl.ctypes.data, nsub)
    ref = {k: h.buffer_numpy(k).astype(np.float64).copy() for k in ("QPOS", "QVEL", "QACC", "EFC_FORCE", "SENSORDATA", "CONTACT_DIST", "ACTUATOR_FORCE")}
    ms = capi.model_to_struct(model)
    nefc = ref["EFC_FORCE"].shape[1]
    got = {k: np.zeros_like(v) for k, v in ref.items()}
    for i in range(n):
        phys = np.zeros(S["PHYS"], np.float32)
        phys[: model.nq] = q0[i]; phys[S["QVEL"]: S["QVEL"] + model.nv] = v0[i]; phys[S["QACCW"]: S["QACCW"] + model.nv] = w0[i]
        out = np.zeros(S["OUT"], np.float32)
        c = np.ascontiguousarray(ctrl[i])
        assert emu.emu_physics(C.addressof(ms), nsub, 1, phys.ctypes.data, c.ctypes.data, out.ctypes.data) == 0
        got["QPOS"][i] = phys[: model.nq]; got["QVEL"][i] = phys[S["QVEL"]: S["QVEL"] + model.nv]
        got["QACC"][i] = out[S["O_QACC"]: S["O_QACC"] + model.nv]; got["EFC_FORCE"][i] = out[S["O_EFC"]: S["O_EFC"] + nefc]
        got["SENSORDATA"][i] = out[S["O_SENS"]: S["O_SENS"] + 24]; got["CONTACT_DIST"][i] = out[S["O_CDIST"]: S["O_CDIST"] + 12]
        got["ACTUATOR_FORCE"][i] = out[S["O_AFRC"]: S["O_AFRC"] + model.nu]
    return got, ref


def _check(got, ref, min_ok):
    """Per-env norm-wise agreement with the GPU parity tolerances (tests/test_parity_gpu.py); envs whose active contact set
    differs (fp32 vs fp64 branch) are allowed up to 1 - min_ok."""
    def rows(a, b, atol, rtol):
        return np.abs(a - b).max(axis=1) <= atol + rtol * np.abs(b).max(axis=1)
    ok = (rows(got["QPOS"], ref["QPOS"], 1e-4, 0) & rows(got["QVEL"], ref["QVEL"], 2e-3, 1e-3) & rows(got["QACC"], ref["QACC"], 1e-3, 2e-3) &
          rows(got["EFC_FORCE"], ref["EFC_FORCE"], 1e-3, 1e-2) & rows(got["SENSORDATA"], ref["SENSORDATA"], 2e-3, 2e-3) &
          rows(got["ACTUATOR_FORCE"], ref["ACTUATOR_FORCE"], 2e-3, 0))
    assert ok.mean() >= min_ok, (ok, np.abs(got["QACC"] - ref["QACC"]).max(axis=1))
    same_set = ((got["CONTACT_DIST"] < 0) == (ref["CONTACT_DIST"] < 0)).all(axis=1)
    assert same_set.mean() >= min_ok
    return ok


def test_emulated_substeps_match_the_oracle_flat(emu, oracle, model_backlash, poly_table):
    got, ref = _run_pair(emu, oracle, model_backlash, poly_table, n=24, nsub=10, seed=100)          # one control step of physics
    assert (ref["CONTACT_DIST"][:, :8] < 0).any(axis=1).mean() >= 0.4     # many states end in floor contact
    _check(got, ref, min_ok=0.95)


def test_emulated_substeps_match_the_oracle_hfield(emu, oracle, poly_table):
    model = CompiledModel.load(constants.task_to_blob("rough_terrain_backlash"))
    got, ref = _run_pair(emu, oracle, model, poly_table, n=24, nsub=10, seed=200)
    assert (ref["CONTACT_DIST"][:, :8] < 0).any(axis=1).mean() >= 0.4
    _check(got, ref, min_ok=0.8)


def test_emulated_foot_foot_path(emu, oracle, model_backlash, poly_table):
    """Poses with the feet pressed together: the FF = true re-run of the substep (csrc/oduck_ffcollide.cuh)."""
    from test_oracle_physics import _ff_poses

    def press(h):
        q = _ff_poses(model_backlash, h.n, 3)
        q[:, 2] = 0.6                                                   # in the air: only the foot-foot contacts act
        v = np.zeros((h.n, model_backlash.nv), np.float32)
        h.set_state(q.ctypes.data, v.ctypes.data, v.ctypes.data)
    got, ref = _run_pair(emu, oracle, model_backlash, poly_table, n=32, nsub=2, seed=300, warm_steps=0, mutate=press)
    hit = (ref["CONTACT_DIST"][:, 8:12] < 0).any(axis=1)
    assert hit.sum() >= 3                                               # the rare path is exercised
    ok = _check(got, ref, min_ok=0.9)
    assert ok[hit].mean() >= 0.75
