"""The warp-level height-field collider (csrc/oduck_hfcollide.cuh) run on the CPU: tests/emu compiles the device header for
the host and runs one warp as 32 threads, every warp intrinsic an exchange between two barriers.  This checks the LOGIC of
hf_collide (box culls, shared-memory in-place clipping of (triangle, face) pairs, twin masking, manifold pick) against
the oracle without a GPU; the real kernel is compared with the oracle on the B200 box (tests/test_hfield.py -m gpu)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import make_handle
from open_duck_playground_b200 import constants, mjcf
from open_duck_playground_b200.mjcf import CompiledModel
from test_hfield import _dump, _poses

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
CSRC = os.path.join(os.path.dirname(EMU), "..", "open_duck_playground_b200", "csrc")
VARIANTS = {"default": []}


def _lib(name):
    out = os.path.join(EMU, "_build", f"libhf_emu_{name}.so")
    srcs = [os.path.join(EMU, f) for f in ("hf_emu.cpp", "cuda_runtime.h")] + [os.path.join(CSRC, f) for f in ("oduck_hfcollide.cuh", "oduck_ffcollide.cuh", "oduck_device.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", f"-I{EMU}", *VARIANTS[name], os.path.join(EMU, "hf_emu.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.emu_hf_collide.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _emu_contacts(lib, model, q):
    """Contacts of both feet for every pose: [n, 8, 8] = dist, pos[3], normal[3], -."""
    A = model.arrays
    nvt, npl = int(A["foot_nvert"]), int(A["foot_nplane"])
    pnv = np.ascontiguousarray(A["foot_plane_nvert"][:npl], np.int32)
    pv = np.ascontiguousarray(A["foot_plane_vert"][:npl, :8], np.int32)
    data = np.ascontiguousarray(A["hfield_data"], np.float32)
    size = np.asarray(A["hfield_size"][:3], np.float32)
    out = np.zeros((len(q), 8, 8), np.float32)
    for i in range(len(q)):
        xpos, xmat, _, _ = mjcf.world_kinematics(model, q[i].astype(np.float64))
        for k in range(2):
            b = int(A["foot_body"][k])
            vert = np.ascontiguousarray(A["foot_vert"][k][:nvt], np.float32)
            nrm = np.ascontiguousarray(A["foot_plane_normal"][k][:npl], np.float32)
            xp, xm = np.ascontiguousarray(xpos[b], np.float32), np.ascontiguousarray(xmat[b].reshape(9), np.float32)
            cen = np.ascontiguousarray(A["foot_center"][k], np.float32)
            o = np.zeros((4, 8), np.float32)
            lib.emu_hf_collide(xp.ctypes.data, xm.ctypes.data, vert.ctypes.data, nvt, npl, pnv.ctypes.data, pv.ctypes.data, nrm.ctypes.data,
                               cen.ctypes.data, float(A["foot_radius"]), data.shape[0], data.shape[1], size.ctypes.data, data.ctypes.data, o.ctypes.data)
            out[i, 4 * k:4 * k + 4] = o
    return out


@pytest.fixture(scope="module")
def scene(oracle, poly_table):
    model = CompiledModel.load(constants.task_to_blob("rough_terrain_backlash"))
    n = 48
    q = _poses(model, n, 21, tilt=0.08, dz=(-0.004, 0.012))
    h = make_handle(oracle, model, poly_table, n)
    v = np.zeros((n, model.nv), np.float32)
    h.set_state(q.ctypes.data, v.ctypes.data, v.ctypes.data)
    d = _dump(oracle, h)
    ref = np.zeros((n, 8, 8))
    ref[:, :, 0] = d[:, 1120:1128]
    ref[:, :, 1:4] = d[:, 1136:1160].reshape(n, 8, 3)
    ref[:, :, 4:7] = d[:, 2560:2584].reshape(n, 8, 3)
    return model, q, ref


@pytest.fixture(scope="module")
def emulated(scene):
    model, q, _ = scene
    return {name: _emu_contacts(_lib(name), model, q) for name in VARIANTS}


def test_emulated_collider_matches_the_oracle(scene, emulated):
    _, q, ref = scene
    got = emulated["default"].astype(np.float64)
    act_r, act_g = ref[:, :, 0] < 0, got[:, :, 0] < 0
    assert act_r.any(axis=1).mean() > 0.7
    feet_r, feet_g = act_r.reshape(-1, 2, 4), act_g.reshape(-1, 2, 4)
    same = (feet_r == feet_g).all(axis=2)
    assert same.mean() > 0.95                                            # fp32 vs fp64 manifold picks differ in a few feet
    both = act_r & act_g & np.repeat(same, 4, axis=1).reshape(act_r.shape)
    assert both.sum() > 100
    assert np.abs(got[both][:, 0] - ref[both][:, 0]).max() < 5e-6
    # a different pick among near-coincident candidates moves a contact by more than rounding: rare, the rest agree to fp32 rounding
    dpos = np.abs(got[both][:, 1:4] - ref[both][:, 1:4]).max(axis=1)
    dnrm = np.abs(got[both][:, 4:7] - ref[both][:, 4:7]).max(axis=1)
    assert (dpos < 2e-5).mean() > 0.97 and (dnrm < 5e-5).mean() > 0.97
