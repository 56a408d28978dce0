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
VARIANTS = {"default": [], "stats": ["-DODUCK_HF_STATS"]}   # stats: counters + the candidate list of the last call exported


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
    return {name: _emu_contacts(_lib(name), model, q) for name in ("default",)}


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


def _candidates_fp32(xp, xm, vert, pnv, pv, nrm, cen, rb, data, size):
    """Candidate list of oracle/oduck_oracle.cpp hfield_convex (lines "for r ... for c ... for q ...": triangles under the bounding
    sphere, down-looking faces, two-buffer Sutherland-Hodgman, points below the triangle plane) restated with numpy float32 scalars in
    the device code's operation order, so that the emulated device collider can be held to it BIT FOR BIT: [nc, 7] = dist, pos, normal."""
    f32 = np.float32
    R = xm.reshape(3, 3).astype(f32)
    V = []
    for v in vert:
        w = np.array([xp[i] + (R[i, 0] * v[0] + R[i, 1] * v[1] + R[i, 2] * v[2]) for i in range(3)], f32)
        V.append(w)
    N = [np.array([(R[i, 0] * n[0] + R[i, 1] * n[1] + R[i, 2] * n[2]) for i in range(3)], f32) for n in nrm]
    C = np.array([xp[i] + (R[i, 0] * cen[0] + R[i, 1] * cen[1] + R[i, 2] * cen[2]) for i in range(3)], f32)
    nrow, ncol = data.shape
    sx, sy, sz = f32(size[0]), f32(size[1]), f32(size[2])
    dx, dy = f32(2.0 * float(size[0]) / (ncol - 1)), f32(2.0 * float(size[1]) / (nrow - 1))
    rb = f32(rb)
    cmin = int(np.floor((C[0] - rb + sx) / dx)); cmax = int(np.floor((C[0] + rb + sx) / dx))
    rmin = int(np.floor((C[1] - rb + sy) / dy)); rmax = int(np.floor((C[1] + rb + sy) / dy))
    cmin = max(cmin, 0); rmin = max(rmin, 0); cmax = min(cmax, ncol - 2); rmax = min(rmax, nrow - 2)
    out = []
    for r in range(rmin, rmax + 1):
        for c in range(cmin, cmax + 1):
            x0, x1, y0, y1 = f32(c) * dx - sx, f32(c + 1) * dx - sx, f32(r) * dy - sy, f32(r + 1) * dy - sy
            H = lambda rr, cc: f32(data[rr, cc]) * sz
            tris = [[(x0, y1, H(r + 1, c)), (x0, y0, H(r, c)), (x1, y1, H(r + 1, c + 1))], [(x0, y0, H(r, c)), (x1, y0, H(r, c + 1)), (x1, y1, H(r + 1, c + 1))]]
            for i in range(2):
                T = [np.array(t, f32) for t in tris[i]]
                top = max(T[0][2], T[1][2], T[2][2])
                if C[2] - rb > top: continue
                e1, e2 = T[1] - T[0], T[2] - T[0]
                n = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]], f32)
                n = (f32(1) / np.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])) * n
                for q in range(len(pnv)):
                    if not (N[q][0] * n[0] + N[q][1] * n[1] + N[q][2] * n[2] < 0): continue
                    poly = [V[pv[q][v]] for v in range(pnv[q])]
                    for e in range(3):
                        if not poly: break
                        r0, r1 = T[e], T[(e + 1) % 3]
                        sd = (r1[1] - r0[1], -(r1[0] - r0[0]))
                        new = []
                        for v in range(len(poly)):
                            p0, p1 = poly[v], poly[(v + 1) % len(poly)]
                            d0 = sd[0] * (p0[0] - r0[0]) + sd[1] * (p0[1] - r0[1]); d1 = sd[0] * (p1[0] - r0[0]) + sd[1] * (p1[1] - r0[1])
                            if d0 <= 0: new.append(p0)
                            if (d0 <= 0) != (d1 <= 0):
                                t = d0 / (d0 - d1)
                                new.append((p0 + t * (p1 - p0)).astype(f32))
                        poly = new
                    for p in poly:
                        dist = n[0] * (p[0] - T[0][0]) + n[1] * (p[1] - T[0][1]) + n[2] * (p[2] - T[0][2])
                        if not dist < 0: continue
                        out.append([dist, *(p - f32(0.5) * dist * n), *n])
    return np.array(out, f32).reshape(-1, 7)


def test_emulated_candidate_lists_equal_the_fp32_restatement_bit_for_bit():
    """The group-parallel clipping (hf_clip_pass: one polygon per 8 lanes, prefix-sum compaction) must produce the serial
    algorithm's candidates: same points, same order, same bits -- the culls (cell box, height, plane side, face box) may only drop
    (triangle, face) pairs that yield no candidate.  Poses: resting, sunk, tilted, fallen, hovering."""
    import ctypes as C
    lib = _lib("stats")
    lib.emu_hf_stats.restype = C.POINTER(C.c_longlong)
    lib.emu_hf_last_candidates.restype = C.POINTER(C.c_float)
    model = CompiledModel.load(constants.task_to_blob("rough_terrain_backlash"))
    A = model.arrays
    nvt, npl = int(A["foot_nvert"]), int(A["foot_nplane"])
    pnv = np.ascontiguousarray(A["foot_plane_nvert"][:npl], np.int32)
    pv = np.ascontiguousarray(A["foot_plane_vert"][:npl, :8], np.int32)
    data = np.ascontiguousarray(A["hfield_data"], np.float32)
    size = np.asarray(A["hfield_size"][:3], np.float32)
    q = np.concatenate([_poses(model, 3, 21, tilt=0.08, dz=(-0.004, 0.012)), _poses(model, 3, 5, tilt=0.3, dz=(-0.03, 0.01)),
                        _poses(model, 2, 7, tilt=1.2, dz=(-0.06, -0.02)), _poses(model, 2, 9, tilt=0.05, dz=(0.008, 0.014))])
    st = lib.emu_hf_stats()
    counts = []
    for i in range(len(q)):
        xpos, xmat, _, _ = mjcf.world_kinematics(model, q[i].astype(np.float64))
        for k in range(2):
            b = int(A["foot_body"][k])
            vert = np.ascontiguousarray(A["foot_vert"][k][:nvt], np.float32)
            nrm = np.ascontiguousarray(A["foot_plane_normal"][k][:npl], np.float32)
            xp, xm = np.ascontiguousarray(xpos[b], np.float32), np.ascontiguousarray(xmat[b].reshape(9), np.float32)
            cen = np.ascontiguousarray(A["foot_center"][k], np.float32)
            o = np.zeros((4, 8), np.float32)
            before = st[3]
            lib.emu_hf_collide(xp.ctypes.data, xm.ctypes.data, vert.ctypes.data, nvt, npl, pnv.ctypes.data, pv.ctypes.data, nrm.ctypes.data,
                               cen.ctypes.data, float(A["foot_radius"]), data.shape[0], data.shape[1], size.ctypes.data, data.ctypes.data, o.ctypes.data)
            nc = st[3] - before
            got = np.ctypeslib.as_array(lib.emu_hf_last_candidates(), shape=(512, 8))[:min(nc, 512), :7]
            want = _candidates_fp32(xp, xm, vert, pnv, pv, nrm, cen, float(A["foot_radius"]), data, size)
            assert len(want) == nc and np.array_equal(want[:512], got), (i, k, len(want), nc)
            counts.append(nc)
    assert max(counts) > 150 and min(counts) == 0 and sum(c > 0 for c in counts) >= 10      # deep, hovering and resting feet all occurred
