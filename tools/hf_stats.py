"""Workload statistics of the height-field collider on realistic states (CPU only): the oracle library rolls the rough-terrain env
out under a random policy, the emulated device collider (tests/emu, -DODUCK_HF_STATS) counts, per foot and substep-equivalent call,
the terrain triangles that survive the culls, the (triangle, face) pairs it clips and the candidates it lists."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from open_duck_playground_b200 import constants, mjcf, rng as jr  # noqa: E402
from open_duck_playground_b200.joystick import Joystick  # noqa: E402
from open_duck_playground_b200.mjcf import CompiledModel  # noqa: E402
from oracle import oracle_lib  # noqa: E402


def main(n=32, steps=40, sample=(0, 1, 2, 4, 9, 19, 39)):
    out = "/tmp/hf/libhf_stats.so"
    os.makedirs("/tmp/hf", exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", f"-I{ROOT}/tests/emu", "-DODUCK_HF_STATS", *[a for a in sys.argv[1:] if a.startswith("-D")],
                           f"{ROOT}/tests/emu/hf_emu.cpp", "-o", out])
    lib = C.CDLL(out)
    lib.emu_hf_collide.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_hf_stats.restype = C.POINTER(C.c_longlong)
    st = lib.emu_hf_stats()
    env = Joystick("rough_terrain_backlash", library=oracle_lib.load())
    env.randomize(jr.split(jr.PRNGKey(1), n))
    s = env.reset(jr.split(jr.PRNGKey(0), n))
    model = CompiledModel.load(constants.task_to_blob("rough_terrain_backlash"))
    A = model.arrays
    nvt, npl = int(A["foot_nvert"]), int(A["foot_nplane"])
    pnv = np.ascontiguousarray(A["foot_plane_nvert"][:npl], np.int32)
    pv = np.ascontiguousarray(A["foot_plane_vert"][:npl, :8], np.int32)
    data = np.ascontiguousarray(A["hfield_data"], np.float32)
    size = np.asarray(A["hfield_size"][:3], np.float32)
    g = torch.Generator().manual_seed(0)
    from open_duck_playground_b200 import ppo
    torch.manual_seed(0)
    weights = ppo.PolicyWeights(ppo.MLP([101, 512, 256, 128, 28]), 101, torch.device("cpu"))     # the bench's actor: random weights, sampled actions
    policy = "--mlp" in sys.argv
    rows = []
    for t in range(steps):
        if policy:
            keys = torch.from_numpy(jr.split(jr.PRNGKey(2000 + t), n).view(np.int32).copy())
            act, _, _ = ppo.policy_forward(env, weights, keys, False)
            s = env.step(s, act)
        else:
            s = env.step(s, torch.rand(n, 14, generator=g) * 2 - 1)
        if t in sample:
            print(f"step {t + 1}: done {float(s.done.mean()):.2f}  base z mean {float(s.data.qpos[:, 2].mean()):.3f}  min {float(s.data.qpos[:, 2].min()):.3f}")
        if t not in sample:
            continue
        q = s.data.qpos.numpy()
        for i in range(n):
            xpos, xmat, _, _ = mjcf.world_kinematics(model, q[i].astype(np.float64))
            for k in range(2):
                b = int(A["foot_body"][k])
                vert = np.ascontiguousarray(A["foot_vert"][k][:nvt], np.float32)
                nrm = np.ascontiguousarray(A["foot_plane_normal"][k][:npl], np.float32)
                xp, xm = np.ascontiguousarray(xpos[b], np.float32), np.ascontiguousarray(xmat[b].reshape(9), np.float32)
                cen = np.ascontiguousarray(A["foot_center"][k], np.float32)
                o = np.zeros((4, 8), np.float32)
                before = [st[j] for j in range(8)]
                lib.emu_hf_collide(xp.ctypes.data, xm.ctypes.data, vert.ctypes.data, nvt, npl, pnv.ctypes.data, pv.ctypes.data, nrm.ctypes.data,
                                   cen.ctypes.data, float(A["foot_radius"]), data.shape[0], data.shape[1], size.ctypes.data, data.ctypes.data, o.ctypes.data)
                rows.append([st[j] - before[j] for j in range(8)] + [int((o[:, 0] < 0).sum()), t])
    r = np.array(rows)
    names = ["triangles (height cull)", "triangles (plane-side cull)", "pairs", "candidates", "in-threshold", "s5", "s6", "s7", "contacts"]
    print(f"{len(r)} foot-collisions, nvert={nvt} nplane={npl}")
    for t in sorted(set(r[:, -1])):
        rr = r[r[:, -1] == t]
        print(f"after control step {t + 1}:")
        for j, nm in enumerate(names):
            c = rr[:, j]
            if c.any():
                print(f"  {nm:30s} mean {c.mean():7.2f}  median {np.median(c):6.1f}  p90 {np.percentile(c, 90):6.1f}  max {c.max():5d}  zero {np.mean(c == 0):.2f}")


if __name__ == "__main__":
    main()
