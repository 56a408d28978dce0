"""Phase-by-phase comparison of liboduck_cuda against the CPU oracle (run under gpurun)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from open_duck_playground_b200 import rng as jr  # noqa: E402
from open_duck_playground_b200.joystick import Joystick  # noqa: E402
from oracle import oracle_lib  # noqa: E402

SECT = {"M": (0, 1024), "qfrc_bias": (1024, 32), "qfrc_smooth": (1056, 32), "qacc_smooth": (1088, 32), "con_dist": (1120, 12),
        "con_pos": (1136, 36), "D_fric": (1184, 32), "D_lim": (1216, 32), "D_con": (1248, 12), "aref_fric": (1264, 32),
        "aref_lim": (1296, 32), "aref_con": (1328, 48), "search": (1376, 32), "grad": (1408, 32), "xpos": (1440, 96),
        "com": (1536, 3), "cdof": (1540, 192), "qacc": (1736, 32), "cost_w_s_alpha_it": (2536, 4), "H": (4096, 1024)}


def cmp(name, a, b, act=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b)
    i = np.unravel_index(np.argmax(d), d.shape)
    print(f"  {name:22s} max|d|={d.max():.3e} at {i} (gpu={a[i]:.6g} ref={b[i]:.6g})  max|ref|={np.abs(b).max():.3e}")


def debug_dump(env):
    h = env.handle
    L = h.L.lib
    stride = L.oduck_debug_stride()
    if h.L.is_device:
        out = np.zeros((h.n, stride), np.float32)
        L.oduck_debug_forward.argtypes = [C.c_void_p, C.c_void_p]
        h.L.check(L.oduck_debug_forward(h.h, out.ctypes.data))
    else:
        out = np.zeros((h.n, stride), np.float64)
        L.oduck_debug_forward.argtypes = [C.c_void_p, C.c_void_p]
        h.L.check(L.oduck_debug_forward(h.h, out.ctypes.data))
    return out


def main():
    n = int(os.environ.get("N", 64))
    task = os.environ.get("TASK", "flat_terrain_backlash")
    gpu = Joystick(task, device="cuda:0")
    ref = Joystick(task, library=oracle_lib.load())
    keys = jr.split(jr.PRNGKey(0), n)
    for e in (gpu, ref):
        e.randomize(jr.split(jr.PRNGKey(1), n))
    print("== randomize")
    a, b = gpu.buffer("DR_PARAMS").cpu().numpy(), ref.buffer("DR_PARAMS").numpy()
    m = gpu.mj_model
    cmp("mass", a[:, :m.nbody], b[:, 1:1 + m.nbody])
    sg, sr = gpu.reset(keys), ref.reset(keys)
    torch.cuda.synchronize()
    print("== reset")
    for name in ("QPOS", "QVEL", "QACC_WARM", "SENSORDATA", "CONTACT_DIST", "EFC_FORCE", "ACTUATOR_FORCE", "OBS_STATE", "OBS_PRIV",
                 "INFO_COMMAND", "INFO_PUSH_INTERVAL", "INFO_REF_MOTION", "INFO_RNG"):
        cmp(name, gpu.buffer(name).cpu().numpy(), ref.buffer(name).numpy())
    print("== debug forward (state after reset)")
    dg, dr = debug_dump(gpu), debug_dump(ref)
    for k, (o, ln) in SECT.items():
        cmp(k, dg[:, o:o + ln], dr[:, o:o + ln])
    print("== 10 physics substeps, random ctrl")
    rs = np.random.default_rng(0)
    ctrl = (m.key_ctrl[:m.nu] + 0.25 * rs.uniform(-1, 1, (n, m.nu))).astype(np.float32)
    gpu.physics_substeps(torch.from_numpy(ctrl).cuda(), 10)
    ref.physics_substeps(torch.from_numpy(ctrl), 10)
    torch.cuda.synchronize()
    for name in ("QPOS", "QVEL", "QACC_WARM", "SENSORDATA", "CONTACT_DIST", "EFC_FORCE"):
        cmp(name, gpu.buffer(name).cpu().numpy(), ref.buffer(name).numpy())
    print("== env.step x 5, random actions (states re-synchronised from the oracle before every step)")
    sg, sr = gpu.reset(keys), ref.reset(keys)
    for t in range(5):
        act = rs.uniform(-1, 1, (n, m.nu)).astype(np.float32)
        gpu.set_state(torch.from_numpy(ref.buffer("QPOS").numpy().astype(np.float32)), torch.from_numpy(ref.buffer("QVEL").numpy().astype(np.float32)),
                      torch.from_numpy(ref.buffer("QACC_WARM").numpy().astype(np.float32)))
        sg, sr = gpu.step(sg, torch.from_numpy(act).cuda()), ref.step(sr, torch.from_numpy(act))
        torch.cuda.synchronize()
        print(f" step {t}")
        for name in ("QPOS", "QVEL", "OBS_STATE", "OBS_PRIV", "REWARD", "DONE", "METRICS", "INFO_COMMAND", "INFO_RNG", "INFO_MOTOR_TARGETS",
                     "INFO_ACTION_HISTORY", "INFO_FEET_AIR_TIME", "INFO_SWING_PEAK", "INFO_STEP", "INFO_PUSH", "INFO_IMITATION_PHASE"):
            cmp(name, gpu.buffer(name).cpu().numpy(), ref.buffer(name).numpy())
    print("== timing")
    for nn in (4096, 8192, 32768):
        e = Joystick(task, device="cuda:0")
        e.randomize(jr.split(jr.PRNGKey(1), nn))
        e.reset(jr.split(jr.PRNGKey(0), nn))
        act = torch.zeros(nn, m.nu, device="cuda")
        for _ in range(3):
            e.step(None, act)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            e.step(None, act)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 20
        print(f"  N={nn}: {ms:.3f} ms / env.step  -> {nn / ms * 1e3:.3e} env-steps/s")
        del e


if __name__ == "__main__":
    main()
