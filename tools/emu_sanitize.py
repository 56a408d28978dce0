"""Memory checks of the device code without a GPU: builds tests/emu/oduck_emu.cpp (csrc/oduck_cuda.cu on CPU threads) with
-fsanitize=bounds and with -fsanitize=address and drives reset / step / physics of every task flavour through it.

    python tools/emu_sanitize.py            # prints one line per configuration; sanitizer reports go to stderr
    python tools/emu_sanitize.py -DODUCK_CHOL_LDL -DODUCK_SYMV_UNROLL -DODUCK_ANC_PIPE     # the same for a variant build

A CPU analogue of `compute-sanitizer --tool memcheck` for the kernels' indexing (shared-memory records are plain arrays in the
emulation, so UBSan's bounds checker sees every `s.con[lane][k]`-style access; ASan guards the HBM-side buffers)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
DRIVER = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np, torch
from open_duck_playground_b200 import capi, rng as jr
from open_duck_playground_b200.joystick import Joystick
from open_duck_playground_b200.standing import Standing
import test_reward_library as T
emu = capi.Library(%(lib)r, False)
for cls, task, cfg in ((Joystick, "flat_terrain_backlash", None), (Joystick, "flat_terrain", None), (Joystick, "rough_terrain_backlash", None),
                       (Standing, "flat_terrain_backlash", None), (Joystick, "flat_terrain_backlash", T.library_config())):
    n = 4
    g = cls(task, library=emu, config=cfg) if cfg is not None else cls(task, library=emu)
    g.randomize(jr.split(jr.PRNGKey(11), n)); st = g.reset(jr.split(jr.PRNGKey(0), n))
    rs = np.random.default_rng(2)
    for t in range(2):
        st = g.step(st, torch.from_numpy(rs.uniform(-1, 1, (n, 14)).astype(np.float32)))
    g.physics_substeps(None, 2); g.forward()
    print(%(tag)r, cls.__name__, task, "library terms" if cfg is not None else "", "ok", flush=True)
'''


def main():
    rc = 0
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    suffix = "".join("_" + a[3:].lower() for a in extra)
    for tag, flag in (("ubsan-bounds", "-fsanitize=bounds"), ("asan", "-fsanitize=address")):
        lib = os.path.join(EMU, "_build", f"liboduck_emu_{tag}{suffix}.so")
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fPIC", "-shared", f"-I{EMU}", "-DWPB=1", flag, *extra, "-x", "c++", os.path.join(EMU, "oduck_emu.cpp"), "-o", lib])
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
        if tag == "asan":
            env["LD_PRELOAD"] = subprocess.check_output(["g++", "-print-file-name=libasan.so"], text=True).strip()
        r = subprocess.run([sys.executable, "-c", DRIVER % {"root": ROOT, "lib": lib, "tag": tag}], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        bad = [l for l in r.stderr.splitlines() if "runtime error" in l or "AddressSanitizer" in l]
        if r.returncode != 0 or bad:
            rc = 1
            sys.stderr.write(r.stderr[-4000:])
        print(f"{tag}: {'CLEAN' if not bad and r.returncode == 0 else 'REPORTS'}")
    sys.exit(rc)


if __name__ == "__main__":
    main()
