"""Kernel-tuning aid: build liboduck_cuda.so variants with extra -D flags and bench them back to back on one GPU box.

  python tools/variants.py build  name1=-DWPB=7 name2="-DWPB=7 -DODUCK_OPAQUE_IDS" ...
  python tools/variants.py bench  [--steps K --warmup W]      # on the GPU box; prints one line per variant

Variants land in open_duck_playground_b200/csrc/variants/ (git-ignored, travels with gpurun snapshots).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "open_duck_playground_b200", "csrc")
VDIR = os.path.join(CSRC, "variants")
sys.path.insert(0, ROOT)


def build(specs):
    from __graft_entry__ import compile_library
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition("=")
        compile_library(os.path.join(VDIR, f"liboduck_cuda_{name}.so"), flags.split())
        print("built", name)


def bench(extra):
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    for lib in libs:
        env = dict(os.environ, ODUCK_CUDA_LIB=os.path.join(VDIR, lib))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(lib, "FAILED", r.stderr[-400:])
            continue
        j = json.loads(line[-1])
        print(f"{lib:44s} value={j['value']:.4g} ms_per_step={j['ms_per_step']:.4f} kernel_ms={j['roofline'].get('kernel_ms')} e2e={j['e2e']['value']:.4g}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        bench(sys.argv[2:])
