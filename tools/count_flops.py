"""Algorithmic FLOPs per env-step, counted: the oracle built with ODUCK_COUNT_FLOPS executes the restated
algorithm on a `real` type that counts every + - * / sqrt and transcendental, single-threaded.

    python tools/count_flops.py [task] [n_envs]
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ODUCK_THREADS"] = "1"
import numpy as np
import torch

from open_duck_playground_b200 import capi, rng as jr
from open_duck_playground_b200.joystick import Joystick

task = sys.argv[1] if len(sys.argv) > 1 else "flat_terrain_backlash"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
so = os.path.join(ROOT, "oracle", "liboduck_oracle_count.so")
subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-DODUCK_COUNT_FLOPS", "-pthread", "-shared", "-o", so, os.path.join(ROOT, "oracle", "oduck_oracle.cpp")])
lib = capi.Library(so, is_device=False)
lib.lib.oduck_flop_count.restype = C.c_uint64
env = Joystick(task, library=lib)
env.randomize(jr.split(jr.PRNGKey(2), n))
st = env.reset(jr.split(jr.PRNGKey(0), n))
rs = np.random.default_rng(1)
act = lambda: torch.from_numpy(rs.uniform(-1, 1, (n, 14)).astype(np.float32))
for _ in range(30):                       # let the feet settle into contact: the contact rows are part of the work
    st = env.step(st, act())
lib.lib.oduck_flop_reset()
K = 10
for _ in range(K):
    st = env.step(st, act())
f = lib.lib.oduck_flop_count()
lib.lib.oduck_flop_reset()
env.physics_substeps(None, 10)
fp = lib.lib.oduck_flop_count()
print(f"{task}: {f / (n * K):.0f} flop per env-step (env.step: 10 substeps + obs + rewards), {fp / n:.0f} per 10 physics substeps alone, {fp / (n * 10):.0f} per substep")
