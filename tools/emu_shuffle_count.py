"""Warp exchanges (SHFL / VOTE / REDUX of the real kernel) per phase of one physics substep, counted on the CPU emulation of the
device code (tests/emu): where the shuffle / shared-memory pipe's load comes from, without a GPU.
    python tools/emu_shuffle_count.py [nsub]
Phases (PHASE_SYNC points of csrc/oduck_physics.cuh): 0 kinematics..M | 1 com_vel, rne, smooth, factor M | 2 collision, rows,
warm start | 3 gradient, Hessian, factor H | 4 line search | 5 outputs, euler."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from conftest import make_handle
from open_duck_playground_b200 import constants
from open_duck_playground_b200.mjcf import CompiledModel
from open_duck_playground_b200.poly_reference_motion import PolyTable
from oracle import oracle_lib
from test_step_emu import load_emu, _run_pair

nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 10
emu = load_emu()
model = CompiledModel.load(constants.task_to_blob("flat_terrain_backlash"))
poly = PolyTable.load(constants.POLY_BLOB)
cnt = (C.c_longlong * 16)()
emu.emu_exchange_counts(cnt, 1)
n = 8
_run_pair(emu, oracle_lib.load(), model, poly, n=n, nsub=nsub, seed=7)
emu.emu_exchange_counts(cnt, 1)
tot = sum(cnt)
names = {0: "kinematics, com, cdof, crb, M", 1: "com_vel, rne, smooth force, factor M + solve", 2: "collision, constraint rows, warm start", 3: "gradient, Hessian, factor H + solve",
         4: "line search", 5: "outputs (last substep), euler", 15: "load_env / store (outside the substep)"}
for k in sorted(names):
    print(f"phase {k:2d}  {cnt[k] / (n * nsub):8.1f} exchanges per substep  {100.0 * cnt[k] / tot:5.1f} %   {names[k]}")
print(f"total    {tot / (n * nsub):8.1f} per substep")
