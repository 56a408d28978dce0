import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from open_duck_playground_b200 import rng as jr
from open_duck_playground_b200.joystick import Joystick
from oracle import oracle_lib
sys.path.insert(0, os.path.dirname(__file__))
from gpu_parity_report import debug_dump, SECT
n = 256
gpu = Joystick("flat_terrain_backlash", device="cuda:0"); ref = Joystick("flat_terrain_backlash", library=oracle_lib.load())
for e in (gpu, ref): e.randomize(jr.split(jr.PRNGKey(11), n))
keys = jr.split(jr.PRNGKey(0), n)
gpu.reset(keys); ref.reset(keys)
# reset leaves qacc_warm = qacc; redo the reset-time forward from zero warmstart for the dump
z = torch.zeros(n, 30)
for e in (gpu, ref): e.set_state(None, None, z if e is ref else z.cuda())
dg, dr = debug_dump(gpu), debug_dump(ref)
err = np.abs(dg[:, 1736:1766] - dr[:, 1736:1766]).max(axis=1)
rel = err / (np.abs(dr[:, 1736:1766]).max(axis=1) + 1e-9)
order = np.argsort(-rel)
print("envs with rel qacc err > 1e-3:", int((rel > 1e-3).sum()), "of", n)
for i in order[:6]:
    print(f"env {i}: abs {err[i]:.3e} rel {rel[i]:.3e}  gpu[costw,costs,alpha,it]={dg[i,2536:2540]}  ref={dr[i,2536:2540]}")
    for k in ("qacc_smooth", "search", "grad", "con_dist", "D_con", "aref_con", "aref_lim"):
        o, ln = SECT[k]; d = np.abs(dg[i, o:o+ln] - dr[i, o:o+ln]); j = d.argmax()
        print(f"     {k:12s} max|d|={d.max():.3e} (gpu {dg[i,o+j]:.6g} ref {dr[i,o+j]:.6g})")
