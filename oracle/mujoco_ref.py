"""Hook for the REAL reference physics (TEST INFRASTRUCTURE, like everything under oracle/; imported only by tests/ and by
bench.py's ``--impl reference`` / ``cpu_baseline`` legs).

The reference's CPU path is plain MuJoCo: ``mujoco.mj_step(model, data)`` in a loop
(/root/reference/playground/open_duck_mini_v2/mujoco_infer.py:170, model loaded at mujoco_infer_base.py:19); ten of them are one
``env.step`` (joystick.py:51-52,420).  Neither ``mujoco`` nor ``mujoco.mjx`` can be installed in this image (no wheel in
/opt/wheelhouse, no network), so today ``available()`` is False and every consumer falls back to the C++ port.  If a driver
ever provides an install -- ``baseline/_ref/`` (git-ignored; travels to the GPU box) or site-packages -- this module

* times ``mj_step`` threaded over envs on the host cores (``MujocoReference.rate``; MuJoCo releases the GIL inside mj_step, one
  ``MjData`` per env) -> ``bench.py --impl reference`` reports ``cpu_baseline.kind = "reference"`` instead of ``"port"``;
* steps a batch of states through ten ``mj_step`` calls and hands back qpos / qvel / efc_force so that a test can diff the
  oracle (and through it the CUDA kernels) against the reference itself (``tests/test_mujoco_hook.py``) -- the step that turns
  "parity unpinned"  into a pinned comparison.

The scene MJCF is needed as well: it is looked up under ``$ODUCK_REFERENCE_ROOT``, ``/root/reference`` and
``baseline/_ref`` (``playground/open_duck_mini_v2/xmls/<scene>.xml``).
"""
from __future__ import annotations

import importlib
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")
_SCENES = {"flat_terrain": "scene_flat_terrain.xml", "flat_terrain_backlash": "scene_flat_terrain_backlash.xml",
           "rough_terrain_backlash": "scene_rough_terrain_backlash.xml"}


def _import_mujoco():
    """``import mujoco`` from the environment or from baseline/_ref (a ``pip install --target`` tree); None when absent."""
    if "mujoco" in sys.modules:
        return sys.modules["mujoco"]
    added = False
    if os.path.isdir(REF_INSTALL) and REF_INSTALL not in sys.path:
        sys.path.append(REF_INSTALL)
        added = True
    try:
        return importlib.import_module("mujoco")
    except Exception:
        if added:
            sys.path.remove(REF_INSTALL)
        return None


def scene_path(task: str) -> Optional[str]:
    """The reference's MJCF of ``task`` (constants.py:28-34) if a reference checkout is reachable."""
    roots = [os.environ.get("ODUCK_REFERENCE_ROOT"), "/root/reference", REF_INSTALL]
    for r in roots:
        if not r:
            continue
        p = os.path.join(r, "playground", "open_duck_mini_v2", "xmls", _SCENES[task])
        if os.path.exists(p):
            return p
    return None


def available(task: str = "flat_terrain_backlash") -> bool:
    return _import_mujoco() is not None and scene_path(task) is not None


def why_unavailable(task: str = "flat_terrain_backlash") -> str:
    if _import_mujoco() is None:
        return "import mujoco fails (not in the image, no baseline/_ref install)"
    if scene_path(task) is None:
        return "no reference checkout with the scene MJCF (ODUCK_REFERENCE_ROOT, /root/reference, baseline/_ref)"
    return ""


class MujocoReference:
    """N independent ``MjData`` of one ``MjModel``, stepped with ``mujoco.mj_step`` on a thread pool (one env per task)."""

    def __init__(self, task: str = "flat_terrain_backlash", threads: Optional[int] = None, sim_dt: float = 0.002):
        self.mj = _import_mujoco()
        if self.mj is None:
            raise RuntimeError("mujoco is not importable: " + why_unavailable(task))
        path = scene_path(task)
        if path is None:
            raise RuntimeError(why_unavailable(task))
        self.model = self.mj.MjModel.from_xml_path(path)          # mujoco_infer_base.py:19
        self.model.opt.timestep = sim_dt                            # base.py:56
        self.threads = int(threads or os.cpu_count() or 1)
        self.datas = []

    # ---------------------------------------------------------------- state marshalling
    def _ensure(self, n: int) -> None:
        while len(self.datas) < n:
            self.datas.append(self.mj.MjData(self.model))

    def set_state(self, qpos: np.ndarray, qvel: np.ndarray, ctrl: np.ndarray, qacc_warmstart: Optional[np.ndarray] = None) -> None:
        n = qpos.shape[0]
        self._ensure(n)
        for i in range(n):
            d = self.datas[i]
            d.qpos[:] = qpos[i, :self.model.nq]
            d.qvel[:] = qvel[i, :self.model.nv]
            d.ctrl[:] = ctrl[i, :self.model.nu]
            if qacc_warmstart is not None:
                d.qacc_warmstart[:] = qacc_warmstart[i, :self.model.nv]
        self.n = n

    def _step_range(self, lo: int, hi: int, n_substeps: int) -> None:
        mj, m = self.mj, self.model
        for i in range(lo, hi):
            d = self.datas[i]
            for _ in range(n_substeps):
                mj.mj_step(m, d)                                    # mujoco_infer.py:170

    def step(self, n_substeps: int = 10) -> None:
        """One env.step worth of physics for every env: ``n_substeps`` x mj_step, envs split over the thread pool."""
        n, t = self.n, max(1, min(self.threads, self.n))
        bounds = [(k * n // t, (k + 1) * n // t) for k in range(t)]
        if t == 1:
            self._step_range(0, n, n_substeps)
            return
        with ThreadPoolExecutor(max_workers=t) as pool:
            list(pool.map(lambda b: self._step_range(b[0], b[1], n_substeps), bounds))

    def state(self) -> Dict[str, np.ndarray]:
        """qpos, qvel, qacc and the non-zero constraint forces (sorted per env: MuJoCo instantiates only active rows, in its own
        order, so the comparable quantity is the multiset of non-zero ``efc_force`` values)."""
        n = self.n
        out = {"qpos": np.stack([np.array(self.datas[i].qpos) for i in range(n)]), "qvel": np.stack([np.array(self.datas[i].qvel) for i in range(n)]),
               "qacc": np.stack([np.array(self.datas[i].qacc) for i in range(n)])}
        out["efc_force_sorted"] = [np.sort(np.asarray(self.datas[i].efc_force)[np.abs(np.asarray(self.datas[i].efc_force)) > 0]) for i in range(n)]
        return out

    # ---------------------------------------------------------------- timing
    def rate(self, n_envs: int, steps: int, n_substeps: int = 10, seed: int = 1) -> Tuple[float, float]:
        """env-steps/s of ``n_substeps`` x mj_step per env on the host threads: keyframe ``home`` states, ctrl = home +
        0.25 U(-1, 1) redrawn every control step.  Returns (env-steps/s, ms per batched step)."""
        mj, m = self.mj, self.model
        self._ensure(n_envs)
        self.n = n_envs
        rs = np.random.default_rng(seed)
        for i in range(n_envs):
            mj.mj_resetDataKeyframe(m, self.datas[i], 0)
        home = np.array(self.datas[0].ctrl)

        def draw():
            c = home[None] + 0.25 * rs.uniform(-1, 1, (n_envs, m.nu))
            for i in range(n_envs):
                self.datas[i].ctrl[:] = c[i]

        draw()
        self.step(n_substeps)                                       # warm-up
        t0 = time.perf_counter()
        for _ in range(steps):
            draw()
            self.step(n_substeps)
        dt = time.perf_counter() - t0
        return n_envs * steps / dt, dt / steps * 1e3


def diff_against(ref_env, n_substeps: int = 10, task: str = "flat_terrain_backlash") -> Dict[str, float]:
    """Step the states held by ``ref_env`` (a Joystick over the oracle -- or any library -- with domain randomisation OFF) through
    MuJoCo and through the library; returns the max abs differences.  ``ref_env`` advances by ``n_substeps`` substeps."""
    import torch
    f = lambda name: np.asarray(ref_env.buffer(name).cpu().numpy(), dtype=np.float64)      # noqa: E731
    mjr = MujocoReference(task)
    mjr.set_state(f("QPOS"), f("QVEL"), f("CTRL"), f("QACC_WARM"))
    mjr.step(n_substeps)
    ref_env.physics_substeps(None, n_substeps)
    if ref_env.device.type == "cuda":
        torch.cuda.synchronize()
    got = mjr.state()
    nq, nv = mjr.model.nq, mjr.model.nv
    out = {"qpos": float(np.abs(got["qpos"] - f("QPOS")[:, :nq]).max()), "qvel": float(np.abs(got["qvel"] - f("QVEL")[:, :nv]).max()),
           "qacc": float(np.abs(got["qacc"] - f("QACC")[:, :nv]).max())}
    ef = f("EFC_FORCE")
    worst, mismatched = 0.0, 0
    for i in range(mjr.n):
        mine = np.sort(ef[i][np.abs(ef[i]) > 0])
        theirs = got["efc_force_sorted"][i]
        if mine.shape != theirs.shape:
            mismatched += 1
            continue
        if mine.size:
            worst = max(worst, float((np.abs(mine - theirs) / (1e-4 + np.abs(theirs))).max()))
    out["efc_force_rel"], out["active_set_mismatch_envs"] = worst, mismatched
    return out
