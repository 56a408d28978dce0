"""oracle/mujoco_ref.py -- the hook that times and diffs against the REAL reference physics (``mujoco.mj_step``,
/root/reference/playground/open_duck_mini_v2/mujoco_infer.py:170) whenever a MuJoCo install is reachable.  The image has none,
so the hook is exercised with a FAKE ``mujoco`` module whose ``mj_step`` is the oracle's own substep: that checks the plumbing
(install discovery, state marshalling, threading over envs, the diff report, bench.py's choice of ``kind``), not the physics.
With a real install ``test_states_match_mujoco`` runs instead of being skipped and pins the oracle to the reference."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from open_duck_playground_b200 import rng as jr
from open_duck_playground_b200.joystick import Joystick
from oracle import mujoco_ref

TASK = "flat_terrain_backlash"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_mujoco(oracle):
    """A module with the five names the hook uses.  One MjData = one 1-env oracle handle (DR off = the nominal model)."""
    mod = types.ModuleType("mujoco")

    class MjModel:
        def __init__(self):
            self._env = Joystick(TASK, library=oracle)
            m = self._env.mj_model
            self.nq, self.nv, self.nu = m.nq, m.nv, m.nu
            self.opt = types.SimpleNamespace(timestep=0.002)

        @staticmethod
        def from_xml_path(path):
            assert os.path.exists(path)
            return MjModel()

    class MjData:
        def __init__(self, model):
            self._env = model._env.spawn()
            self._env.reset(jr.split(jr.PRNGKey(0), 1))
            self._env.handle.set_rollout_sink(None)
            m = model
            self.qpos, self.qvel, self.ctrl = np.zeros(m.nq), np.zeros(m.nv), np.zeros(m.nu)
            self.qacc_warmstart, self.qacc, self.efc_force = np.zeros(m.nv), np.zeros(m.nv), np.zeros(0)

    def mj_resetDataKeyframe(model, data, key):
        mm = model._env.mj_model
        data.qpos[:] = mm.key_qpos[:model.nq]; data.qvel[:] = 0; data.ctrl[:] = mm.key_ctrl[:model.nu]; data.qacc_warmstart[:] = 0

    def mj_step(model, data):
        e = data._env
        f = lambda a: torch.from_numpy(np.asarray(a, np.float32)[None].copy())      # noqa: E731
        e.set_state(f(data.qpos), f(data.qvel), f(data.qacc_warmstart))
        d = e.physics_substeps(f(data.ctrl), 1)
        data.qpos[:] = d.qpos[0].numpy(); data.qvel[:] = d.qvel[0].numpy(); data.qacc[:] = d.qacc[0].numpy()
        data.qacc_warmstart[:] = d.qacc_warmstart[0].numpy()
        ef = d.efc_force[0].numpy()
        data.efc_force = ef[np.abs(ef) > 0].copy()

    mod.MjModel, mod.MjData, mod.mj_step, mod.mj_resetDataKeyframe = MjModel, MjData, mj_step, mj_resetDataKeyframe
    return mod


@pytest.fixture()
def fake_install(oracle, tmp_path, monkeypatch):
    xml = tmp_path / "playground" / "open_duck_mini_v2" / "xmls"
    xml.mkdir(parents=True)
    (xml / "scene_flat_terrain_backlash.xml").write_text("<mujoco/>")
    monkeypatch.setenv("ODUCK_REFERENCE_ROOT", str(tmp_path))
    monkeypatch.setitem(sys.modules, "mujoco", _fake_mujoco(oracle))
    return tmp_path


def test_hook_reports_unavailable_without_an_install(monkeypatch):
    monkeypatch.delitem(sys.modules, "mujoco", raising=False)
    try:
        import mujoco  # noqa: F401
        pytest.skip("a real MuJoCo install is present")
    except ImportError:
        pass
    assert not mujoco_ref.available(TASK) and "import mujoco fails" in mujoco_ref.why_unavailable(TASK)
    with pytest.raises(RuntimeError, match="not importable"):
        mujoco_ref.MujocoReference(TASK)


def test_hook_times_and_diffs_through_a_fake_module(oracle, fake_install):
    assert mujoco_ref.available(TASK) and mujoco_ref.scene_path(TASK).startswith(str(fake_install))
    ref = mujoco_ref.MujocoReference(TASK, threads=3)
    rate, ms = ref.rate(n_envs=6, steps=2, n_substeps=2)
    assert rate > 0 and ms > 0 and len(ref.datas) == 6
    q = np.stack([np.array(d.qpos) for d in ref.datas])
    assert np.isfinite(q).all() and np.abs(q[0] - q[1]).max() > 0          # every env got its own ctrl draw and was stepped
    # the diff: the same states through the hook (fake mj_step = oracle substep, f32 state hand-over) and through the library
    env = Joystick(TASK, library=oracle)
    env.reset(jr.split(jr.PRNGKey(4), 5))                                   # domain randomisation off: the nominal model MuJoCo loads
    d = mujoco_ref.diff_against(env, n_substeps=3, task=TASK)
    assert d["active_set_mismatch_envs"] == 0 and d["qpos"] < 1e-5 and d["qvel"] < 1e-3 and d["efc_force_rel"] < 1e-2, d


def test_bench_reference_arm_prefers_the_real_reference(oracle, fake_install, monkeypatch):
    """bench.py --impl reference / cpu_baseline: kind 'reference' (mj_step) when the hook is available, else the C++ port."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "TASK", TASK)
    r = bench.cpu_reference(n_envs=4, steps=1)
    assert r["kind"] == "reference" and r["value"] > 0 and "mujoco.mj_step" in r["sample"]
    monkeypatch.delenv("ODUCK_REFERENCE_ROOT")
    monkeypatch.setattr(mujoco_ref, "scene_path", lambda task: None)
    r = bench.cpu_reference(n_envs=4, steps=1)
    assert r["kind"] == "port" and r["value"] > 0 and r["cores"] >= 1


@pytest.mark.skipif(not mujoco_ref.available(TASK) or "mujoco" not in sys.modules or not hasattr(sys.modules.get("mujoco"), "__file__"),
                    reason="no MuJoCo install (image has none: see oracle/mujoco_ref.py); runs when baseline/_ref provides one")
def test_states_match_mujoco(oracle):
    """Pins the oracle to the reference's own CPU physics: 10 substeps from shared states (SURVEY.md 8c tolerances)."""
    env = Joystick(TASK, library=oracle)
    env.reset(jr.split(jr.PRNGKey(0), 64))
    d = mujoco_ref.diff_against(env, n_substeps=10, task=TASK)
    print(d)
    assert d["qpos"] < 1e-4 and d["qvel"] < 1e-3 and d["efc_force_rel"] < 1e-2 and d["active_set_mismatch_envs"] <= 1
