"""The reference CLI (open_duck_mini_v2/runner.py:35-64, common/runner.py:24-118) end to end on the CPU: ``runner.main()`` with the
oracle library standing in for liboduck_cuda.so (the product itself has no CPU path: the stand-in is injected here, in the test).
Covers what no other test touches: argument names, env selection, checkpoint + ONNX files per evaluation, ``--restore_checkpoint_path``."""
import functools
import glob
import os
import sys

import numpy as np
import pytest
import torch

from open_duck_playground_b200 import capi, ppo, runner
from open_duck_playground_b200.export_onnx import run_onnx


@pytest.fixture()
def cpu_runner(oracle, monkeypatch):
    monkeypatch.setattr(capi, "load_cuda_library", lambda: oracle)
    # (unroll 8, not less: a shorter first batch leaves the oldest action-history features at the normaliser's floor std)
    small = functools.partial(ppo.PPOConfig, unroll_length=8, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=4, episode_length=12, num_evals=2)
    monkeypatch.setattr(runner, "PPOConfig", small)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return runner


@pytest.mark.parametrize("env_name", ["joystick", "standing"])
def test_cli_trains_checkpoints_and_resumes(cpu_runner, tmp_path, monkeypatch, capsys, env_name):
    out = tmp_path / "ckpt"
    argv = ["runner", "--output_dir", str(out), "--num_timesteps", str(8 * 8 * 3), "--env", env_name, "--task", "flat_terrain_backlash", "--num_envs", "8"]
    monkeypatch.setattr(sys, "argv", argv)
    cpu_runner.main()
    text = capsys.readouterr().out
    assert "Observation size: " + ("101" if env_name == "joystick" else "85") in text and "STEP: 0 reward:" in text and "STEP: 192 reward:" in text and "Saving checkpoint (step: 192)" in text
    pts, onnxs = sorted(glob.glob(str(out / "*.pt"))), sorted(glob.glob(str(out / "*.onnx")))
    assert len(pts) >= 1 and len(onnxs) == len(pts)                          # one ONNX next to every checkpoint (common/runner.py:76-84)
    ck = torch.load(pts[-1], weights_only=False)
    assert {"normalizer", "policy", "value", "optimizer", "env_steps"} <= set(ck)
    assert ck["env_steps"] == 192
    obs_dim = 101 if env_name == "joystick" else 85
    act = run_onnx(onnxs[-1], np.zeros((1, obs_dim), np.float32))
    assert act.shape == (1, 14) and np.all(np.abs(act) <= 1.0)              # deterministic tanh(loc) policy
    # resume: the restored run continues from the checkpoint's step count and weights
    monkeypatch.setattr(sys, "argv", argv + ["--restore_checkpoint_path", pts[-1]])
    seen = {}
    orig_load = ppo.PPOTrainer.load

    def spy(self, params):
        orig_load(self, params)
        seen["env_steps"] = self.env_steps
        seen["same"] = all(torch.equal(a, b) for a, b in zip(self.policy.state_dict().values(), params["policy"].values()))
    monkeypatch.setattr(ppo.PPOTrainer, "load", spy)
    cpu_runner.main()
    assert seen == {"env_steps": 192, "same": True}
    text = capsys.readouterr().out
    assert "STEP: 192 reward:" in text and "STEP: 384 reward:" in text           # evaluation of the restored policy, then 3 more training steps of 8 x 8


def test_cli_rejects_an_unknown_env(cpu_runner, tmp_path, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["runner", "--output_dir", str(tmp_path), "--env", "hopping"])
    with pytest.raises(ValueError, match="Unknown env hopping"):
        cpu_runner.main()


def test_cli_keeps_the_reference_flags():
    """open_duck_mini_v2/runner.py:36-56: --output_dir --num_timesteps --env --task --restore_checkpoint_path, same defaults."""
    import inspect
    src = inspect.getsource(runner.main)
    for flag, default in (("--output_dir", '"checkpoints"'), ("--num_timesteps", "150000000"), ("--env", '"joystick"'), ("--task", '"flat_terrain"'),
                          ("--restore_checkpoint_path", "None")):
        line = next(ln for ln in src.splitlines() if f'"{flag}"' in ln)
        assert f"default={default}" in line, line
