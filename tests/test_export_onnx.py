"""Checkpoint <-> deployment format: the Brax-style parameter tree the reference exporter reads
(common/export_onnx.py:91-92,132-146) and the ONNX policy graph it produces (export_onnx.py:64-72: tanh of the loc half of a
swish MLP on normalised observations), checked against the torch policy and the library's deterministic actor."""
import numpy as np
import pytest
import torch

from open_duck_playground_b200 import export_onnx as ex, ppo, rng as jr
from open_duck_playground_b200.joystick import Joystick


def _checkpoint(seed=0):
    torch.manual_seed(seed)
    pol = ppo.MLP([101, 512, 256, 128, 28])
    with torch.no_grad():
        for lin in pol.layers:
            lin.bias.uniform_(-0.1, 0.1)
    mean, std = torch.randn(101) * 0.3, torch.rand(101) + 0.5
    params = {"normalizer": {"state": {"mean": mean, "std": std, "count": 10.0}}, "policy": {k: v for k, v in pol.state_dict().items()}}
    return pol, mean, std, params


def test_brax_param_tree_layout():
    pol, mean, std, params = _checkpoint()
    norm, tree = ex.brax_param_tree(params)
    assert set(tree["params"]) == {"hidden_0", "hidden_1", "hidden_2", "hidden_3"}
    assert tree["params"]["hidden_0"]["kernel"].shape == (101, 512) and tree["params"]["hidden_3"]["kernel"].shape == (128, 28)   # flax [in][out]
    assert np.array_equal(tree["params"]["hidden_1"]["kernel"], pol.layers[1].weight.detach().numpy().T)
    assert np.array_equal(norm.mean["state"], mean.numpy()) and np.array_equal(norm.std["state"], std.numpy())


def test_onnx_file_reproduces_the_deterministic_policy(tmp_path):
    pol, mean, std, params = _checkpoint(1)
    path = ex.export_onnx(ex.brax_param_tree(params), 14, (512, 256, 128), 101, str(tmp_path / "policy.onnx"))
    raw = open(path, "rb").read()
    assert raw[:2] == b"\x08\x08" and b"continuous_actions" in raw and b"hidden_3/kernel" in raw      # ir_version = 8 first, named like the reference graph
    obs = np.random.default_rng(0).normal(0, 1, (5, 1, 101)).astype(np.float32)
    with torch.no_grad():
        ref = torch.tanh(pol((torch.from_numpy(obs[:, 0]) - mean) / std)[:, :14]).numpy()
    out = np.concatenate([ex.run_onnx(path, o) for o in obs])
    assert out.shape == (5, 14) and np.abs(out - ref).max() < 2e-6


def test_onnx_matches_library_actor(oracle, tmp_path):
    """mujoco_infer.py feeds the ONNX policy what the env hands out as obs['state']: same action as oduck_policy_forward."""
    pol, mean, std, params = _checkpoint(2)
    env = Joystick("flat_terrain_backlash", library=oracle)
    st = env.reset(jr.split(jr.PRNGKey(0), 4))
    w = ppo.PolicyWeights(pol, 101, env.device)
    w.refresh(mean, std)
    act, _, _ = ppo.policy_forward(env, w, None, deterministic=True)
    path = ex.export_onnx(ex.brax_param_tree(params), 14, (512, 256, 128), 101, str(tmp_path / "p.onnx"))
    out = np.concatenate([ex.run_onnx(path, st.obs["state"][i:i + 1].float().numpy()) for i in range(4)])
    assert np.abs(out - act.numpy()).max() < 1e-5


def test_headless_inference_loop(oracle, tmp_path):
    """Checkpoint -> ONNX -> sim-to-sim loop (the role of mujoco_infer.py), on the oracle env: the exported policy drives the env
    exactly like the library's deterministic actor."""
    from open_duck_playground_b200 import infer
    pol, mean, std, params = _checkpoint(3)
    with torch.no_grad():
        pol.layers[-1].weight.mul_(0.05)                                  # small actions: the duck keeps standing for the test's 12 steps
        params["policy"] = {k: v for k, v in pol.state_dict().items()}
    path = ex.export_onnx(ex.brax_param_tree(params), 14, (512, 256, 128), 101, str(tmp_path / "p.onnx"))
    env = Joystick("flat_terrain_backlash", library=oracle)
    out = infer.run_policy(env, path, steps=12, seed=5, command=[0.1, 0, 0, 0, 0, 0, 0], num_envs=2)
    assert out["steps"] == 12 and np.isfinite(out["mean_step_reward"]) and 0.05 < out["final_base_height"] < 0.3
    assert torch.allclose(env.buffer("INFO_COMMAND")[:, 0].double(), torch.full((2,), 0.1, dtype=torch.float64), atol=1e-6)
    # same rollout with the library actor (deterministic) from the same reset: identical trajectory
    env2 = Joystick("flat_terrain_backlash", library=oracle)
    st = env2.reset(jr.split(jr.PRNGKey(5), 2))
    w = ppo.PolicyWeights(pol, 101, env2.device)
    w.refresh(mean, std)
    for _ in range(12):
        st.info["command"][:] = torch.tensor([0.1, 0, 0, 0, 0, 0, 0], dtype=st.info["command"].dtype)
        act, _, _ = ppo.policy_forward(env2, w, None, deterministic=True)
        st = env2.step(st, act)
    assert torch.allclose(env.buffer("QPOS").double(), env2.buffer("QPOS").double(), atol=1e-4)
