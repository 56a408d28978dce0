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


def test_onnx_file_parses_with_the_protobuf_runtime_against_the_onnx_schema(tmp_path):
    """An INDEPENDENT decoder: Google's protobuf runtime parses the exported bytes against the ONNX schema (tests/onnx_schema.py:
    the published message names and field numbers), nothing of export_onnx.py's own reader is used.  The parsed model must be a
    well-formed opset-13 graph (no unknown fields anywhere, every node input defined before use, one input / one output with
    static shapes), carry the checkpoint's tensors bit for bit, and -- evaluated from the PARSED messages -- reproduce the torch
    policy."""
    import onnx_schema
    M = onnx_schema.onnx_messages()
    pol, mean, std, params = _checkpoint(3)
    norm, tree = ex.brax_param_tree(params)
    path = ex.export_onnx((norm, tree), 14, (512, 256, 128), 101, str(tmp_path / "policy.onnx"))
    model = M["ModelProto"]()
    model.ParseFromString(open(path, "rb").read())

    from google.protobuf import unknown_fields

    def no_unknown(msg):                                                  # recursively: every byte was claimed by a declared field
        assert len(unknown_fields.UnknownFieldSet(msg)) == 0, type(msg).__name__
        for fd, v in msg.ListFields():
            if fd.message_type is not None:
                for sub in (v if (fd.is_repeated if hasattr(fd, "is_repeated") else fd.label == fd.LABEL_REPEATED) else [v]):
                    no_unknown(sub)
    no_unknown(model)
    assert model.SerializeToString() != b"" and model.ir_version == 8 and model.producer_name == "open_duck_playground_b200"
    assert len(model.opset_import) == 1 and model.opset_import[0].domain == "" and model.opset_import[0].version == 13
    g = model.graph
    assert g.name == "open_duck_policy" and len(g.input) == 1 and len(g.output) == 1
    dims = lambda vi: [d.dim_value for d in vi.type.tensor_type.shape.dim]  # noqa: E731
    assert g.input[0].name == "obs" and g.input[0].type.tensor_type.elem_type == 1 and dims(g.input[0]) == [1, 101]
    assert g.output[0].name == "continuous_actions" and g.output[0].type.tensor_type.elem_type == 1 and dims(g.output[0]) == [1, 14]
    # initializers: float tensors with raw little-endian data of exactly prod(dims) elements, equal to the checkpoint's
    init = {}
    for t in g.initializer:
        assert t.data_type == 1 and len(t.raw_data) == 4 * int(np.prod(list(t.dims))) and not t.float_data
        init[t.name] = np.frombuffer(t.raw_data, "<f4").reshape(list(t.dims))
    assert np.array_equal(init["mean"], norm.mean["state"]) and np.array_equal(init["std"], norm.std["state"])
    for i in range(4):
        k, b = tree["params"][f"hidden_{i}"]["kernel"], tree["params"][f"hidden_{i}"]["bias"]
        if i == 3:
            k, b = k[:, :14], b[:14]                                     # the loc half of the head
        assert np.array_equal(init[f"hidden_{i}/kernel"], k) and np.array_equal(init[f"hidden_{i}/bias"], b)
    # nodes: standard-domain ops of opset 13, topologically ordered, unique output names
    assert [n.op_type for n in g.node] == ["Sub", "Div"] + ["MatMul", "Add", "Sigmoid", "Mul"] * 3 + ["MatMul", "Add", "Tanh"]
    known, outs = set(init) | {"obs"}, set()
    for n in g.node:
        assert n.domain == "" and len(n.attribute) == 0 and all(i in known for i in n.input) and len(n.output) == 1 and n.output[0] not in outs
        known.add(n.output[0]); outs.add(n.output[0])
    assert g.output[0].name in outs
    # evaluate the PARSED graph
    ops = {"Sub": np.subtract, "Div": np.divide, "MatMul": np.matmul, "Add": np.add, "Mul": np.multiply,
           "Sigmoid": lambda a: 1.0 / (1.0 + np.exp(-a)), "Tanh": np.tanh}
    obs = np.random.default_rng(1).normal(0, 1, (1, 101)).astype(np.float32)
    env = dict(init, obs=obs)
    for n in g.node:
        env[n.output[0]] = ops[n.op_type](*[env[i] for i in n.input]).astype(np.float32)
    with torch.no_grad():
        ref = torch.tanh(pol((torch.from_numpy(obs) - mean) / std)[:, :14]).numpy()
    assert np.abs(env["continuous_actions"] - ref).max() < 2e-6
    assert np.array_equal(env["continuous_actions"], ex.run_onnx(path, obs))      # and the module's own reader agrees with it
