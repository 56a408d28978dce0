"""Checkpoint <-> deployment format (reference playground/common/export_onnx.py:6-189).

The reference turns the Brax parameter tuple ``(normalizer_params, policy_params)`` into a Keras MLP and then into an ONNX file
that ``mujoco_infer.py`` / the robot runtime load with onnxruntime: input ``obs [1, obs_size]``, graph
``tanh(split(MLP((obs - mean) / std))[0])`` with swish hidden layers (export_onnx.py:64-72,97-102).  Neither tensorflow, tf2onnx
nor onnx exist in this image, and the graph is tiny, so this module

  * ``brax_param_tree(params)``  -- re-expresses a trainer checkpoint in the tree the reference exporter reads
    (``params[0].mean / .std`` per obs key, ``params[1]["params"]["hidden_i"]["kernel" | "bias"]``, kernels ``[in][out]``,
    export_onnx.py:91-92,132-146), so reference tooling keeps working on our checkpoints;
  * ``export_onnx(params, act_size, hidden_sizes, obs_size, output_path)`` -- writes the same graph as a standard ONNX file
    (opset 13: Sub, Div, MatMul, Add, Sigmoid, Mul, Tanh) with a few lines of protobuf wire encoding; only the ``loc`` half of
    the head is exported (the reference graph computes and drops the other half);
  * ``run_onnx(path, obs)`` -- decodes such a file and evaluates it with NumPy (tests, and CPU inference without onnxruntime).
"""
from __future__ import annotations

import struct
from types import SimpleNamespace
from typing import Dict, Sequence, Tuple

import numpy as np

# ------------------------------------------------------------------------------------------------- protobuf wire format
_VARINT, _LEN = 0, 2


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _f_int(field: int, v: int) -> bytes:
    return _varint((field << 3) | _VARINT) + _varint(v)


def _f_bytes(field: int, v: bytes) -> bytes:
    return _varint((field << 3) | _LEN) + _varint(len(v)) + v


def _f_str(field: int, v: str) -> bytes:
    return _f_bytes(field, v.encode())


def _decode(buf: bytes) -> Dict[int, list]:
    """Generic message -> {field: [values]} (varints as int, length-delimited as bytes)."""
    out: Dict[int, list] = {}
    i = 0
    while i < len(buf):
        key = 0; shift = 0
        while True:
            b = buf[i]; i += 1
            key |= (b & 0x7F) << shift; shift += 7
            if not b & 0x80:
                break
        field, wt = key >> 3, key & 7
        if wt == _VARINT:
            v = 0; shift = 0
            while True:
                b = buf[i]; i += 1
                v |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
        elif wt == _LEN:
            n = 0; shift = 0
            while True:
                b = buf[i]; i += 1
                n |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            v = bytes(buf[i:i + n]); i += n
        elif wt == 5:
            v = struct.unpack("<I", buf[i:i + 4])[0]; i += 4
        elif wt == 1:
            v = struct.unpack("<Q", buf[i:i + 8])[0]; i += 8
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


# ------------------------------------------------------------------------------------------------- ONNX messages (onnx.proto3)
_FLOAT = 1


def _tensor(name: str, a: np.ndarray) -> bytes:                      # TensorProto: dims=1, data_type=2, name=8, raw_data=9
    a = np.ascontiguousarray(a, dtype=np.float32)
    return b"".join(_f_int(1, d) for d in a.shape) + _f_int(2, _FLOAT) + _f_str(8, name) + _f_bytes(9, a.tobytes())


def _node(op: str, inputs: Sequence[str], outputs: Sequence[str], name: str) -> bytes:   # NodeProto: input=1 output=2 name=3 op_type=4
    return b"".join(_f_str(1, i) for i in inputs) + b"".join(_f_str(2, o) for o in outputs) + _f_str(3, name) + _f_str(4, op)


def _value_info(name: str, shape: Sequence[int]) -> bytes:            # ValueInfoProto{name=1, type=2{tensor_type=1{elem_type=1, shape=2{dim=1{dim_value=1}}}}}
    dims = b"".join(_f_bytes(1, _f_int(1, d)) for d in shape)
    ttype = _f_int(1, _FLOAT) + _f_bytes(2, dims)
    return _f_str(1, name) + _f_bytes(2, _f_bytes(1, ttype))


def brax_param_tree(params: dict, obs_key: str = "state") -> Tuple[SimpleNamespace, dict]:
    """Trainer checkpoint (``PPOTrainer.params()``) -> ``(normalizer_params, policy_params)`` as Brax lays them out."""
    norm = params["normalizer"]
    n = SimpleNamespace(mean={k: np.asarray(v["mean"], np.float32) for k, v in norm.items()},
                        std={k: np.asarray(v["std"], np.float32) for k, v in norm.items()},
                        count={k: v["count"] for k, v in norm.items()})
    pol = params["policy"]
    nl = len([k for k in pol if k.endswith(".weight")])
    tree = {"params": {f"hidden_{i}": {"kernel": np.asarray(pol[f"layers.{i}.weight"], np.float32).T.copy(),       # flax: [in][out]
                                       "bias": np.asarray(pol[f"layers.{i}.bias"], np.float32)} for i in range(nl)}}
    return n, tree


def export_onnx(params, act_size: int, hidden_layer_sizes: Sequence[int], obs_size: int, output_path: str = "ONNX.onnx", obs_key: str = "state") -> str:
    """Same call shape as the reference's ``export_onnx(params, act_size, ppo_params, obs_size, output_path)``; ``params`` is the
    Brax-style tuple (``brax_param_tree``) and ``hidden_layer_sizes`` stands for ``ppo_params.network_factory.policy_hidden_layer_sizes``."""
    norm, pol = params[0], params[1]
    mean, std = np.asarray(norm.mean[obs_key], np.float32), np.asarray(norm.std[obs_key], np.float32)
    layers = pol["params"]
    nl = len(hidden_layer_sizes) + 1
    assert mean.shape == (obs_size,) and len(layers) == nl
    inits, nodes = [_tensor("mean", mean), _tensor("std", std)], []
    nodes.append(_node("Sub", ["obs", "mean"], ["centered"], "center"))
    nodes.append(_node("Div", ["centered", "std"], ["x0"], "normalize"))
    x = "x0"
    for i in range(nl):
        k, b = np.asarray(layers[f"hidden_{i}"]["kernel"], np.float32), np.asarray(layers[f"hidden_{i}"]["bias"], np.float32)
        if i == nl - 1:                                               # loc, _ = split(logits, 2): only the loc half reaches the output
            assert k.shape[1] == 2 * act_size
            k, b = k[:, :act_size], b[:act_size]
        else:
            assert k.shape[1] == hidden_layer_sizes[i]
        inits += [_tensor(f"hidden_{i}/kernel", k), _tensor(f"hidden_{i}/bias", b)]
        nodes.append(_node("MatMul", [x, f"hidden_{i}/kernel"], [f"mm{i}"], f"hidden_{i}/MatMul"))
        nodes.append(_node("Add", [f"mm{i}", f"hidden_{i}/bias"], [f"z{i}"], f"hidden_{i}/BiasAdd"))
        if i < nl - 1:                                                # swish = z * sigmoid(z) (export_onnx.py:101)
            nodes.append(_node("Sigmoid", [f"z{i}"], [f"s{i}"], f"hidden_{i}/Sigmoid"))
            nodes.append(_node("Mul", [f"z{i}", f"s{i}"], [f"h{i}"], f"hidden_{i}/Swish"))
            x = f"h{i}"
    nodes.append(_node("Tanh", [f"z{nl - 1}"], ["continuous_actions"], "tanh"))
    graph = (b"".join(_f_bytes(1, n) for n in nodes) + _f_str(2, "open_duck_policy") + b"".join(_f_bytes(5, t) for t in inits) +
             _f_bytes(11, _value_info("obs", [1, obs_size])) + _f_bytes(12, _value_info("continuous_actions", [1, act_size])))
    model = (_f_int(1, 8) + _f_str(2, "open_duck_playground_b200") + _f_bytes(7, graph) +
             _f_bytes(8, _f_str(1, "") + _f_int(2, 13)))                # ir_version 8, opset 13
    with open(output_path, "wb") as f:
        f.write(model)
    return output_path


def run_onnx(path: str, obs: np.ndarray) -> np.ndarray:
    """Evaluate a file written by ``export_onnx`` with NumPy (the op subset above)."""
    model = _decode(open(path, "rb").read())
    graph = _decode(model[7][0])
    env: Dict[str, np.ndarray] = {}
    for t in graph.get(5, []):
        d = _decode(t)
        assert d[2][0] == _FLOAT
        env[d[8][0].decode()] = np.frombuffer(d[9][0], np.float32).reshape([int(x) for x in d.get(1, [])])
    inp = _decode(graph[11][0])[1][0].decode()
    out = _decode(graph[12][0])[1][0].decode()
    env[inp] = np.asarray(obs, np.float32)
    ops = {"Sub": lambda a, b: a - b, "Div": lambda a, b: a / b, "MatMul": lambda a, b: a @ b, "Add": lambda a, b: a + b,
           "Mul": lambda a, b: a * b, "Sigmoid": lambda a: 1.0 / (1.0 + np.exp(-a)), "Tanh": np.tanh}
    for nb in graph[1]:
        n = _decode(nb)
        args = [env[i.decode()] for i in n[1]]
        env[n[2][0].decode()] = ops[n[4][0].decode()](*args).astype(np.float32)
    return env[out]
