"""PPO rollout + update around the fused env step (reference: Brax ``ppo.train`` as driven by common/runner.py:86-118).

The hot path -- policy forward (``oduck_policy_forward``) and ``env.step``, fused as ``oduck_rollout_step`` -- runs in the CUDA
library, and so does the update (``DeviceLearner`` over include/oduck_ppo.h: GAE, clipped surrogate, backward, clip + Adam);
the PyTorch fp32 twin of the update in this file is its checker and the CPU path of the tests.  Hyper-parameters default to ``locomotion_params.brax_ppo_config("BerkeleyHumanoidJoystickFlatTerrain")`` -- the table
the reference looks up at common/runner.py:87-89 (values: SURVEY.md 3.1).

Multi-GPU: ranks own disjoint env shards; ``update_mode = "replicated"``: one ``all_gather`` of the rollout per
training step over NCCL, then the same update on every rank; ``"sharded"`` (the CUDA default): every rank updates on its own
shard and the flat gradient is all-reduced per minibatch inside the captured update graph (Brax's pmean).
"""
from __future__ import annotations

import copy
import math
import os
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import capi, rng as jr


@dataclass
class PPOConfig:
    num_timesteps: int = 150_000_000
    num_envs: int = 8192
    unroll_length: int = 20
    batch_size: int = 256
    num_minibatches: int = 32
    num_updates_per_batch: int = 4
    discounting: float = 0.97
    gae_lambda: float = 0.95
    learning_rate: float = 3e-4
    entropy_cost: float = 0.005
    clipping_epsilon: float = 0.2
    max_grad_norm: float = 1.0
    reward_scaling: float = 1.0
    normalize_observations: bool = True
    episode_length: int = 1000
    num_evals: int = 15
    num_resets_per_eval: int = 1    # Brax ppo.train: training epochs per evaluation, each followed by a hard reset of every training env with
                                    # fresh keys (the AutoReset wrapper only ever returns an env to the first state of its LAST hard reset);
                                    # 0 = one epoch per evaluation and no hard resets (Brax's default; the reference's table says 1)
    num_eval_envs: int = 128        # Brax ppo.train default: evaluation episodes run on their own envs (rank 0); 0 = estimate from the rollout
    deterministic_eval: bool = False   # Brax default: the evaluator samples actions like the behaviour policy
    policy_hidden_layer_sizes: Tuple[int, ...] = (512, 256, 128)
    value_hidden_layer_sizes: Tuple[int, ...] = (512, 256, 128)
    policy_obs_key: str = "state"
    value_obs_key: str = "privileged_state"
    seed: int = 0
    update_mode: str = "auto"       # "replicated": all_gather rollouts, identical update on every rank;
                                    # "sharded": every rank updates on its own env shard, gradients all-reduced per minibatch (Brax's pmean);
                                    # "auto": sharded on CUDA with world > 1, else replicated
    cuda_graph: bool = True         # (torch learner) capture one minibatch update (fwd + bwd + clip + Adam) in a CUDA graph on GPU
    rollout_pipeline: int = 0       # the rank's envs as P sub-batches with their own handle, CUDA graphs and stream, so that sub-batch 0
                                    # starts unroll step t + 1 under the partial last wave of sub-batch P - 1's step t (8192 envs = 3.46
                                    # waves of k_step).  Same transitions as P = 1, bit for bit (envs are independent, keys sliced).
                                    # 0 = auto: 2 on CUDA from 2048 envs per rank (measured on B200: rollout 26.9 -> 25.8 ms at 8192), else 1
    kernel_rollout_writes: bool = True   # the step / actor kernels store each Transition in the rollout buffers themselves
                                    # (oduck_rollout_step); False = the 7 torch copies per step of round 1 (kept as the checker)
    learner_fused_tail: bool = True  # eager device-learner steps end with the fused cooperative reduce + clip + Adam launch; False: the two plain
                                     # kernels a captured graph uses (same gradients; the global norm is summed in another grouping, so the
                                     # clip factor -- and with it every weight -- may differ in its last bit)
    learner_matmul: str = "fp32"    # device learner GEMMs: "fp32" = fp32-faithful (3 tf32 tensor-core passes; the parity-tested default),
                                    # "tf32" = one tf32 pass, XLA's default arithmetic for f32 dots on NVIDIA GPUs (the reference's own)
    learner: str = "auto"           # "device": the fused learner step of include/oduck_ppo.h (tcgen05 GEMMs, fused GAE/loss/Adam kernels);
                                    # "torch": the PyTorch fp32 twin (the checker of the device learner, and the CPU path of the tests);
                                    # "auto": device on CUDA, torch otherwise


class MLP(torch.nn.Module):
    """flax ``MLP``: Dense layers with swish, lecun_uniform kernels, zero biases, no final activation."""

    def __init__(self, sizes):
        super().__init__()
        self.layers = torch.nn.ModuleList(torch.nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))
        for lin in self.layers:
            bound = math.sqrt(3.0 / lin.in_features)
            torch.nn.init.uniform_(lin.weight, -bound, bound)
            torch.nn.init.zeros_(lin.bias)

    def forward(self, x):
        for i, lin in enumerate(self.layers):
            x = lin(x)
            if i + 1 < len(self.layers):
                x = torch.nn.functional.silu(x)
        return x


class RunningStats:
    """brax.training.acme.running_statistics.update in its own arithmetic (fp32, batch-Welford): mean/std per feature, std
    clipped to [1e-6, 1e6]; ``reduce`` merges the ranks' shards (Brax: psum over the pmap axis)."""

    def __init__(self, dim, device):
        self.count = torch.zeros((), device=device)
        self.mean = torch.zeros(dim, device=device)
        self.m2 = torch.zeros(dim, device=device)           # summed_variance
        self.std = torch.ones(dim, device=device)

    def update(self, batch: torch.Tensor, reduce: bool = False) -> None:
        """Brax: count += n; mean += sum(x - mean_old) / count; summed_variance += sum((x - mean_old) (x - mean_new)).  Both sums come
        out of ONE fused pass over the batch (``torch.var_mean``: the batch is 17 - 35 M floats per update, and the literal form reads
        and writes it six times): with S1 = sum(x - mean_old) = n (m_b - mean_old) and S2 = sum((x - mean_old)^2) = n (v_b + (m_b -
        mean_old)^2), sum((x - mean_old) (x - mean_new)) = S2 - (mean_new - mean_old) S1."""
        b = batch.reshape(-1, batch.shape[-1])
        n = torch.full((), float(b.shape[0]), device=b.device)
        v_b, m_b = torch.var_mean(b, dim=0, correction=0)
        dm = m_b - self.mean
        s1, s2 = n * dm, n * (v_b + dm * dm)
        if reduce:
            pack = torch.cat([n[None], s1, s2]); dist.all_reduce(pack)
            d = s1.shape[0]
            n, s1, s2 = pack[0], pack[1:1 + d], pack[1 + d:]
        self.count = self.count + n
        delta = s1 / self.count
        self.mean = self.mean + delta
        self.m2 = self.m2 + (s2 - delta * s1)
        self.std.copy_(torch.sqrt(self.m2.clamp_min(0.0) / self.count).clamp(1e-6, 1e6))   # rounding can leave a tiny negative sum

    @property
    def mean32(self):
        return self.mean


class PolicyWeights:
    """Packs a policy MLP + normaliser into ``OduckPolicyWeights`` (row-major [in][out] kernels like flax)."""

    def __init__(self, policy: MLP, obs_dim: int, device, external=None, packed=None):
        self.policy, self.device = policy, device
        self.external = external          # [(W [in][out], b)] x 4 views of the device learner's master weights (updated in place)
        self.packed = packed              # [ptr] x 4: the learner's tensor-core operand blocks (always current: no repack, no copy)
        self.struct = capi.OduckPolicyWeights()
        self.struct.obs_dim = obs_dim
        hs = [l.out_features for l in policy.layers]
        for i in range(3):
            self.struct.hidden[i] = hs[i]
        self.struct.out_dim = hs[3]
        self.version = 0                  # bumped by every refresh(); policy_forward() invalidates a handle's repack cache when it changes
        self.w = self.b = None
        self.refresh(torch.zeros(obs_dim, device=device), torch.ones(obs_dim, device=device))

    @torch.no_grad()
    def refresh(self, mean: torch.Tensor, std: torch.Tensor) -> None:
        self.mean, self.std = mean.float().contiguous(), std.float().contiguous()
        self.version += 1
        if self.external is not None:
            self.w, self.b = [w for w, _ in self.external], [b for _, b in self.external]
        elif self.w is None:
            self.w = [l.weight.detach().t().contiguous() for l in self.policy.layers]     # [in][out]
            self.b = [l.bias.detach().clone().contiguous() for l in self.policy.layers]
        else:
            # persistent buffers rewritten in place: the addresses in the struct (and in any captured CUDA graph) stay valid, and
            # the library's repack cache -- keyed on the w[0] address -- is invalidated through the version tag, never by the
            # accident of the caching allocator handing out a new address
            for dst, l in zip(self.w, self.policy.layers):
                dst.copy_(l.weight.detach().t())
            for dst, l in zip(self.b, self.policy.layers):
                dst.copy_(l.bias.detach())
        self.struct.obs_mean, self.struct.obs_std = self.mean.data_ptr(), self.std.data_ptr()
        for i in range(4):
            self.struct.w[i], self.struct.b[i] = self.w[i].data_ptr(), self.b[i].data_ptr()
            self.struct.packed[i] = self.packed[i] if self.packed else None


def policy_forward(env, weights: PolicyWeights, keys: Optional[torch.Tensor], deterministic: bool, obs: Optional[torch.Tensor] = None):
    """A15 through the C-ABI.  Returns (action, raw_action, log_prob) tensors on the env's device."""
    n, na = env.handle.n, env.action_size
    dev = env.device
    act = torch.empty(n, na, device=dev)
    raw = torch.empty(n, na, device=dev)
    logp = torch.empty(n, device=dev)
    kp = 0 if keys is None else env._ptr(keys, torch.int32, (n, 2))
    op = 0 if obs is None else env._ptr(obs, torch.float32, (n, weights.struct.obs_dim))
    if obs is None and weights.struct.obs_dim > env.observation_size["state"][0]:
        raise ValueError(f"policy reads {weights.struct.obs_dim} features but the handle's own obs['state'] records hold "
                         f"{env.observation_size['state'][0]}: pass obs= explicitly (e.g. state.obs['privileged_state'])")
    tag = (id(weights), weights.version)
    if getattr(env, "_policy_tag", None) != tag:             # weights behind the same addresses changed (or another weight set): drop
        env.handle.policy_invalidate()                       # this handle's cached tensor-core repack (rarely used handles included)
        env._policy_tag = tag
    env.handle.policy_forward(weights.struct, op, kp, deterministic, act.data_ptr(), raw.data_ptr(), logp.data_ptr(), env._stream())
    return act, raw, logp


def attach_rollout_sink(env, buf: Optional[Dict[str, torch.Tensor]], env_offset: int = 0) -> None:
    """Point the env's kernels at rollout buffers ``buf`` ([T(+1), n, ...], as made by ``PPOTrainer._new_buffers``): from now on
    ``rollout_step`` stores every Transition there itself (include/oduck.h OduckRolloutSink).  ``None`` detaches."""
    if buf is None:
        env.handle.set_rollout_sink(None)
        env._sink_keep = None
        return
    T, n = buf["reward"].shape
    sk = capi.OduckRolloutSink()
    sk.unroll, sk.num_envs, sk.env_offset = int(T), int(n), int(env_offset)
    sk.policy_dim, sk.value_dim = int(buf["obs_p"].shape[-1]), int(buf["obs_v"].shape[-1])
    for attr, key in (("obs_policy", "obs_p"), ("obs_value", "obs_v"), ("raw_action", "raw"), ("log_prob", "logp"), ("reward", "reward"),
                      ("done", "done"), ("truncation", "trunc")):
        t = buf[key]
        if not t.is_contiguous() or t.dtype != torch.float32 or t.device != env.device:
            raise ValueError(f"rollout buffer {key} must be a contiguous float32 tensor on {env.device}")
        setattr(sk, attr, t.data_ptr())
    env.handle.set_rollout_sink(sk)
    env._sink_keep = (sk, buf)               # the library holds raw pointers: keep the tensors alive with the env


def rollout_step(env, weights: PolicyWeights, keys: torch.Tensor, t: int):
    """A17 through the C-ABI (``oduck_rollout_step``): sample an action from the env's own obs["state"], step, and let the kernels
    write the Transition into the attached sink -- raw action / log-prob by the actor's head, reward / done / truncation and the
    next observations by the step kernel.  No per-step copies."""
    n = env.handle.n
    kp = env._ptr(keys, torch.int32, (n, 2))
    tag = (id(weights), weights.version)
    if getattr(env, "_policy_tag", None) != tag:
        env.handle.policy_invalidate()
        env._policy_tag = tag
    env.handle.rollout_step(weights.struct, kp, t, env._stream())
    return env._state()


def torch_policy_logprob(policy: MLP, obs_n: torch.Tensor, raw: torch.Tensor, noise: Optional[torch.Tensor] = None):
    """log-prob / entropy of ``raw`` (pre-tanh) under NormalTanh(policy(obs)) -- the differentiable twin of the kernel's head.
    ``noise``: standard normals for Brax's sampled entropy term (default: fresh ``randn``)."""
    out = policy(obs_n)
    loc, sp = out.chunk(2, dim=-1)
    scale = torch.nn.functional.softplus(sp) + 0.001
    z = (raw - loc) / scale
    logn = -0.5 * z * z - torch.log(scale) - 0.5 * math.log(2 * math.pi)
    ldj = 2.0 * (math.log(2.0) - raw - torch.nn.functional.softplus(-2.0 * raw))
    logp = (logn - ldj).sum(-1)
    # Brax entropy: normal entropy + E[log det jacobian] estimated at a fresh sample
    sample = loc + scale * (torch.randn_like(loc) if noise is None else noise)
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale) + 2.0 * (math.log(2.0) - sample - torch.nn.functional.softplus(-2.0 * sample))).sum(-1)
    return logp, ent


def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_, discount):
    """brax ppo losses.compute_gae; inputs [T, N]."""
    mask = 1.0 - truncation
    values_tp1 = torch.cat([values[1:], bootstrap_value[None]], 0)
    deltas = (rewards + discount * (1 - termination) * values_tp1 - values) * mask
    acc = torch.zeros_like(bootstrap_value)
    vs_minus = []
    for t in reversed(range(rewards.shape[0])):
        acc = deltas[t] + discount * (1 - termination[t]) * mask[t] * lambda_ * acc
        vs_minus.append(acc)
    vs_minus = torch.stack(vs_minus[::-1], 0)
    vs = vs_minus + values
    vs_tp1 = torch.cat([vs[1:], bootstrap_value[None]], 0)
    adv = (rewards + discount * (1 - termination) * vs_tp1 - values) * mask
    return vs.detach(), adv.detach()


def shard_keys(seed: int, world: int, rank: int, n_per_rank: int) -> np.ndarray:
    """Per-env keys independent of the GPU count: split(seed, world * n) sliced per rank."""
    return jr.split(jr.PRNGKey(seed), world * n_per_rank)[rank * n_per_rank:(rank + 1) * n_per_rank]


class RolloutBuffers(dict):
    """A rank's rollout buffers ([T(+1), n, ...] per field, time-major like Brax's stacked Transition) as views of ONE flat fp32
    allocation, so that the exchange of SURVEY.md 8e is one ``all_gather_into_tensor`` of ``flat`` and nothing is packed."""

    FIELDS = ("obs_p", "obs_v", "raw", "logp", "reward", "done", "trunc")

    def __init__(self, T: int, n: int, dp: int, dv: int, na: int, device):
        shapes = {"obs_p": (T + 1, n, dp), "obs_v": (T + 1, n, dv), "raw": (T, n, na), "logp": (T, n), "reward": (T, n), "done": (T, n), "trunc": (T, n)}
        total = sum(int(np.prod(sh)) for sh in shapes.values())
        self.flat = torch.zeros(total, device=device)
        self.offsets, off = {}, 0
        super().__init__()
        for k in self.FIELDS:
            cnt = int(np.prod(shapes[k]))
            self[k] = self.flat[off:off + cnt].view(*shapes[k])
            self.offsets[k] = (off, shapes[k])
            off += cnt
        self.bytes = total * 4
        self.T, self.n = T, n
        # The policy observation is the first dp columns of the value observation (both reference envs: privileged_state =
        # hstack([state, ...]), joystick.py:596-615): when the owner says so, the exchange leaves obs_p at home -- the gathered part
        # of ``flat`` starts behind it -- and the learner reads the policy rows out of obs_v (OduckRollout.obs_policy_ld).
        self.policy_prefix = False

    @property
    def skip(self) -> int:
        """Leading floats of ``flat`` that do not travel in the all-gather."""
        return self.offsets["obs_v"][0] if self.policy_prefix else 0


class GatheredRollout:
    """What one all-gather of every rank's ``RolloutBuffers.flat`` leaves behind: ``world`` blocks, block r = rank r's buffers.
    The device learner reads it in place (OduckRollout.block_envs / block_stride); ``[key]`` builds the time-major
    [T, world * n, ...] tensor of a field on demand (PyTorch twin learner, scalar summaries)."""

    def __init__(self, gathered: torch.Tensor, local: RolloutBuffers, world: int):
        self.flat, self.local, self.world = gathered, local, world
        self.skip = local.skip
        self.block_stride = local.flat.numel() - self.skip
        self._cache: Dict[str, torch.Tensor] = {}

    def blocks(self, key: str) -> torch.Tensor:
        """[world, T(+1), n, ...] view of a field."""
        if key == "obs_p" and self.skip:                                # stayed at home: the leading columns of obs_v
            return self.blocks("obs_v")[..., :self.local["obs_p"].shape[-1]]
        off, shape = self.local.offsets[key]
        cnt = int(np.prod(shape))
        return self.flat.view(self.world, self.block_stride)[:, off - self.skip:off - self.skip + cnt].view(self.world, *shape)

    def __getitem__(self, key: str) -> torch.Tensor:
        if key not in self._cache:
            b = self.blocks(key)                                        # [world, T, n, ...] -> [T, world * n, ...]
            self._cache[key] = b.transpose(0, 1).reshape(b.shape[1], self.world * b.shape[2], *b.shape[3:]).contiguous()
        return self._cache[key]

    def keys(self):
        return self.local.keys()

    def items(self):
        return ((k, self[k]) for k in self.local.keys())


def all_gather_rollout(batch, world: int, out: Optional[torch.Tensor] = None):
    """The one exchange of a training step: every rank contributes its env shard.  ``RolloutBuffers`` are gathered
    with ONE collective on their flat allocation and stay in rank-major blocks (``GatheredRollout``); a plain dict of tensors is
    gathered field by field into [T, world * n, ...]."""
    if world == 1:
        return batch
    if isinstance(batch, RolloutBuffers):
        part = batch.flat[batch.skip:]
        g = out if out is not None else torch.empty(world * part.numel(), device=batch.flat.device)
        dist.all_gather_into_tensor(g, part)
        return GatheredRollout(g, batch, world)
    res = {}
    for k, v in batch.items():
        v = v.contiguous()
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v)
        res[k] = torch.cat(parts, dim=1)
    return res


class DeviceLearner:
    """Host mirror of include/oduck_ppo.h: one SGD step of Brax PPO per ``minibatch`` call, entirely on the device
    (gather/pack, tcgen05 forward + backward GEMMs, GAE/loss, clip + Adam).  Master weights live in the library's flat
    parameter vector (per net, per layer: W[in][out] like flax, then b); ``tensor()`` hands out zero-copy views."""

    def __init__(self, cfg: PPOConfig, policy: MLP, value: MLP, batch_envs: int, num_actions: int, device):
        from .joystick import _CudaView
        self._view_cls = _CudaView
        self.device = torch.device(device)
        c = capi.OduckPpoConfig()
        c.batch_envs, c.unroll, c.num_actions = int(batch_envs), int(cfg.unroll_length), int(num_actions)
        for net, mlp in ((0, policy), (1, value)):
            dims = [mlp.layers[0].in_features] + [l.out_features for l in mlp.layers]
            for i, d in enumerate(dims):
                (c.policy_dims if net == 0 else c.value_dims)[i] = d
        c.normalize_advantage = 1
        c.discounting, c.gae_lambda, c.clipping_epsilon, c.entropy_cost = cfg.discounting, cfg.gae_lambda, cfg.clipping_epsilon, cfg.entropy_cost
        c.reward_scaling, c.learning_rate, c.max_grad_norm = cfg.reward_scaling, cfg.learning_rate, cfg.max_grad_norm if cfg.max_grad_norm else 0.0
        c.adam_b1, c.adam_b2, c.adam_eps = 0.9, 0.999, 1e-8
        if cfg.learner_matmul not in ("fp32", "tf32"):
            raise ValueError("learner_matmul must be 'fp32' or 'tf32'")
        c.matmul_tf32 = 1 if cfg.learner_matmul == "tf32" else 0
        self.batch_envs = int(batch_envs)
        self.h = capi.PpoHandle(capi.load_cuda_library(), c, self.device.index or 0)
        self.params, self.grads = self.view("PARAMS"), self.view("GRADS")
        self.losses, self.step = self.view("LOSSES"), self.view("STEP")
        self.load_from(policy, value, reset_opt=True)

    def view(self, name: str) -> torch.Tensor:
        ptr, count, dt = self.h.buffer_info(name)
        with torch.cuda.device(self.device):
            return torch.as_tensor(self._view_cls(ptr, (count,), (1,), dt), device=self.device)

    def tensor(self, net: int, layer: int, which: int) -> torch.Tensor:
        off, r, c = self.h.param_info(net, layer, which)
        t = self.params[off:off + r * c]
        return t.view(r, c) if which == 0 else t

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @torch.no_grad()
    def load_from(self, policy: MLP, value: MLP, reset_opt: bool) -> None:
        for net, mlp in ((0, policy), (1, value)):
            for l, lin in enumerate(mlp.layers):
                self.tensor(net, l, 0).copy_(lin.weight.t())
                self.tensor(net, l, 1).copy_(lin.bias)
        self.h.set_params(self.params.data_ptr(), reset_opt, self._stream())

    @torch.no_grad()
    def store_to(self, policy: MLP, value: MLP) -> None:
        for net, mlp in ((0, policy), (1, value)):
            for l, lin in enumerate(mlp.layers):
                lin.weight.copy_(self.tensor(net, l, 0).t())
                lin.bias.copy_(self.tensor(net, l, 1))

    def policy_views(self):
        return [(self.tensor(0, l, 0), self.tensor(0, l, 1)) for l in range(4)]

    def policy_packed(self):
        return [self.h.packed_weights(0, l) for l in range(4)]

    def minibatch(self, rollout: "capi.OduckRollout", norm: "capi.OduckNormalizer", env_idx: int, noise: int = 0, key: int = 0, stages: int = capi.PPO_ALL) -> None:
        self.h.minibatch(rollout, norm, env_idx, noise, key, stages, self._stream())

    def prefetch(self, rollout: "capi.OduckRollout", norm: "capi.OduckNormalizer", next_env_idx: int) -> None:
        """Pack the NEXT minibatch's observation operands beside the current minibatch's kernels (call right after ``minibatch``)."""
        self.h.prefetch(rollout, norm, next_env_idx, self._stream())


def rollout_struct(batch) -> "capi.OduckRollout":
    """OduckRollout over the trainer's rollout buffers -- a dict of [T(+1), N, ...] tensors, or a ``GatheredRollout`` read in place
    as rank-major blocks (tensors must stay alive while the learner runs)."""
    ro = capi.OduckRollout()
    if isinstance(batch, GatheredRollout):
        src = batch.local
        T, n = src["reward"].shape
        ro.num_envs, ro.unroll, ro.block_envs, ro.block_stride = int(batch.world * n), int(T), int(n), int(batch.block_stride)
        ptr = lambda k: batch.flat.data_ptr() + 4 * (src.offsets[k][0] - batch.skip)     # noqa: E731  (block 0's field)
        if batch.skip:                                                                    # policy rows = leading columns of the value rows
            ro.obs_policy_ld = int(src["obs_v"].shape[-1])
    else:
        T, N = batch["reward"].shape
        for k, v in batch.items():
            if not v.is_contiguous() or v.dtype != torch.float32:
                raise ValueError(f"rollout tensor {k} must be contiguous float32")
        ro.num_envs, ro.unroll, ro.block_envs, ro.block_stride = int(N), int(T), 0, 0
        ptr = lambda k: batch[k].data_ptr()                                               # noqa: E731
    ro.obs_policy, ro.obs_value, ro.raw_action = ptr("obs_v" if ro.obs_policy_ld else "obs_p"), ptr("obs_v"), ptr("raw")
    ro.log_prob, ro.reward, ro.done, ro.truncation = ptr("logp"), ptr("reward"), ptr("done"), ptr("trunc")
    return ro


class Evaluator:
    """Brax ``acting.Evaluator`` + ``EvalWrapper`` for the fused env: ``num_eval_envs`` episodes of ``episode_length`` steps on
    their own env instance; rewards and metrics are summed per env while its FIRST episode is active (``active *= 1 - done``),
    exactly what ``eval/episode_reward`` / ``eval/episode_<metric>`` / ``eval/avg_episode_length`` report in the reference's
    progress callback (common/runner.py:56-66)."""

    def __init__(self, env, cfg: PPOConfig):
        self.cfg = cfg
        self.env = env.spawn()            # same compiled model (incl. a user xml_path), config, library, auto-reset
        self.n = cfg.num_eval_envs
        self.key = jr.PRNGKey(cfg.seed + 101)
        self.env.randomize(jr.split(jr.PRNGKey(cfg.seed + 102), self.n))

    @torch.no_grad()
    def run(self, weights: "PolicyWeights") -> Dict[str, float]:
        env, n, T = self.env, self.n, self.cfg.episode_length
        dev = env.device
        self.key, k_reset, k_act = jr.split(self.key, 3)
        st = env.reset(jr.split(k_reset, n))
        step_keys = jr.split(jr.split(k_act, T), n).view(np.int32)              # [T, n, 2]
        keys = torch.from_numpy(np.ascontiguousarray(step_keys)).to(dev)
        active = torch.ones(n, device=dev)
        ep_reward = torch.zeros(n, device=dev)
        ep_steps = torch.zeros(n, device=dev)
        ep_metrics = torch.zeros(n, len(env.METRICS), device=dev)
        met = env.buffer("METRICS")
        for t in range(T):
            act, _, _ = policy_forward(env, weights, None if self.cfg.deterministic_eval else keys[t], deterministic=self.cfg.deterministic_eval,
                                       obs=None if self.cfg.policy_obs_key == "state" else st.obs[self.cfg.policy_obs_key])
            st = env.step(st, act)
            ep_reward += st.reward.float() * active
            ep_metrics += met[:, :len(env.METRICS)].float() * active[:, None]
            ep_steps += active
            active = active * (1.0 - st.done.float())
        out = {"eval/episode_reward": float(ep_reward.mean()), "eval/episode_reward_std": float(ep_reward.std(unbiased=False)),
               "eval/avg_episode_length": float(ep_steps.mean())}
        for i, name in enumerate(env.METRICS):
            out[f"eval/episode_{name}"] = float(ep_metrics[:, i].mean())
        return out


class PPOTrainer:
    def __init__(self, env, cfg: PPOConfig, rank: int = 0, world: int = 1, progress_fn: Optional[Callable] = None,
                 policy_params_fn: Optional[Callable] = None):
        self.env, self.cfg, self.rank, self.world = env, cfg, rank, world
        self.progress_fn, self.policy_params_fn = progress_fn, policy_params_fn
        self.n_local = cfg.num_envs // world
        dev = env.device
        torch.manual_seed(cfg.seed)                                     # identical initial weights on every rank
        na = env.action_size
        self.policy = MLP([env.observation_size[cfg.policy_obs_key][0], *cfg.policy_hidden_layer_sizes, 2 * na]).to(dev)
        self.value = MLP([env.observation_size[cfg.value_obs_key][0], *cfg.value_hidden_layer_sizes, 1]).to(dev)
        self.opt = torch.optim.Adam(list(self.policy.parameters()) + list(self.value.parameters()), lr=cfg.learning_rate, capturable=dev.type == "cuda")
        self.stats = {k: RunningStats(env.observation_size[k][0], dev) for k in (cfg.policy_obs_key, cfg.value_obs_key)}
        self._mean32 = {k: torch.zeros(env.observation_size[k][0], device=dev) for k in self.stats}
        self.learner = cfg.learner if cfg.learner != "auto" else ("device" if dev.type == "cuda" else "torch")
        self.dev_learner: Optional[DeviceLearner] = None
        if self.learner == "device":
            mode = cfg.update_mode if cfg.update_mode != "auto" else ("sharded" if world > 1 else "replicated")
            n_update = self.n_local if (mode == "sharded" and world > 1) else cfg.num_envs
            self.dev_learner = DeviceLearner(cfg, self.policy, self.value, n_update // cfg.num_minibatches, na, dev)
        self.weights = PolicyWeights(self.policy, env.observation_size[cfg.policy_obs_key][0], dev,
                                     external=self.dev_learner.policy_views() if self.dev_learner else None,
                                     packed=self.dev_learner.policy_packed() if self.dev_learner else None)
        self._roll = None                 # persistent rollout buffers + CUDA graphs (device learner on CUDA)
        self._ones = {k: torch.ones(env.observation_size[k][0], device=dev) for k in self.stats}
        self._zeros = {k: torch.zeros(env.observation_size[k][0], device=dev) for k in self.stats}
        self.key = jr.PRNGKey(cfg.seed + 17)
        self.env_steps = 0
        self.P = int(cfg.rollout_pipeline)
        if self.P <= 0:
            self.P = 2 if (dev.type == "cuda" and self.n_local >= 2048 and self.n_local % 2 == 0) else 1
        if self.P == 1:
            self._envs = [env]
            env.randomize(shard_keys(cfg.seed + 1, world, rank, self.n_local))
            self.state = env.reset(shard_keys(cfg.seed, world, rank, self.n_local))
        else:
            if self.n_local % self.P:
                raise ValueError("num_envs per rank must be a multiple of rollout_pipeline")
            m = self.n_local // self.P
            dr, rk = shard_keys(cfg.seed + 1, world, rank, self.n_local), shard_keys(cfg.seed, world, rank, self.n_local)
            self._envs = [env.spawn() for _ in range(self.P)]
            self.state = []
            for q, e in enumerate(self._envs):
                e.randomize(dr[q * m:(q + 1) * m])
                self.state.append(e.reset(rk[q * m:(q + 1) * m]))
        # kernels write the rollout buffers themselves (OduckRolloutSink) when the nets read the observations the step kernel
        # produces under those names; any other pairing keeps the copy path
        self._use_sink = cfg.policy_obs_key == "state" and cfg.value_obs_key == "privileged_state" and cfg.kernel_rollout_writes
        self._sink_buf = None
        self.timing = {"rollout_ms": 0.0, "gather_ms": 0.0, "update_ms": 0.0}
        self.evaluator = Evaluator(env, cfg) if (rank == 0 and cfg.num_eval_envs > 0) else None

    # ------------------------------------------------------------------ A17: unroll
    def _rollout_keys(self) -> np.ndarray:
        """Sampling keys of one unroll, [T, n_local, 2]: split(step_key_t, world * n) sliced per rank (independent of the GPU count)."""
        pre = getattr(self, "_prefetched_keys", None)
        if pre is not None:                                   # made while the GPU was busy with the previous update
            self._prefetched_keys = None
            return pre
        T, n = self.cfg.unroll_length, self.n_local
        self.key, sub = jr.split(self.key, 2)
        step_keys = jr.split(sub, T)
        return np.ascontiguousarray(jr.split(step_keys, self.world * n)[:, self.rank * n:(self.rank + 1) * n]).view(np.int32)

    def _new_buffers(self) -> RolloutBuffers:
        cfg, env = self.cfg, self.env
        buf = RolloutBuffers(cfg.unroll_length, self.n_local, env.observation_size[cfg.policy_obs_key][0], env.observation_size[cfg.value_obs_key][0],
                             env.action_size, env.device)
        # obs["privileged_state"] starts with obs["state"] in both reference envs (joystick.py:596-615, standing.py): the exchange
        # then carries one observation tensor (device learner only: the PyTorch twin reads obs_p as a tensor of its own)
        buf.policy_prefix = bool(getattr(env, "privileged_obs_has_state_prefix", False) and cfg.policy_obs_key == "state"
                                 and cfg.value_obs_key == "privileged_state" and self.dev_learner is not None)
        return buf

    def _attach(self, buf) -> None:
        if self._sink_buf is not buf:
            m = self.n_local // self.P
            for q, e in enumerate(self._envs):
                attach_rollout_sink(e, buf, q * m)
            self._sink_buf = buf

    def _rollout_step(self, buf, t, st, keys_t):
        pk, vk = self.cfg.policy_obs_key, self.cfg.value_obs_key
        if self._use_sink:
            self._attach(buf)
            return rollout_step(self.env, self.weights, keys_t, t)
        buf["obs_p"][t].copy_(st.obs[pk]); buf["obs_v"][t].copy_(st.obs[vk])
        act, raw, logp = policy_forward(self.env, self.weights, keys_t, deterministic=False, obs=None if pk == "state" else st.obs[pk])
        st = self.env.step(st, act)
        buf["raw"][t].copy_(raw); buf["logp"][t].copy_(logp)
        buf["reward"][t].copy_(st.reward); buf["done"][t].copy_(st.done); buf["trunc"][t].copy_(st.info["truncation"])
        return st

    def rollout(self) -> Dict[str, torch.Tensor]:
        cfg, env = self.cfg, self.env
        T, n = cfg.unroll_length, self.n_local
        pk, vk = cfg.policy_obs_key, cfg.value_obs_key
        graphed = self.dev_learner is not None and cfg.cuda_graph and env.device.type == "cuda"
        if graphed:
            # static normaliser buffers (updated in place by update()), static weights (the learner's own): capturable
            mean, std = (self._mean32[pk], self.stats[pk].std) if cfg.normalize_observations else (self._zeros[pk], self._ones[pk])
        else:
            mean, std = (self.stats[pk].mean32, self.stats[pk].std) if cfg.normalize_observations else (torch.zeros_like(self.stats[pk].std), torch.ones_like(self.stats[pk].std))
        self.weights.refresh(mean, std)
        keys = self._rollout_keys()
        if self.P > 1:
            return self._rollout_pipelined(keys, graphed)
        st = self.state
        if not graphed:
            buf = self._new_buffers()
            for t in range(T):
                st = self._rollout_step(buf, t, st, torch.from_numpy(keys[t]))
            if not self._use_sink:
                buf["obs_p"][T].copy_(st.obs[pk]); buf["obs_v"][T].copy_(st.obs[vk])
            self.state = st
            return buf
        # CUDA-graph path: the T steps of an unroll are captured once (per-step graphs over persistent buffers); an unroll is
        # then one key upload + T replays, which keeps the ranks kernel-bound when many of them share the host's cores
        if self._roll is None:
            R = {"buf": self._new_buffers(), "keys": torch.empty(T, n, 2, dtype=torch.int32, device=env.device),
                 "host_keys": torch.empty(T, n, 2, dtype=torch.int32).pin_memory(), "graphs": None}
            self._roll = R
        R = self._roll
        R["host_keys"].copy_(torch.from_numpy(keys))
        R["keys"].copy_(R["host_keys"], non_blocking=True)
        buf = R["buf"]
        if R["graphs"] is None:
            for t in range(T):                               # first unroll: eager (also warms every kernel up before the capture)
                st = self._rollout_step(buf, t, st, R["keys"][t])
            if not self._use_sink:
                buf["obs_p"][T].copy_(st.obs[pk]); buf["obs_v"][T].copy_(st.obs[vk])
            torch.cuda.synchronize(env.device)
            pool, graphs = torch.cuda.graph_pool_handle(), []
            for t in range(T):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    self._rollout_step(buf, t, st, R["keys"][t])
                    if t == T - 1 and not self._use_sink:
                        buf["obs_p"][T].copy_(st.obs[pk]); buf["obs_v"][T].copy_(st.obs[vk])
                graphs.append(g)
            R["graphs"] = graphs
        else:
            for g in R["graphs"]:
                g.replay()
        self.state = st
        return buf

    def _sub_step(self, buf, t, q, st, keys_t):
        """_rollout_step for sub-batch q: its env handle, its slice of the rollout buffers."""
        pk, vk = self.cfg.policy_obs_key, self.cfg.value_obs_key
        m = self.n_local // self.P
        sl = slice(q * m, (q + 1) * m)
        e = self._envs[q]
        if self._use_sink:
            self._attach(buf)
            return rollout_step(e, self.weights, keys_t[sl], t)
        buf["obs_p"][t, sl].copy_(st.obs[pk]); buf["obs_v"][t, sl].copy_(st.obs[vk])
        act, raw, logp = policy_forward(e, self.weights, keys_t[sl], deterministic=False, obs=None if pk == "state" else st.obs[pk])
        st = e.step(st, act)
        buf["raw"][t, sl].copy_(raw); buf["logp"][t, sl].copy_(logp)
        buf["reward"][t, sl].copy_(st.reward); buf["done"][t, sl].copy_(st.done); buf["trunc"][t, sl].copy_(st.info["truncation"])
        if t == self.cfg.unroll_length - 1:
            buf["obs_p"][t + 1, sl].copy_(st.obs[pk]); buf["obs_v"][t + 1, sl].copy_(st.obs[vk])
        return st

    def _rollout_pipelined(self, keys: np.ndarray, graphed: bool) -> Dict[str, torch.Tensor]:
        """rollout() for rollout_pipeline = P > 1: every sub-batch runs its own chain of T (actor forward + env.step) steps; on CUDA
        each chain is T graphs replayed on the sub-batch's stream, forked from / joined to the caller's stream with events."""
        T, P, dev = self.cfg.unroll_length, self.P, self.env.device
        if not graphed:
            buf = self._new_buffers()
            for t in range(T):
                kt = torch.from_numpy(keys[t])
                for q in range(P):
                    self.state[q] = self._sub_step(buf, t, q, self.state[q], kt)
            return buf
        if self._roll is None:
            self._roll = {"buf": self._new_buffers(), "keys": torch.empty(T, self.n_local, 2, dtype=torch.int32, device=dev),
                          "host_keys": torch.empty(T, self.n_local, 2, dtype=torch.int32).pin_memory(), "graphs": None,
                          "streams": [torch.cuda.Stream(device=dev) for _ in range(P)]}
        R = self._roll
        R["host_keys"].copy_(torch.from_numpy(keys))
        R["keys"].copy_(R["host_keys"], non_blocking=True)
        buf = R["buf"]
        if R["graphs"] is None:
            for t in range(T):                               # first unroll: eager on the caller's stream (warms every kernel up)
                for q in range(P):
                    self.state[q] = self._sub_step(buf, t, q, self.state[q], R["keys"][t])
            torch.cuda.synchronize(dev)
            graphs = []
            for q in range(P):
                pool, chain = torch.cuda.graph_pool_handle(), []
                for t in range(T):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool, stream=R["streams"][q]):
                        self._sub_step(buf, t, q, self.state[q], R["keys"][t])
                    chain.append(g)
                graphs.append(chain)
            R["graphs"] = graphs
            torch.cuda.synchronize(dev)
            return buf
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)                                    # after the key upload and the previous update
        for t in range(T):
            for q in range(P):
                with torch.cuda.stream(R["streams"][q]):
                    if t == 0:
                        R["streams"][q].wait_event(fork)
                    R["graphs"][q][t].replay()
        for q in range(P):
            join = torch.cuda.Event()
            join.record(R["streams"][q])
            main.wait_event(join)
        return buf

    # ------------------------------------------------------------------ update
    def _minibatch_loss(self, mb: Dict[str, torch.Tensor], noise: Optional[torch.Tensor] = None):
        cfg = self.cfg
        pk, vk = cfg.policy_obs_key, cfg.value_obs_key
        if cfg.normalize_observations:
            obs_p = (mb["obs_p"] - self._mean32[pk]) / self.stats[pk].std
            obs_v = (mb["obs_v"] - self._mean32[vk]) / self.stats[vk].std
        else:
            obs_p, obs_v = mb["obs_p"], mb["obs_v"]
        values = self.value(obs_v).squeeze(-1)
        baseline, bootstrap = values[:-1], values[-1]
        trunc, done = mb["trunc"], mb["done"]
        termination = done * (1 - trunc)
        rewards = mb["reward"] * cfg.reward_scaling
        vs, adv = compute_gae(trunc, termination, rewards, baseline.detach(), bootstrap.detach(), cfg.gae_lambda, cfg.discounting)
        adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)      # jnp.std: population std
        logp, ent = torch_policy_logprob(self.policy, obs_p[:-1], mb["raw"], noise)
        rho = torch.exp(logp - mb["logp"])
        policy_loss = -torch.min(rho * adv, rho.clamp(1 - cfg.clipping_epsilon, 1 + cfg.clipping_epsilon) * adv).mean()
        v_loss = ((vs - baseline) ** 2).mean() * 0.5 * 0.5
        ent_mean = ent.mean()
        return policy_loss + v_loss - cfg.entropy_cost * ent_mean, policy_loss, v_loss, ent_mean

    def _params(self):
        return list(self.policy.parameters()) + list(self.value.parameters())

    def _apply_grads(self):
        torch.nn.utils.clip_grad_norm_(self._params(), self.cfg.max_grad_norm, foreach=True)
        self.opt.step()

    def _build_graphs(self, mb_shapes: Dict[str, torch.Size], sharded: bool) -> None:
        """Capture (forward + backward) and (clip + Adam) of one minibatch; the sharded mode all-reduces the flat gradient in between."""
        dev = self.env.device
        self._static = {k: torch.zeros(shp, device=dev) for k, shp in mb_shapes.items()}
        self._out = torch.zeros(4, device=dev)
        snap = [p.detach().clone() for p in self._params()]
        for prm in self._params():
            prm.grad = torch.zeros_like(prm)
        # The Adam state must EXIST before the capture (the captured step updates these very tensors by address) and must never be
        # replaced afterwards.  A fresh optimiser has none yet: one throw-away step on zero gradients creates it; then the state
        # is snapshotted by VALUE (state_dict() hands out references), warm-up + capture run, and the snapshot is copied back in
        # place -- zero moments / step 0 for a fresh optimiser, the loaded moments after load().
        fresh = len(self.opt.state) == 0
        if fresh:
            self.opt.step()
            with torch.no_grad():
                for prm, sv in zip(self._params(), snap):
                    prm.copy_(sv)
        opt_snap = {prm: {k: (v.detach().clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in self.opt.state[prm].items()} for prm in self._params()}
        if fresh:
            for st_ in opt_snap.values():
                for v in st_.values():
                    if torch.is_tensor(v):
                        v.zero_()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):                                   # warm-up outside capture (allocator, cuBLAS handles, Adam state)
                for prm in self._params():
                    prm.grad.zero_()
                loss, *_ = self._minibatch_loss(self._static)
                loss.backward()
                self._apply_grads()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._g_fb, self._g_step = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_fb):
            for prm in self._params():
                prm.grad.zero_()
            loss, pl, vl, en = self._minibatch_loss(self._static)
            loss.backward()
            self._out.copy_(torch.stack([loss.detach(), pl.detach(), vl.detach(), en.detach()]))
        with torch.cuda.graph(self._g_step):
            self._apply_grads()
        with torch.no_grad():                                    # undo the warm-up / capture-time updates
            for prm, sv in zip(self._params(), snap):
                prm.copy_(sv)
            for prm, saved in opt_snap.items():                  # in place: the graph keeps updating these tensors
                for k, v in saved.items():
                    if torch.is_tensor(v):
                        self.opt.state[prm][k].copy_(v)
                    else:
                        self.opt.state[prm][k] = v
        self._graph_sharded = sharded

    def update(self, batch: Dict[str, torch.Tensor], sharded: bool = False) -> Dict[str, float]:
        """4 epochs x num_minibatches clipped-PPO updates on ``batch`` ([T(+1), N, ...]).  sharded: ``batch`` is this rank's env shard and
        gradients are averaged over ranks (Brax's pmean); otherwise ``batch`` is the gathered global batch and every rank does the same update."""
        cfg = self.cfg
        pk, vk = cfg.policy_obs_key, cfg.value_obs_key
        blocked = isinstance(batch, GatheredRollout)
        if cfg.normalize_observations:
            if blocked and self.dev_learner is not None:          # [world, T + 1, n, d]: the same rows, rank-major (the moments are sums over rows;
                                                                  # the PyTorch twin keeps the time-major order = the single-process bits)
                self.stats[pk].update(batch.blocks("obs_p")[:, :-1], reduce=False); self.stats[vk].update(batch.blocks("obs_v")[:, :-1], reduce=False)
            else:
                self.stats[pk].update(batch["obs_p"][:-1], reduce=sharded); self.stats[vk].update(batch["obs_v"][:-1], reduce=sharded)
        for k in (pk, vk):
            self._mean32[k].copy_(self.stats[k].mean32)
        T, N = (batch.local["reward"].shape[0], batch.world * batch.local["reward"].shape[1]) if blocked else batch["reward"].shape
        gen = torch.Generator(device="cpu").manual_seed(cfg.seed + self.env_steps + (self.rank if sharded else 0))
        if self.dev_learner is not None:
            return self._update_device(batch, sharded, gen)
        mb = N // cfg.num_minibatches
        use_graph = cfg.cuda_graph and self.env.device.type == "cuda"
        keys = ("obs_p", "obs_v", "raw", "logp", "reward", "done", "trunc")
        if use_graph and getattr(self, "_g_fb", None) is None:
            self._build_graphs({k: batch[k][:, :mb].shape for k in keys}, sharded)
        params = self._params()
        for _ in range(cfg.num_updates_per_batch):
            perm = torch.randperm(N, generator=gen).to(batch["reward"].device)
            for i in range(cfg.num_minibatches):
                idx = perm[i * mb:(i + 1) * mb]
                if use_graph:
                    for k in keys:
                        torch.index_select(batch[k], 1, idx, out=self._static[k])
                    self._g_fb.replay()
                    if sharded:
                        flat = torch.cat([p.grad.flatten() for p in params])
                        dist.all_reduce(flat)
                        flat /= self.world
                        off = 0
                        for p in params:
                            p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
                    self._g_step.replay()
                else:
                    loss, pl, vl, en = self._minibatch_loss({k: batch[k][:, idx] for k in keys})
                    self.opt.zero_grad(set_to_none=True)
                    loss.backward()
                    if sharded:
                        for p in params:
                            dist.all_reduce(p.grad); p.grad /= self.world
                    self._apply_grads()
        if use_graph:
            o = self._out.tolist()
            return dict(loss=o[0], policy_loss=o[1], v_loss=o[2], entropy=o[3])
        return dict(loss=float(loss.detach()), policy_loss=float(pl.detach()), v_loss=float(vl.detach()), entropy=float(en.detach()))

    def _update_device(self, batch: Dict[str, torch.Tensor], sharded: bool, gen: torch.Generator) -> Dict[str, float]:
        """The update through include/oduck_ppo.h: one library call per minibatch, no host sync until the metrics are read."""
        cfg, L = self.cfg, self.dev_learner
        pk, vk = cfg.policy_obs_key, cfg.value_obs_key
        blocked = isinstance(batch, GatheredRollout)
        T, N = (batch.local["reward"].shape[0], batch.world * batch.local["reward"].shape[1]) if blocked else batch["reward"].shape
        B = N // cfg.num_minibatches
        if B != L.batch_envs:
            raise ValueError(f"device learner was built for {L.batch_envs} envs per minibatch, got {B}")
        dev = self.env.device
        ro = rollout_struct(batch)
        nm = capi.OduckNormalizer()
        if cfg.normalize_observations:
            nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = self._mean32[pk].data_ptr(), self.stats[pk].std.data_ptr(), self._mean32[vk].data_ptr(), self.stats[vk].std.data_ptr()
        else:
            nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = self._zeros[pk].data_ptr(), self._ones[pk].data_ptr(), self._zeros[vk].data_ptr(), self._ones[vk].data_ptr()
        n_mb = cfg.num_updates_per_batch * cfg.num_minibatches
        self.key, sub = jr.split(self.key, 2)
        if sharded:
            sub = jr.split(sub, self.world)[self.rank]
        keys = torch.from_numpy(jr.split(sub, n_mb).view(np.int32).copy()).to(dev)        # entropy-sample keys, one per minibatch
        keep = [keys]
        FLB = capi.PPO_STAGE_FORWARD | capi.PPO_STAGE_LOSS | capi.PPO_STAGE_BACKWARD
        # The minibatch step reads a static index / key buffer and is captured once into CUDA graphs (the library forks and
        # joins its side streams inside the capture); a minibatch is then two small copies + one replay (two around the
        # gradient all-reduce when sharded).  Only with the persistent rollout buffers of the graphed unroll (static pointers).
        graphed = cfg.cuda_graph and self._roll is not None and (batch is self._roll["buf"] or (blocked and batch.local is self._roll["buf"] and batch.flat is self._roll.get("gathered")))
        if graphed and getattr(self, "_upd", None) is None:
            nmb, E = cfg.num_minibatches, cfg.num_updates_per_batch
            U = {"idx": torch.zeros(B, dtype=torch.int32, device=dev), "key": torch.zeros(2, dtype=torch.int32, device=dev), "ro": ro, "nm": nm,
                 "perm": torch.zeros(E * N, dtype=torch.int32, device=dev), "keys": torch.zeros(E * nmb, 2, dtype=torch.int32, device=dev)}
            mb_idx = lambda j: U["perm"].data_ptr() + 4 * ((j // nmb) * N + (j % nmb) * B)        # noqa: E731  minibatch j = (epoch, i): slice i of epoch's permutation
            mb_key = lambda j: U["keys"].data_ptr() + 8 * j                                        # noqa: E731
            # plain kernel nodes only: the two-kernel reduce / Adam tail.  (The fused cooperative launch can be captured on this
            # driver as a kernel node with the cooperative attribute, but inside the graph it ran 49 us against 29 + 18 us for the two
            # plain kernels and serialised the side streams: measured on B200, profiles/r02d_launches_ppo_tf32.csv.)
            try:
                L.minibatch(ro, nm, U["idx"].data_ptr(), 0, U["key"].data_ptr(), capi.PPO_STAGE_FORWARD)    # warm-up outside the capture
                torch.cuda.synchronize(dev)
                U["graphs"], U["epoch"] = [], None
                if sharded and os.environ.get("ODUCK_PPO_GRAPH_NCCL", "1") != "0":
                    # sharded update, ONE graph for the whole update with the gradient all-reduces inside: forward / loss / backward,
                    # NCCL all-reduce (average) of the flat gradient, clip + Adam, epochs x num_minibatches times -- NCCL collectives
                    # are capturable, so the host issues two copies and one replay per update instead of six calls per minibatch
                    dist.all_reduce(torch.zeros(8, device=dev), op=dist.ReduceOp.AVG)       # communicator set up outside the capture
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for j in range(E * nmb):
                            L.minibatch(ro, nm, mb_idx(j), 0, mb_key(j), FLB)
                            if j + 1 < E * nmb:
                                L.prefetch(ro, nm, mb_idx(j + 1))
                            dist.all_reduce(L.grads, op=dist.ReduceOp.AVG)
                            L.minibatch(ro, nm, mb_idx(j), 0, mb_key(j), capi.PPO_STAGE_ADAM)
                    U["epoch"] = g
                elif sharded:
                    for stg in (FLB, capi.PPO_STAGE_ADAM):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            L.minibatch(ro, nm, U["idx"].data_ptr(), 0, U["key"].data_ptr(), stg)
                        U["graphs"].append(g)
                else:
                    # ONE graph for the whole update: the epochs x num_minibatches SGD steps back to back, minibatch (epoch, i) reading
                    # slice i of that epoch's permutation in the static index buffer -- per update two small copies and one replay
                    # instead of three host calls per minibatch (and no idle GPU while the host shuffles between epochs)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for j in range(E * nmb):
                            L.minibatch(ro, nm, mb_idx(j), 0, mb_key(j), capi.PPO_ALL | capi.PPO_NO_COOP)
                            if j + 1 < E * nmb:                          # the next minibatch's inputs are packed beside this one's kernels
                                L.prefetch(ro, nm, mb_idx(j + 1))
                    U["epoch"] = g
            except Exception as e:                                                       # capture unsupported: stay eager
                import sys
                print(f"[ppo] minibatch graph capture unavailable ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
                U["graphs"], U["epoch"] = None, None
            self._upd = U
        U = getattr(self, "_upd", None) if graphed else None
        perms = [torch.randperm(N, generator=gen).to(torch.int32) for _ in range(cfg.num_updates_per_batch)]   # (the same draws, in the same order, as epoch by epoch)
        if U is not None and U.get("epoch") is not None:
            U["perm"].copy_(torch.cat(perms), non_blocking=True)
            U["keys"].copy_(keys, non_blocking=True)
            U["epoch"].replay()
            perms = []
        for e, perm in enumerate(perms):
            perm = perm.to(dev)
            keep.append(perm)
            for i in range(cfg.num_minibatches):
                j = e * cfg.num_minibatches + i
                if U is not None and U["graphs"]:
                    U["idx"].copy_(perm[i * B:(i + 1) * B], non_blocking=True)
                    U["key"].copy_(keys[j], non_blocking=True)
                    U["graphs"][0].replay()
                    if sharded:
                        dist.all_reduce(L.grads)
                        L.grads.div_(self.world)
                        U["graphs"][1].replay()
                    continue
                idx = perm.data_ptr() + 4 * i * B
                key = keys.data_ptr() + 8 * j
                nxt = perm.data_ptr() + 4 * (i + 1) * B if i + 1 < cfg.num_minibatches else 0
                if sharded:
                    L.minibatch(ro, nm, idx, 0, key, FLB)
                    if nxt:
                        L.prefetch(ro, nm, nxt)
                    dist.all_reduce(L.grads)
                    L.grads.div_(self.world)
                    L.minibatch(ro, nm, idx, 0, key, capi.PPO_STAGE_ADAM)
                else:
                    L.minibatch(ro, nm, idx, 0, key, capi.PPO_ALL if cfg.learner_fused_tail else (capi.PPO_ALL | capi.PPO_NO_COOP))
                    if nxt:
                        L.prefetch(ro, nm, nxt)
        [e.handle.policy_invalidate() for e in self._envs]                                                # the actor repacks the new weights on its next forward
        self._prefetched_keys = self._rollout_keys()                                       # host work of the next unroll, under the GPU's update
        o = L.losses.tolist()                                                              # host sync: the update is done, `keep` may go
        del keep
        return dict(loss=o[0], policy_loss=o[1], v_loss=o[2], entropy=o[3], clip_fraction=o[5])

    def training_step(self) -> Dict[str, float]:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.env.device.type == "cuda" else None
        if ev: ev[0].record()
        local = self.rollout()
        if ev: ev[1].record()
        mode = self.cfg.update_mode
        if mode == "auto":
            mode = "sharded" if (self.world > 1 and self.env.device.type == "cuda") else "replicated"
        sharded = mode == "sharded" and self.world > 1
        self.last_update_mode = "sharded" if sharded else ("replicated" if self.world > 1 else "single")
        if sharded or self.world == 1:
            batch = local
        else:
            out = None
            if self._roll is not None and local is self._roll["buf"]:            # persistent (graph-static) gather buffer
                if self._roll.get("gathered") is None:
                    self._roll["gathered"] = torch.empty(self.world * (local.flat.numel() - local.skip), device=local.flat.device)
                out = self._roll["gathered"]
            batch = all_gather_rollout(local, self.world, out)
        if ev: ev[2].record()
        summary = torch.stack([batch["reward"].mean(), batch["done"].mean()])      # enqueued ahead of the update: read after it without another round trip
        m = self.update(batch, sharded=sharded)
        if ev:
            ev[3].record(); torch.cuda.synchronize()
            self.timing = {"rollout_ms": ev[0].elapsed_time(ev[1]), "gather_ms": ev[1].elapsed_time(ev[2]), "update_ms": ev[2].elapsed_time(ev[3])}
        self.env_steps += self.cfg.num_envs * self.cfg.unroll_length
        m["reward_per_step"], m["episode_done_rate"] = summary.tolist()             # (this rank's shard in sharded mode)
        return m

    def evaluate(self) -> Dict[str, float]:
        """One evaluation epoch with the current policy and normaliser (rank 0; Brax ``evaluator.run_evaluation``)."""
        pk = self.cfg.policy_obs_key
        if self.dev_learner is None:
            mean, std = (self.stats[pk].mean32, self.stats[pk].std) if self.cfg.normalize_observations else (self._zeros[pk], self._ones[pk])
        else:
            mean, std = (self._mean32[pk], self.stats[pk].std) if self.cfg.normalize_observations else (self._zeros[pk], self._ones[pk])
        self.weights.refresh(mean, std)
        return self.evaluator.run(self.weights)

    def reset_training_envs(self) -> None:
        """Hard reset of every training env with fresh keys (Brax ppo.train: ``env_state = reset_fn(key_envs)`` after each training
        epoch when num_resets_per_eval > 0).  Keys are split(seed', world * n) sliced per rank: independent of the GPU count."""
        self._hard_resets = getattr(self, "_hard_resets", 0) + 1
        rk = shard_keys(self.cfg.seed + 7919 * self._hard_resets, self.world, self.rank, self.n_local)
        if self.P == 1:
            self.state = self.env.reset(rk)
        else:
            m = self.n_local // self.P
            for q, e in enumerate(self._envs):
                self.state[q] = e.reset(rk[q * m:(q + 1) * m])

    def train(self):
        """Brax ``ppo.train``'s outer loop: an evaluation of the initial policy (num_evals > 1), then ``num_evals - 1`` iterations of
        [max(num_resets_per_eval, 1) training epochs, each ending in a hard env reset if num_resets_per_eval > 0] + evaluation +
        ``progress_fn`` / ``policy_params_fn`` (rank 0).  An epoch is ceil(num_timesteps / (iterations x epochs x env-steps per
        training step)) training steps, so at least ``num_timesteps`` env-steps are run, like Brax."""
        cfg = self.cfg
        per_step = cfg.num_envs * cfg.unroll_length
        iters, epochs = max(cfg.num_evals - 1, 1), max(cfg.num_resets_per_eval, 1)
        steps_per_epoch = max(1, -(-cfg.num_timesteps // (iters * epochs * per_step)))

        def report(m):
            if self.rank != 0:
                return
            if self.evaluator is not None:
                ev = self.evaluate()
            elif m is not None:                              # no eval envs: extrapolate the behaviour policy's mean step reward
                ev = {"eval/episode_reward": m["reward_per_step"] * cfg.episode_length, "eval/episode_reward_std": 0.0}
            else:
                return
            metrics = dict(ev)
            if m is not None:
                metrics.update({f"training/{k}": v for k, v in m.items()}, **{f"time/{k}": v for k, v in self.timing.items()})
            if self.progress_fn:
                self.progress_fn(self.env_steps, metrics)
            if m is not None and self.policy_params_fn:
                self.policy_params_fn(self.env_steps, None, self.params())

        if cfg.num_evals > 1:
            report(None)                                     # the initial policy (Brax: progress_fn(0, metrics))
        for _ in range(iters):
            for _ in range(epochs):
                for _ in range(steps_per_epoch):
                    m = self.training_step()
                if cfg.num_resets_per_eval > 0:
                    self.reset_training_envs()
            report(m)
        return self.params()

    def params(self):
        pk = self.cfg.policy_obs_key
        if self.dev_learner is not None:
            self.dev_learner.store_to(self.policy, self.value)
        return {"normalizer": {k: {"mean": s.mean32.cpu(), "std": s.std.cpu(), "count": float(s.count)} for k, s in self.stats.items()},
                "policy": {k: v.cpu() for k, v in self.policy.state_dict().items()}, "value": {k: v.cpu() for k, v in self.value.state_dict().items()},
                "optimizer": self.opt.state_dict() if self.dev_learner is None else
                {"device_adam": {"m": self.dev_learner.view("ADAM_M").cpu(), "v": self.dev_learner.view("ADAM_V").cpu(), "step": int(self.dev_learner.step.item())}},
                "env_steps": self.env_steps, "policy_obs_key": pk}

    def load(self, params) -> None:
        self.policy.load_state_dict(params["policy"]); self.value.load_state_dict(params["value"])
        if self.dev_learner is not None:
            L = self.dev_learner
            L.load_from(self.policy, self.value, reset_opt=True)
            ad = params["optimizer"].get("device_adam") if isinstance(params["optimizer"], dict) else None
            if ad is not None:
                L.view("ADAM_M").copy_(ad["m"].to(self.env.device)); L.view("ADAM_V").copy_(ad["v"].to(self.env.device)); L.step.fill_(int(ad["step"]))
            [e.handle.policy_invalidate() for e in self._envs]
        else:
            self.opt.load_state_dict(params["optimizer"])
            self._g_fb = self._g_step = None                  # load_state_dict REPLACES the Adam state tensors a captured step updates by address: capture again on the next update
        for k, s in self.stats.items():
            d = params["normalizer"][k]
            s.mean = d["mean"].float().to(self.env.device); s.std.copy_(d["std"].to(self.env.device)); s.count = torch.tensor(float(d["count"]), device=self.env.device)
            s.m2 = (s.std ** 2) * s.count
            self._mean32[k].copy_(s.mean32)                   # the static buffer the graphs / the learner read
        self.env_steps = int(params["env_steps"])
