"""The on-device PPO learner step (include/oduck_ppo.h) against its checker, the PyTorch fp32 twin.

Floating-point path => the oracle is a plain PyTorch fp32 restatement of Brax's compute_ppo_loss + optax clip/Adam
(ppo.py: compute_gae, torch_policy_logprob; brax/training/agents/ppo/losses.py is an un-vendored dependency of the
reference, call site playground/common/runner.py:104-118).  Stated tolerances (3xTF32 GEMMs, fp32 everywhere else):
head outputs 5e-5 abs, losses 3e-5, gradients 5e-4 (tensor cores) / 5e-5 (CUDA-core twin) of the gradient norm per tensor,
parameters after Adam 2e-6 abs for all but a 1e-4 fraction of sign-sensitive elements.
Every stage is run twice: tensor cores (product path) and the CUDA-core twin of the GEMM (ODUCK_PPO_DEBUG_SIMT), so a
failure localises to the MMA/descriptor code or to the operand layouts / epilogues.
"""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest
import torch

from open_duck_playground_b200 import capi, ppo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "oduck_ppo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oduck_ppo_[a-z_]+)\s*\(", text)))


def test_cuda_library_exports_the_learner_surface():
    path = capi.cuda_library_path()
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = capi.Library(path, is_device=True)
    names = _declared()
    assert "oduck_ppo_minibatch" in names and "oduck_ppo_create" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib.lib, n), n
    assert lib.has_ppo


def test_config_struct_matches_header():
    # field order / count of OduckPpoConfig in the header and in capi.py
    text = open(os.path.join(ROOT, "include", "oduck_ppo.h")).read()
    body = text[text.index("typedef struct OduckPpoConfig {"):text.index("} OduckPpoConfig;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct OduckPpoConfig {", "").strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        fields += [re.sub(r"\[.*\]", "", x).strip() for x in names.split(",")]
    assert fields == [f[0] for f in capi.OduckPpoConfig._fields_]


# ----------------------------------------------------------------------------------------------------- GPU parity
def _make(seed, N, T, nmb, dev, dp=101, dv=212):
    torch.manual_seed(seed)
    cfg = ppo.PPOConfig(num_envs=N, unroll_length=T, num_minibatches=nmb)
    policy = ppo.MLP([dp, 512, 256, 128, 28]).to(dev)
    value = ppo.MLP([dv, 512, 256, 128, 1]).to(dev)
    with torch.no_grad():                                        # non-zero biases so that their gradients / updates are exercised
        for m in (policy, value):
            for lin in m.layers:
                lin.bias.uniform_(-0.1, 0.1)
        # keep the head well conditioned (scale = softplus(~0) ~ 0.7 like a freshly initialised policy on normalised obs): with
        # scale -> 0.001 the log-prob reaches 1e5 and exp(logp - logp_old) amplifies fp32 rounding by 1e5 in BOTH implementations
        policy.layers[-1].weight.mul_(0.05)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    r = lambda *s: torch.randn(*s, generator=g)
    batch = {"obs_p": r(T + 1, N, dp) * 2 + 0.3, "obs_v": r(T + 1, N, dv) * 3 - 0.5, "raw": r(T, N, 14) * 0.8, "logp": r(T, N) * 0.3 - 12.0,
             "reward": torch.rand(T, N, generator=g) * 0.2, "done": (torch.rand(T, N, generator=g) < 0.1).float(), "trunc": torch.zeros(T, N)}
    batch["trunc"] = batch["done"] * (torch.rand(T, N, generator=g) < 0.3).float()
    batch = {k: v.to(dev).contiguous() for k, v in batch.items()}
    norm = {"pm": (r(dp) * 0.2).to(dev), "ps": (torch.rand(dp, generator=g) + 0.5).to(dev), "vm": (r(dv) * 0.2).to(dev), "vs": (torch.rand(dv, generator=g) + 0.5).to(dev)}
    return cfg, policy, value, batch, norm


def _twin_loss(cfg, policy, value, batch, norm, idx, noise):
    """Brax compute_ppo_loss on the minibatch made of env columns ``idx`` (PyTorch fp32)."""
    mb = {k: v[:, idx] for k, v in batch.items()}
    obs_p = (mb["obs_p"] - norm["pm"]) / norm["ps"]
    obs_v = (mb["obs_v"] - norm["vm"]) / norm["vs"]
    logits = policy(obs_p[:-1])
    values = value(obs_v).squeeze(-1)
    baseline, bootstrap = values[:-1], values[-1]
    trunc, done = mb["trunc"], mb["done"]
    termination = done * (1 - trunc)
    vs, adv_raw = ppo.compute_gae(trunc, termination, mb["reward"] * cfg.reward_scaling, baseline.detach(), bootstrap.detach(), cfg.gae_lambda, cfg.discounting)
    adv = (adv_raw - adv_raw.mean()) / (adv_raw.std(unbiased=False) + 1e-8)
    logp, ent = ppo.torch_policy_logprob(policy, obs_p[:-1], mb["raw"], noise)
    # behaviour log-prob close to the target one so that ratios are O(1) and both clip branches occur
    rho = torch.exp(logp - mb["logp"])
    pl = -torch.min(rho * adv, rho.clamp(1 - cfg.clipping_epsilon, 1 + cfg.clipping_epsilon) * adv).mean()
    vl = ((vs - baseline) ** 2).mean() * 0.25
    en = ent.mean()
    return pl + vl - cfg.entropy_cost * en, pl, vl, en, logits, values, adv, vs


def _flat(policy, value, grad=False):
    out = []
    for m in (policy, value):
        for lin in m.layers:
            w, b = (lin.weight.grad, lin.bias.grad) if grad else (lin.weight.detach(), lin.bias.detach())
            out += [w.t().reshape(-1), b.reshape(-1)]
    return torch.cat(out)


def _segments(L):
    return [(net, l, which, *L.h.param_info(net, l, which)) for net in (0, 1) for l in range(4) for which in (0, 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("simt", [True, False], ids=["cuda-core-twin", "tcgen05"])
@pytest.mark.parametrize("shape", [(48, 5, 3, 101, 212), (512, 20, 2, 101, 212), (96, 7, 2, 85, 153)], ids=["ragged-M80", "B256xT20", "standing-dims-85-153"])
def test_learner_stages_match_torch(simt, shape):
    N, T, nmb, dp, dv = shape
    dev = torch.device("cuda:0")
    cfg, policy, value, batch, norm = _make(3, N, T, nmb, dev, dp, dv)
    B = N // nmb
    L = ppo.DeviceLearner(cfg, policy, value, B, 14, dev)
    L2 = ppo.DeviceLearner(cfg, policy, value, B, 14, dev)        # takes the same step through the one-call product path
    assert L.h.num_params == sum(p.numel() for m in (policy, value) for p in m.parameters())
    assert torch.equal(L.params, _flat(policy, value))
    # make the behaviour log-prob consistent with the current policy (ratios near 1, some clipped)
    with torch.no_grad():
        lp, _ = ppo.torch_policy_logprob(policy, (batch["obs_p"][:-1] - norm["pm"]) / norm["ps"], batch["raw"])
        batch["logp"] = (lp + 0.25 * torch.randn_like(lp)).contiguous()
    idx = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:B].to(dev)
    noise = torch.randn(T * B, 14, generator=torch.Generator().manual_seed(6)).to(dev)
    ro = ppo.rollout_struct(batch)
    nm = capi.OduckNormalizer()
    nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = norm["pm"].data_ptr(), norm["ps"].data_ptr(), norm["vm"].data_ptr(), norm["vs"].data_ptr()
    idx32 = idx.to(torch.int32).contiguous()
    dbg = capi.PPO_DEBUG_SIMT if simt else 0
    F, Ls, Bk, Ad = capi.PPO_STAGE_FORWARD, capi.PPO_STAGE_LOSS, capi.PPO_STAGE_BACKWARD, capi.PPO_STAGE_ADAM

    loss, pl, vl, en, logits, values, adv, vs = _twin_loss(cfg, policy, value, batch, norm, idx, noise.view(T, B, 14))
    loss.backward()

    # ---- forward
    L.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, F | dbg)
    torch.cuda.synchronize()
    lg = L.view("LOGITS").view(-1, 32)[:T * B, :28]
    vv = L.view("VALUES").view(-1, 32)[:(T + 1) * B, 0]
    e_lg = (lg - logits.detach().reshape(T * B, 28)).abs().max().item()
    e_v = (vv - values.detach().reshape(-1)).abs().max().item()
    print(f"forward: max|dlogits|={e_lg:.2e} max|dvalue|={e_v:.2e}")
    assert e_lg < 5e-5 and e_v < 5e-5
    # ---- loss
    L.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, Ls | dbg)
    torch.cuda.synchronize()
    o = L.losses.tolist()
    e_adv = (L.view("ADV").view(T, B) - adv).abs().max().item()
    e_vs = (L.view("VS").view(T, B) - vs).abs().max().item()
    print(f"loss: device {o[:4]} torch {[loss.item(), pl.item(), vl.item(), en.item()]} max|dadv|={e_adv:.2e} max|dvs|={e_vs:.2e}")
    assert e_adv < 2e-4 and e_vs < 1e-4
    for a, b in zip(o[:4], (loss, pl, vl, en)):
        assert abs(a - b.item()) < 3e-5 * max(1.0, abs(b.item()))
    # ---- backward
    L.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, Bk | dbg)
    torch.cuda.synchronize()
    gref = _flat(policy, value, grad=True)
    gdev = L.grads.clone()
    worst = 0.0
    for net, l, which, off, r, c in _segments(L):
        a, b = gdev[off:off + r * c], gref[off:off + r * c]
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        worst = max(worst, rel)
        print(f"grad net={net} layer={l} {'b' if which else 'W'}: |ref|={b.norm().item():.3e} rel err={rel:.2e}")
        assert rel < (5e-5 if simt else 5e-4), (net, l, which, rel)
    # ---- clip + Adam, two steps (bias correction / step counter)
    params_ref = _flat(policy, value).clone()
    m1, m2 = torch.zeros_like(params_ref), torch.zeros_like(params_ref)
    for step in (1, 2):
        if step == 2:                                           # the second step re-runs the whole pipeline with the updated weights
            L.store_to(policy, value)
            for m in (policy, value):
                m.zero_grad()
            loss2, *_ = _twin_loss(cfg, policy, value, batch, norm, idx, noise.view(T, B, 14))
            loss2.backward()
            gref = _flat(policy, value, grad=True)
            params_ref = _flat(policy, value).clone()
            L.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, F | Ls | Bk | dbg)
        gn = gref.norm()
        g = gref * (cfg.max_grad_norm / gn) if gn >= cfg.max_grad_norm else gref            # optax.clip_by_global_norm
        m1 = 0.9 * m1 + 0.1 * g
        m2 = 0.999 * m2 + 0.001 * g * g
        upd = cfg.learning_rate * (m1 / (1 - 0.9 ** step)) / (torch.sqrt(m2 / (1 - 0.999 ** step)) + 1e-8)
        L.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, Ad | dbg)
        torch.cuda.synchronize()
        assert int(L.step.item()) == step
        d = (L.params - (params_ref - upd)).abs()
        err, frac = d.max().item(), (d > 2e-6).float().mean().item()
        print(f"adam step {step}: |g|={gn.item():.3e} max|dparam|={err:.2e} fraction>2e-6: {frac:.2e}")
        # Adam's first steps are sign-like: an element whose gradient is ~1e-6 of the typical size amplifies the GEMM rounding
        # error, so a handful of the 0.5 M elements may move differently (by at most two lr: opposite signs); everything else agrees to 2e-6
        assert frac < 3e-4 and err < 2.2 * cfg.learning_rate
        if step == 1:
            # ODUCK_PPO_ALL in one call (streams forked/joined inside, fused cooperative reduce + Adam) == the staged calls
            L2.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, capi.PPO_ALL | dbg)
            torch.cuda.synchronize()
            d2 = (L2.params - L.params).abs().max().item()
            print(f"one-call path vs staged path: max|dparam|={d2:.2e}, step={int(L2.step.item())}")
            assert d2 < 1e-7 and int(L2.step.item()) == 1


@pytest.mark.gpu
def test_in_kernel_entropy_noise_is_standard_normal_and_keyed():
    dev = torch.device("cuda:0")
    cfg, policy, value, batch, norm = _make(4, 256, 20, 1, dev)
    L = ppo.DeviceLearner(cfg, policy, value, 256, 14, dev)
    ro = ppo.rollout_struct(batch)
    nm = capi.OduckNormalizer()
    nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = norm["pm"].data_ptr(), norm["ps"].data_ptr(), norm["vm"].data_ptr(), norm["vs"].data_ptr()
    idx = torch.arange(256, dtype=torch.int32, device=dev)
    ents = []
    for k in (1, 1, 2):
        key = torch.tensor([0, k], dtype=torch.int32, device=dev)
        L.minibatch(ro, nm, idx.data_ptr(), 0, key.data_ptr(), capi.PPO_STAGE_FORWARD | capi.PPO_STAGE_LOSS)
        torch.cuda.synchronize()
        ents.append(L.losses.tolist()[3])
    # same key -> same draw (the f64 loss sums are atomics: equal up to summation order), new key -> new draw
    assert abs(ents[0] - ents[1]) < 1e-9 and abs(ents[0] - ents[2]) > 1e-6
    # E[entropy] over fresh normals (torch) agrees with the in-kernel draw within Monte-Carlo error
    with torch.no_grad():
        _, ent = ppo.torch_policy_logprob(policy, (batch["obs_p"][:-1] - norm["pm"]) / norm["ps"], batch["raw"])
    assert abs(ent.mean().item() - ents[0]) < 0.15


@pytest.mark.gpu
def test_trainer_device_learner_tracks_torch_learner():
    """Same seeds, one training step each: the device learner and the torch twin see the same rollout and the same
    minibatch permutation; only the entropy-sample noise differs (threefry vs torch.randn), which enters with weight 0.005."""
    from open_duck_playground_b200.joystick import Joystick
    res = {}
    for kind in ("torch", "device"):
        env = Joystick("flat_terrain_backlash", device="cuda:0")
        cfg = ppo.PPOConfig(num_envs=512, unroll_length=5, num_minibatches=4, num_updates_per_batch=2, learner=kind, cuda_graph=False)
        tr = ppo.PPOTrainer(env, cfg)
        m = tr.training_step()
        p = tr.params()
        res[kind] = (m, torch.cat([v.flatten() for v in p["policy"].values()]), torch.cat([v.flatten() for v in p["value"].values()]))
        assert math.isfinite(m["loss"])
    mt, pt, vt = res["torch"]
    md, pd_, vd = res["device"]
    print("torch", mt, "device", md)
    assert abs(mt["v_loss"] - md["v_loss"]) < 0.05 * max(1e-3, abs(mt["v_loss"]))
    # 8 Adam steps of lr 3e-4 move a weight by at most ~2.4e-3; on average the two learners must agree far inside that
    # (sign-like first steps amplify the entropy-noise difference on individual near-zero-gradient elements)
    assert (pt - pd_).abs().mean().item() < 1e-4 and (vt - vd).abs().mean().item() < 1e-4
    assert (pt - pd_).abs().max().item() < 2.5e-3 and (vt - vd).abs().max().item() < 2.5e-3


@pytest.mark.gpu
def test_graphed_rollout_equals_eager_rollout():
    """The CUDA-graph training step (per-step rollout graphs over persistent buffers, the learner's packed weights shared with the
    actor, one graph per update epoch with the input prefetch inside) replays the same launches as the eager loop: the very same
    parameters, bit for bit, after three training steps.  Both trainers use the two-kernel reduce / Adam tail (the one a graph
    captures): with the fused cooperative tail the eager path sums the global norm in another grouping, its clip factor differs
    in the last bit after step 1 (measured: weights 1.5e-8 apart, gradients identical), the policies then sample slightly
    different actions and the physics amplifies that to 1e-5 by step 3 -- chaos, not a defect (tools/dbg_graph_vs_eager.py)."""
    from open_duck_playground_b200.joystick import Joystick
    out = {}
    for graph in (False, True):
        env = Joystick("flat_terrain_backlash", device="cuda:0")
        cfg = ppo.PPOConfig(num_envs=256, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, learner="device", cuda_graph=graph, learner_fused_tail=False)
        tr = ppo.PPOTrainer(env, cfg)
        for _ in range(3):
            m = tr.training_step()
        torch.cuda.synchronize()
        out[graph] = (tr.dev_learner.params.clone(), m)
        assert math.isfinite(m["loss"])
    assert torch.equal(out[False][0], out[True][0])
    # and the fused tail moves the weights by no more than its last-bit clip factor in ONE step
    res = []
    for fused in (False, True):
        env = Joystick("flat_terrain_backlash", device="cuda:0")
        tr = ppo.PPOTrainer(env, ppo.PPOConfig(num_envs=256, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, learner="device", cuda_graph=False, learner_fused_tail=fused))
        tr.training_step()
        torch.cuda.synchronize()
        res.append(tr.dev_learner.params.clone())
    assert (res[0] - res[1]).abs().max().item() < 1e-6


@pytest.mark.gpu
def test_device_learner_checkpoint_roundtrip_and_standing_training():
    """params() / load() carry the device learner's weights and Adam state; the Standing task (85 / 153-wide observations) trains
    through the same learner."""
    from open_duck_playground_b200.joystick import Joystick
    from open_duck_playground_b200.standing import Standing
    cfg = ppo.PPOConfig(num_envs=256, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, learner="device")
    tr = ppo.PPOTrainer(Joystick("flat_terrain_backlash", device="cuda:0"), cfg)
    for _ in range(2):
        tr.training_step()
    ck = tr.params()
    tr2 = ppo.PPOTrainer(Joystick("flat_terrain_backlash", device="cuda:0"), cfg)
    tr2.load(ck)
    torch.cuda.synchronize()
    assert torch.equal(tr2.dev_learner.params, tr.dev_learner.params)
    assert torch.equal(tr2.dev_learner.view("ADAM_M"), tr.dev_learner.view("ADAM_M")) and torch.equal(tr2.dev_learner.view("ADAM_V"), tr.dev_learner.view("ADAM_V"))
    assert int(tr2.dev_learner.step.item()) == int(tr.dev_learner.step.item()) == 4 and tr2.env_steps == tr.env_steps
    # the restored actor acts like the original one on the same observations
    obs = tr.state.obs["state"].contiguous()
    for t_ in (tr, tr2):
        t_.weights.refresh(t_._mean32["state"], t_.stats["state"].std)
    a1, _, _ = ppo.policy_forward(tr.env, tr.weights, None, True, obs=obs)
    a2, _, _ = ppo.policy_forward(tr2.env, tr2.weights, None, True, obs=obs)
    assert torch.equal(a1, a2)
    ts = ppo.PPOTrainer(Standing("flat_terrain_backlash", device="cuda:0"), cfg)
    assert ts.policy.layers[0].in_features == 85 and ts.value.layers[0].in_features == 153
    for _ in range(3):
        m = ts.training_step()
    assert math.isfinite(m["loss"]) and math.isfinite(m["v_loss"]) and 0.0 <= m["clip_fraction"] <= 1.0


@pytest.mark.gpu
def test_torch_learner_graphed_update_equals_eager_update():
    """the torch learner's captured minibatch step (forward + backward graph, clip + Adam graph) must leave the same
    parameters AND the same Adam state as the eager loop -- the warm-up / capture-time optimiser steps are undone in place, the
    state tensors the graph updates are never replaced, and params() still carries the moments."""
    from open_duck_playground_b200.joystick import Joystick
    out = {}
    for graph in (False, True):
        env = Joystick("flat_terrain_backlash", device="cuda:0")
        cfg = ppo.PPOConfig(num_envs=256, unroll_length=4, num_minibatches=2, num_updates_per_batch=2, learner="torch", cuda_graph=graph, num_eval_envs=0,
                            entropy_cost=0.0)                          # no sampled entropy term: the two runs then differ by summation order only
        tr = ppo.PPOTrainer(env, cfg)
        for _ in range(2):
            m = tr.training_step()
        torch.cuda.synchronize()
        p = tr.params()
        st = p["optimizer"]["state"]
        assert len(st) == 16, "the optimiser state was dropped"
        out[graph] = (torch.cat([v.flatten() for v in p["policy"].values()] + [v.flatten() for v in p["value"].values()]),
                      torch.cat([st[k]["exp_avg"].flatten().cpu() for k in sorted(st)]), torch.cat([st[k]["exp_avg_sq"].flatten().cpu() for k in sorted(st)]),
                      [float(st[k]["step"]) for k in sorted(st)], m)
        assert math.isfinite(m["loss"])
    assert out[True][3] == out[False][3] == [8.0] * 16                  # 2 training steps x 2 epochs x 2 minibatches, no warm-up step left over
    # same rollout, same permutation, no noise term: fp32 summation order only (Adam's early sign-like steps amplify a flipped
    # near-zero gradient element to one step = 3e-4, hence the looser max)
    assert (out[True][0] - out[False][0]).abs().mean().item() < 2e-5 and (out[True][0] - out[False][0]).abs().max().item() < 2.5e-3
    assert (out[True][1] - out[False][1]).abs().max().item() < 5e-2 * max(1e-6, out[False][1].abs().max().item()) + 1e-4


@pytest.mark.gpu
def test_tf32_learner_mode_is_tf32_accurate_and_trains():
    """OduckPpoConfig.matmul_tf32 = 1 (PPOConfig.learner_matmul = "tf32"): one tensor-core pass on operands truncated to tf32 --
    XLA's default arithmetic for f32 dots on NVIDIA GPUs, i.e. the reference's.  Tolerances are tf32's (10-bit mantissa, errors
    accumulate over three hidden layers): forward 2e-2 abs on O(1) logits, gradients 3e-2 of the tensor norm (measured: up to 2.05e-2); the fp32-faithful
    default is held to 5e-5 / 5e-4 by test_learner_stages_match_torch."""
    N, T, nmb = 512, 20, 2
    dev = torch.device("cuda:0")
    cfg, policy, value, batch, norm = _make(3, N, T, nmb, dev)
    cfg.learner_matmul = "tf32"
    B = N // nmb
    L = ppo.DeviceLearner(cfg, policy, value, B, 14, dev)
    with torch.no_grad():
        lp, _ = ppo.torch_policy_logprob(policy, (batch["obs_p"][:-1] - norm["pm"]) / norm["ps"], batch["raw"])
        batch["logp"] = (lp + 0.25 * torch.randn_like(lp)).contiguous()
    idx = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:B].to(dev)
    noise = torch.randn(T * B, 14, generator=torch.Generator().manual_seed(6)).to(dev)
    ro = ppo.rollout_struct(batch)
    nm = capi.OduckNormalizer()
    nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = norm["pm"].data_ptr(), norm["ps"].data_ptr(), norm["vm"].data_ptr(), norm["vs"].data_ptr()
    idx32 = idx.to(torch.int32).contiguous()
    loss, pl, vl, en, logits, values, adv, vs = _twin_loss(cfg, policy, value, batch, norm, idx, noise.view(T, B, 14))
    loss.backward()
    L.minibatch(ro, nm, idx32.data_ptr(), noise.data_ptr(), 0, capi.PPO_STAGE_FORWARD | capi.PPO_STAGE_LOSS | capi.PPO_STAGE_BACKWARD)
    torch.cuda.synchronize()
    lg = L.view("LOGITS").view(-1, 32)[:T * B, :28]
    e_lg = (lg - logits.detach().reshape(T * B, 28)).abs().max().item()
    assert 1e-6 < e_lg < 2e-2, e_lg                                  # really the one-pass mode (not bit-faithful), and tf32-accurate
    gref, gdev = _flat(policy, value, grad=True), L.grads.clone()
    for net, l, which, off, r, c in _segments(L):
        a, b = gdev[off:off + r * c], gref[off:off + r * c]
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        print(f"tf32 grad net={net} layer={l} {'b' if which else 'W'}: rel err={rel:.2e}")
        assert rel < 3e-2, (net, l, which, rel)
    # and a trainer in this mode takes finite steps that track the fp32-faithful learner's
    from open_duck_playground_b200.joystick import Joystick
    res = {}
    for mode in ("fp32", "tf32"):
        tr = ppo.PPOTrainer(Joystick("flat_terrain_backlash", device="cuda:0"),
                            ppo.PPOConfig(num_envs=256, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, learner="device", learner_matmul=mode, num_eval_envs=0))
        m = tr.training_step()
        assert math.isfinite(m["loss"])
        res[mode] = tr.dev_learner.params.clone()
    assert (res["fp32"] - res["tf32"]).abs().mean().item() < 1e-4


@pytest.mark.gpu
def test_learner_reads_the_gathered_rank_major_blocks_in_place():
    """the all-gather leaves world blocks of rank-local [T, n, ...] buffers.  The learner indexes them in place
    (OduckRollout.block_envs / block_stride); gradients must equal, bit for bit, those from the time-major [T, world * n, ...]
    copy of the same data (same minibatch, same arithmetic -- only the gather addresses differ)."""
    dev = torch.device("cuda:0")
    T, n, world, nmb = 6, 96, 3, 2
    torch.manual_seed(11)
    bufs = [ppo.RolloutBuffers(T, n, 101, 212, 14, dev) for _ in range(world)]
    for b in bufs:
        b.flat.normal_()
        b["done"].copy_((torch.rand(T, n, device=dev) < 0.1).float()); b["trunc"].copy_(b["done"] * (torch.rand(T, n, device=dev) < 0.3).float())
        b["reward"].uniform_(0, 0.2); b["logp"].mul_(0.3).sub_(12.0)
    g = ppo.GatheredRollout(torch.cat([b.flat for b in bufs]), bufs[0], world)
    assert g["obs_p"].shape == (T + 1, world * n, 101) and torch.equal(g["obs_p"][:, n:2 * n], bufs[1]["obs_p"])
    tm = {k: g[k] for k in g.keys()}                                   # time-major copy
    N = world * n
    cfg = ppo.PPOConfig(num_envs=N, unroll_length=T, num_minibatches=nmb)
    policy, value = ppo.MLP([101, 512, 256, 128, 28]).to(dev), ppo.MLP([212, 512, 256, 128, 1]).to(dev)
    with torch.no_grad():
        policy.layers[-1].weight.mul_(0.05)
    B = N // nmb
    idx = torch.randperm(N, generator=torch.Generator().manual_seed(5))[:B].to(torch.int32).to(dev)
    assert len(set((idx.cpu() // n).tolist())) == world                # the minibatch draws from every block
    noise = torch.randn(T * B, 14, device=dev)
    nm = capi.OduckNormalizer()
    pm, ps, vm, vs_ = torch.zeros(101, device=dev), torch.ones(101, device=dev), torch.zeros(212, device=dev), torch.ones(212, device=dev)
    nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = pm.data_ptr(), ps.data_ptr(), vm.data_ptr(), vs_.data_ptr()
    grads = []
    for batch in (tm, g):
        L = ppo.DeviceLearner(cfg, policy, value, B, 14, dev)
        ro = ppo.rollout_struct(batch)
        assert (ro.block_envs, ro.block_stride) == ((n, bufs[0].flat.numel()) if batch is g else (0, 0))
        L.minibatch(ro, nm, idx.data_ptr(), noise.data_ptr(), 0, capi.PPO_STAGE_FORWARD | capi.PPO_STAGE_LOSS | capi.PPO_STAGE_BACKWARD)
        torch.cuda.synchronize()
        grads.append((L.grads.clone(), L.view("ADV").clone(), list(L.losses.tolist())))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    assert np.allclose(grads[0][2], grads[1][2], rtol=1e-12, atol=0)      # the reported losses are fp64 atomic sums: order-dependent in the last bits
    # The same with the policy observation left at home: obs["privileged_state"] starts with obs["state"] (joystick.py:596-615), so the
    # exchange may carry obs_v only and the learner reads its policy rows out of it (OduckRollout.obs_policy_ld = 212).
    for b in bufs:
        b["obs_v"][..., :101].copy_(b["obs_p"])
        b.policy_prefix = True
    assert bufs[0].skip == bufs[0]["obs_p"].numel()
    g2 = ppo.GatheredRollout(torch.cat([b.flat[b.skip:] for b in bufs]), bufs[0], world)
    assert g2.block_stride == bufs[0].flat.numel() - bufs[0].skip and torch.equal(g2["obs_p"], torch.cat([b["obs_p"] for b in bufs], dim=1))
    tm2 = {k: g2[k] for k in g2.keys()}
    grads2 = []
    for batch in (tm2, g2):
        L = ppo.DeviceLearner(cfg, policy, value, B, 14, dev)
        ro = ppo.rollout_struct(batch)
        assert ro.obs_policy_ld == (212 if batch is g2 else 0) and (batch is not g2 or ro.obs_policy == ro.obs_value)
        L.minibatch(ro, nm, idx.data_ptr(), noise.data_ptr(), 0, capi.PPO_STAGE_FORWARD | capi.PPO_STAGE_LOSS | capi.PPO_STAGE_BACKWARD)
        torch.cuda.synchronize()
        grads2.append(L.grads.clone())
    assert torch.equal(grads2[0], grads2[1]) and grads2[0].abs().max().item() > 0
    assert grads[0][0].abs().max().item() > 0


@pytest.mark.gpu
def test_pipelined_graphed_rollout_equals_plain_rollout():
    """PPOConfig.rollout_pipeline = 2: the rank's envs as two sub-batches with their own handles, graph chains and streams.  Every
    transition must equal the single-batch trainer's bit for bit -- eager and capture iterations, then graph replays (first
    hardware runs: xpassed in a child process, profiles/r02a..r02g_pytest_gpu.log)."""
    from open_duck_playground_b200.joystick import Joystick
    kw = dict(num_envs=512, unroll_length=5, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=0)
    a = ppo.PPOTrainer(Joystick("flat_terrain_backlash", device="cuda:0"), ppo.PPOConfig(**kw))
    b = ppo.PPOTrainer(Joystick("flat_terrain_backlash", device="cuda:0"), ppo.PPOConfig(rollout_pipeline=2, **kw))
    for it in range(4):
        ra, rb = a.rollout(), b.rollout()
        torch.cuda.synchronize()
        for k in ra:
            assert torch.equal(ra[k], rb[k]), (it, k)


@pytest.mark.gpu
def test_prefetched_minibatch_inputs_change_nothing():
    """oduck_ppo_prefetch packs the next minibatch's observation operands into the alternate input buffers beside the current
    minibatch's kernels; a run of SGD steps with it must leave the very same parameters as the same run without it (the pack
    kernel, its inputs and the GEMMs are identical -- only the buffer set and the stream differ)."""
    N, T, nmb = 512, 20, 4
    dev = torch.device("cuda:0")
    cfg, policy, value, batch, norm = _make(3, N, T, nmb, dev)
    B = N // nmb
    with torch.no_grad():
        lp, _ = ppo.torch_policy_logprob(policy, (batch["obs_p"][:-1] - norm["pm"]) / norm["ps"], batch["raw"])
        batch["logp"] = (lp + 0.25 * torch.randn_like(lp)).contiguous()
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(5)).to(torch.int32).to(dev)
    key = torch.tensor([7, 9], dtype=torch.int32, device=dev)
    ro = ppo.rollout_struct(batch)
    nm = capi.OduckNormalizer()
    nm.policy_mean, nm.policy_std, nm.value_mean, nm.value_std = norm["pm"].data_ptr(), norm["ps"].data_ptr(), norm["vm"].data_ptr(), norm["vs"].data_ptr()
    res = []
    for prefetch in (False, True):
        L = ppo.DeviceLearner(cfg, policy, value, B, 14, dev)
        for epoch in range(2):
            for i in range(nmb):
                L.minibatch(ro, nm, perm.data_ptr() + 4 * i * B, 0, key.data_ptr(), capi.PPO_ALL | capi.PPO_NO_COOP)
                if prefetch and i + 1 < nmb:
                    L.prefetch(ro, nm, perm.data_ptr() + 4 * (i + 1) * B)
        torch.cuda.synchronize()
        res.append((L.params.clone(), L.grads.clone()))
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][0], res[1][0])
    assert (res[0][0] - ppo.DeviceLearner(cfg, policy, value, B, 14, dev).params).abs().max().item() > 1e-5       # the steps really moved the weights
