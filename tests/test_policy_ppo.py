"""A15 (actor MLP + NormalTanh head) and A17 (unroll / GAE / gather) host logic, on the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from open_duck_playground_b200 import ppo, rng as jr
from open_duck_playground_b200.joystick import Joystick


@pytest.fixture()
def env(oracle):
    e = Joystick("flat_terrain_backlash", library=oracle)
    e.reset(jr.split(jr.PRNGKey(0), 8))
    return e


def _weights(env, seed=0):
    torch.manual_seed(seed)
    pol = ppo.MLP([101, 512, 256, 128, 28])
    w = ppo.PolicyWeights(pol, 101, env.device)
    w.refresh(torch.randn(101) * 0.1, torch.rand(101) + 0.5)
    return pol, w


def test_deterministic_policy_matches_torch(env):
    pol, w = _weights(env)
    act, raw, logp = ppo.policy_forward(env, w, None, deterministic=True)
    x = (env.buffer("OBS_STATE").float() - w.mean) / w.std
    loc = pol(x)[:, :14]
    assert torch.allclose(act, torch.tanh(loc), atol=2e-5) and torch.allclose(raw, loc, atol=5e-5)


def test_stochastic_policy_logprob_is_consistent(env):
    pol, w = _weights(env, 1)
    keys = torch.from_numpy(jr.split(jr.PRNGKey(5), 8).view(np.int32))
    act, raw, logp = ppo.policy_forward(env, w, keys, deterministic=False)
    act2, raw2, _ = ppo.policy_forward(env, w, keys, deterministic=False)
    assert torch.equal(raw, raw2)                                        # same keys, same noise
    x = (env.buffer("OBS_STATE").float() - w.mean) / w.std
    lp_ref, _ = ppo.torch_policy_logprob(pol, x, raw)
    assert torch.allclose(logp, lp_ref.detach(), atol=2e-3)
    assert torch.allclose(act, torch.tanh(raw), atol=1e-6)
    # the noise is a standard normal: z = (raw - loc) / scale over many draws
    out = pol(x); loc, sp = out[:, :14], out[:, 14:]
    z = ((raw - loc) / (torch.nn.functional.softplus(sp) + 0.001)).detach().flatten()
    assert abs(float(z.mean())) < 0.35 and 0.6 < float(z.std()) < 1.4


def test_gae_matches_reference_recursion():
    T, N = 6, 5
    g = torch.Generator().manual_seed(0)
    rew, val, boot = torch.rand(T, N, generator=g), torch.rand(T, N, generator=g), torch.rand(N, generator=g)
    term = (torch.rand(T, N, generator=g) < 0.2).float()
    trunc = torch.zeros(T, N); trunc[3, 1] = 1.0; term[3, 1] = 0.0
    vs, adv = ppo.compute_gae(trunc, term, rew, val, boot, 0.95, 0.97)
    # direct transcription of brax compute_gae
    mask = 1 - trunc
    v_tp1 = torch.cat([val[1:], boot[None]])
    deltas = (rew + 0.97 * (1 - term) * v_tp1 - val) * mask
    acc = torch.zeros(N); out = []
    for t in reversed(range(T)):
        acc = deltas[t] + 0.97 * (1 - term[t]) * mask[t] * 0.95 * acc
        out.append(acc)
    vs_ref = torch.stack(out[::-1]) + val
    assert torch.allclose(vs, vs_ref)
    vs_tp1 = torch.cat([vs_ref[1:], boot[None]])
    assert torch.allclose(adv, (rew + 0.97 * (1 - term) * vs_tp1 - val) * mask)


def test_shard_keys_do_not_depend_on_world_size():
    full = ppo.shard_keys(3, 1, 0, 16)
    parts = np.concatenate([ppo.shard_keys(3, 4, r, 4) for r in range(4)])
    assert np.array_equal(full, parts)


def test_training_step_runs_and_learns_signal(oracle):
    env = Joystick("flat_terrain_backlash", library=oracle)
    cfg = ppo.PPOConfig(num_envs=16, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, num_timesteps=16 * 4 * 2)
    tr = ppo.PPOTrainer(env, cfg)
    before = [p.clone() for p in tr.policy.parameters()]
    m = tr.training_step()
    assert math.isfinite(m["loss"]) and tr.env_steps == 64
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.policy.parameters()))
    assert float(tr.stats["state"].count) == 64
    p = tr.params()
    tr2 = ppo.PPOTrainer(Joystick("flat_terrain_backlash", library=oracle), cfg)
    tr2.load(p)
    assert all(torch.equal(a, b) for a, b in zip(tr.policy.state_dict().values(), tr2.policy.state_dict().values()))


def test_evaluator_reports_first_episode_sums(oracle):
    """Brax EvalWrapper semantics: per-env sums over the first episode only, averaged over the eval envs."""
    env = Joystick("flat_terrain_backlash", library=oracle)
    cfg = ppo.PPOConfig(num_envs=8, unroll_length=3, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=6, episode_length=15)
    tr = ppo.PPOTrainer(env, cfg)
    ev = tr.evaluate()
    assert {"eval/episode_reward", "eval/episode_reward_std", "eval/avg_episode_length", "eval/episode_reward/alive", "eval/episode_swing_peak"} <= set(ev)
    assert 1.0 <= ev["eval/avg_episode_length"] <= 15.0
    # alive pays 20 per active step, so its episode sum is 20 x the episode length (joystick.py:83)
    assert abs(ev["eval/episode_reward/alive"] - 20.0 * ev["eval/avg_episode_length"]) < 1e-3
    assert ev["eval/episode_reward"] >= 0.0 and math.isfinite(ev["eval/episode_reward_std"])
    # the evaluator owns its envs: the training envs' state is untouched
    before = tr.state.data.qpos.clone()
    tr.evaluate()
    assert torch.equal(before, tr.state.data.qpos)
    m = []
    tr.progress_fn = lambda steps, metrics: m.append(metrics)
    tr.cfg.num_timesteps = 8 * 3 * 2
    tr.train()
    assert m and "eval/episode_reward" in m[-1] and "training/loss" in m[-1] and "eval/avg_episode_length" in m[-1]


def test_pipelined_rollout_equals_plain_rollout(oracle):
    """PPOConfig.rollout_pipeline = P: the rank's envs as P sub-batches with their own handles.  Envs are independent
    and the keys are sliced, so transitions and the parameters after a training step equal the single-batch trainer's bit for bit."""
    kw = dict(num_envs=12, unroll_length=3, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=0)
    plain = ppo.PPOTrainer(Joystick("flat_terrain_backlash", library=oracle), ppo.PPOConfig(**kw))
    piped = ppo.PPOTrainer(Joystick("flat_terrain_backlash", library=oracle), ppo.PPOConfig(rollout_pipeline=3, **kw))
    for _ in range(2):
        a, b = plain.rollout(), piped.rollout()
        for k in a:
            assert torch.equal(a[k], b[k]), k
    torch.manual_seed(5); plain.training_step()                       # the torch learner draws its permutation / entropy noise from the global generator
    torch.manual_seed(5); piped.training_step()
    for pa, pb in zip(plain.policy.parameters(), piped.policy.parameters()):
        assert torch.equal(pa, pb)
    with pytest.raises(ValueError, match="multiple of rollout_pipeline"):
        ppo.PPOTrainer(Joystick("flat_terrain_backlash", library=oracle), ppo.PPOConfig(rollout_pipeline=5, **kw))
