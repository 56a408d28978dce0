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


def test_train_loop_follows_brax_outer_loop(oracle):
    """Brax ppo.train: evaluation of the initial policy at step 0, then num_evals - 1 iterations of num_resets_per_eval epochs (each
    ended by a hard reset of every training env with fresh keys) + evaluation + progress_fn + policy_params_fn; an epoch is
    ceil(num_timesteps / (iterations x epochs x env-steps per training step)) training steps."""
    def run(resets, pipeline=1):
        cfg = ppo.PPOConfig(num_envs=8, unroll_length=8, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=2, episode_length=5, num_evals=3,
                            num_resets_per_eval=resets, num_timesteps=8 * 8 * 7, rollout_pipeline=pipeline)
        prog, saved, hard = [], [], []
        tr = ppo.PPOTrainer(Joystick("flat_terrain_backlash", library=oracle), cfg, progress_fn=lambda s, m: prog.append((s, set(m))),
                            policy_params_fn=lambda s, mk, p: saved.append((s, p["env_steps"])))
        orig = tr.reset_training_envs
        tr.reset_training_envs = lambda: (hard.append(tr.env_steps), orig())[1]
        tr.train()
        return tr, prog, saved, hard
    # 2 iterations x 2 epochs x ceil(448 / (2 x 2 x 64)) = 2 training steps each: 8 training steps = 512 >= 448 env-steps
    tr, prog, saved, hard = run(2)
    assert [s for s, _ in prog] == [0, 256, 512] and tr.env_steps == 512
    assert not any(k.startswith("training/") for k in prog[0][1]) and "eval/episode_reward" in prog[0][1]      # initial policy: evaluation only
    assert {"training/loss", "time/rollout_ms", "eval/avg_episode_length"} <= prog[1][1]
    assert saved == [(256, 256), (512, 512)]
    assert hard == [128, 256, 384, 512]                                    # after every epoch
    st = tr.state
    assert int(st.info["step"].max()) == 0 and float(st.done.max()) == 0.0    # the training envs are freshly reset at the end
    # num_resets_per_eval = 0: one epoch per evaluation, no hard resets (Brax's default)
    tr0, prog0, _, hard0 = run(0)
    assert [s for s, _ in prog0] == [0, 256, 512] and hard0 == [] and int(tr0.state.info["step"].max()) > 0
    # the sub-batch pipeline resets every sub-env, with the same keys as the single batch
    trp, _, _, hardp = run(2, pipeline=2)
    assert hardp == hard
    q = torch.cat([s_.data.qpos for s_ in trp.state])
    assert torch.equal(q, tr.state.data.qpos)


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


def test_refresh_rewrites_weights_in_place_and_invalidates_every_handle(oracle):
    """the library caches the repacked weights per handle keyed on the w[0] address; refresh() must keep the address
    (persistent buffers) and make EVERY env that uses the weights drop its cache, however rarely it runs a forward."""
    a = Joystick("flat_terrain_backlash", library=oracle); a.reset(jr.split(jr.PRNGKey(0), 4))
    b = a.spawn(); b.reset(jr.split(jr.PRNGKey(0), 4))
    pol, w = _weights(a)
    ptrs = [t.data_ptr() for t in w.w + w.b]
    calls = []
    for e, name in ((a, "a"), (b, "b")):
        orig = e.handle.policy_invalidate
        e.handle.policy_invalidate = (lambda o=orig, n=name: (calls.append(n), o())[1])
    ppo.policy_forward(a, w, None, deterministic=True); ppo.policy_forward(b, w, None, deterministic=True)
    assert calls == ["a", "b"]
    ppo.policy_forward(a, w, None, deterministic=True)
    assert calls == ["a", "b"]                                           # unchanged weights: the cache stays
    with torch.no_grad():
        for lin in pol.layers:
            lin.weight.mul_(0.5); lin.bias.add_(0.1)
    w.refresh(w.mean, w.std)
    assert [t.data_ptr() for t in w.w + w.b] == ptrs                     # same buffers, new contents
    assert torch.equal(w.w[0], pol.layers[0].weight.detach().t()) and torch.equal(w.b[3], pol.layers[3].bias.detach())
    act_b, _, _ = ppo.policy_forward(b, w, None, deterministic=True)     # b first this time: it must not run stale weights
    assert calls == ["a", "b", "b"]
    x = (b.buffer("OBS_STATE").float() - w.mean) / w.std
    assert torch.allclose(act_b, torch.tanh(pol(x)[:, :14]), atol=2e-5)


def test_spawn_keeps_model_config_and_autoreset(oracle):
    e = Joystick("flat_terrain_backlash", library=oracle, auto_reset=False, config_overrides={"noise_config.level": 0.0})
    e.reset(jr.split(jr.PRNGKey(0), 3))
    s = e.spawn()
    assert s._mj_model is e._mj_model and s._xml_path == e._xml_path and s._auto_reset is False and s._lib is e._lib
    assert s._config.noise_config.level == 0.0 and s._config is not e._config and s._handle is None
    s.reset(jr.split(jr.PRNGKey(0), 5))
    assert s.num_envs == 5 and e.num_envs == 3
    from open_duck_playground_b200.standing import Standing
    st = Standing("flat_terrain_backlash", library=oracle)
    assert type(st.spawn()) is Standing


def test_policy_obs_key_other_than_state_is_passed_explicitly(oracle):
    """with policy_obs_key='privileged_state' the policy is 212 wide; the rollout must hand those observations to the
    library instead of letting it read 212 floats out of the 101-wide obs['state'] records."""
    env = Joystick("flat_terrain_backlash", library=oracle)
    cfg = ppo.PPOConfig(num_envs=8, unroll_length=2, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=0, policy_obs_key="privileged_state", learner="torch")
    tr = ppo.PPOTrainer(env, cfg)
    assert tr.policy.layers[0].in_features == 212
    buf = tr.rollout()
    x = (buf["obs_p"][0] - tr.stats["privileged_state"].mean32) / tr.stats["privileged_state"].std
    lp, _ = ppo.torch_policy_logprob(tr.policy, x, buf["raw"][0])
    assert torch.allclose(lp.detach(), buf["logp"][0], atol=2e-3)
    with pytest.raises(ValueError, match="pass obs="):
        ppo.policy_forward(env, tr.weights, None, deterministic=True)


@pytest.mark.parametrize("P", [1, 2])
def test_kernel_side_rollout_writes_equal_the_copy_path(oracle, P):
    """A17: with a rollout sink attached the library stores every Transition itself (oduck_rollout_step); the buffers must equal
    the ones the round-1 path filled with seven copies per step -- over two unrolls (slot 0 of the second comes from the handle's
    observations), single batch and sub-batches with env offsets."""
    kw = dict(num_envs=12, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, num_eval_envs=0, learner="torch", rollout_pipeline=P,
              episode_length=6)
    outs = []
    for sink in (False, True):
        env = Joystick("flat_terrain_backlash", library=oracle, config_overrides={"episode_length": 6})   # auto-reset inside the second unroll
        tr = ppo.PPOTrainer(env, ppo.PPOConfig(kernel_rollout_writes=sink, **kw))
        assert tr._use_sink is sink
        outs.append([{k: v.clone() for k, v in tr.rollout().items()} for _ in range(2)])
    for a, b in zip(*outs):
        for k in a:
            assert torch.equal(a[k], b[k]), k
    assert outs[1][1]["done"].sum() > 0 and outs[1][1]["trunc"].sum() > 0          # the truncation / auto-reset rows were exercised


def test_rollout_sink_argument_checks(oracle):
    env = Joystick("flat_terrain_backlash", library=oracle)
    env.reset(jr.split(jr.PRNGKey(0), 4))
    _, w = _weights(env)
    keys = torch.from_numpy(jr.split(jr.PRNGKey(5), 4).view(np.int32))
    with pytest.raises(Exception, match="no sink attached"):
        ppo.rollout_step(env, w, keys, 0)
    T = 3
    mk = lambda n, dp=101: {"obs_p": torch.zeros(T + 1, n, dp), "obs_v": torch.zeros(T + 1, n, 212), "raw": torch.zeros(T, n, 14), "logp": torch.zeros(T, n),   # noqa: E731
                            "reward": torch.zeros(T, n), "done": torch.zeros(T, n), "trunc": torch.zeros(T, n)}
    with pytest.raises(Exception, match="do not fit"):
        ppo.attach_rollout_sink(env, mk(6), env_offset=3)
    with pytest.raises(Exception, match="row widths"):
        ppo.attach_rollout_sink(env, mk(4, dp=85))
    buf = mk(6)
    ppo.attach_rollout_sink(env, buf, env_offset=2)
    with pytest.raises(Exception, match="outside the unroll"):
        ppo.rollout_step(env, w, keys, T)
    st = ppo.rollout_step(env, w, keys, 1)
    assert torch.equal(buf["obs_p"][2, 2:6], st.obs["state"].float()) and torch.equal(buf["reward"][1, 2:6], st.reward.float())
    assert (buf["obs_p"][2, :2] == 0).all() and (buf["obs_p"][0] == 0).all()       # other envs' rows / slot 0 untouched at t = 1
    ppo.attach_rollout_sink(env, None)
    with pytest.raises(Exception, match="no sink attached"):
        ppo.rollout_step(env, w, keys, 0)


def test_running_statistics_one_pass_update_equals_the_literal_brax_form():
    """RunningStats.update takes its two sums out of one fused pass (torch.var_mean); the result must be brax
    running_statistics.update's -- count += n, mean += sum(x - mean_old) / count, summed_variance += sum((x - mean_old)(x - mean_new)),
    std = sqrt(summed_variance / count) -- here evaluated literally in float64, over several batches with offsets up to 100 sigma."""
    from open_duck_playground_b200.ppo import RunningStats
    torch.manual_seed(0)
    scale = torch.tensor([1, 2, 0.1, 5, 1, 1, 30.0])
    shift = torch.tensor([0, 3, -2, 100, 0, 1e-3, -50.0])
    rs = RunningStats(7, "cpu")
    cnt, mean, m2 = 0.0, torch.zeros(7, dtype=torch.float64), torch.zeros(7, dtype=torch.float64)
    for it in range(5):
        x = torch.randn(5, 600, 7) * scale + shift + 0.1 * it
        rs.update(x)
        xd = x.double().reshape(-1, 7)
        d_old = xd - mean
        cnt += xd.shape[0]
        mean = mean + d_old.sum(0) / cnt
        m2 = m2 + (d_old * (xd - mean)).sum(0)
    std = torch.sqrt(m2 / cnt)
    assert float(rs.count) == cnt
    assert ((rs.mean.double() - mean).abs() / (std + mean.abs())).max().item() < 1e-6
    assert ((rs.std.double() - std) / std).abs().max().item() < 2e-5


def test_gathered_rollout_without_the_policy_observation():
    """Host logic of the reduced exchange: when obs["privileged_state"] starts with obs["state"] the rank's flat buffer
    travels from obs_v on (RolloutBuffers.skip), and the gathered blocks still answer for every field -- obs_p as the leading columns
    of obs_v -- with the pointers / strides the device learner needs (OduckRollout.block_envs, block_stride, obs_policy_ld)."""
    from open_duck_playground_b200 import capi
    from open_duck_playground_b200.ppo import GatheredRollout, RolloutBuffers, rollout_struct
    T, n, world = 3, 5, 2
    bufs = [RolloutBuffers(T, n, 101, 212, 14, "cpu") for _ in range(world)]
    g0 = torch.Generator().manual_seed(3)
    for b in bufs:
        b.flat.copy_(torch.randn(b.flat.shape, generator=g0))
        b["obs_v"][..., :101].copy_(b["obs_p"])
        b.policy_prefix = True
    assert bufs[0].skip == (T + 1) * n * 101 and bufs[0].offsets["obs_v"][0] == bufs[0].skip
    g = GatheredRollout(torch.cat([b.flat[b.skip:] for b in bufs]), bufs[0], world)        # what all_gather_into_tensor leaves behind
    assert g.block_stride == bufs[0].flat.numel() - bufs[0].skip
    for k in bufs[0].keys():
        assert torch.equal(g[k], torch.cat([b[k] for b in bufs], dim=1)), k
    ro = rollout_struct(g)
    assert (ro.num_envs, ro.unroll, ro.block_envs, ro.block_stride, ro.obs_policy_ld) == (world * n, T, n, g.block_stride, 212)
    assert ro.obs_policy == ro.obs_value == g.flat.data_ptr() and ro.raw_action == g.flat.data_ptr() + 4 * (bufs[0].offsets["raw"][0] - bufs[0].skip)
    # without the prefix property everything travels and the learner gets two dense observation tensors
    for b in bufs:
        b.policy_prefix = False
    g = GatheredRollout(torch.cat([b.flat for b in bufs]), bufs[0], world)
    ro = rollout_struct(g)
    assert ro.obs_policy_ld == 0 and ro.obs_policy == g.flat.data_ptr() and ro.obs_value == g.flat.data_ptr() + 4 * bufs[0].offsets["obs_v"][0]
    assert torch.equal(g["obs_p"], torch.cat([b["obs_p"] for b in bufs], dim=1))
