"""Small driver for `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py [what ...]`: a handful of envs
through every kernel family of liboduck_cuda.so (the sanitizer slows kernels by 10-100x).  what: flat hf policy ppo (default: all)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from open_duck_playground_b200 import ppo, rng as jr
from open_duck_playground_b200.joystick import Joystick

what = sys.argv[1:] or ["flat", "hf", "policy", "ppo"]
n = int(os.environ.get("SAN_ENVS", 40))                                   # not a multiple of the 8-env CTA: the grid tail is covered
dev = "cuda:0"
if "flat" in what or "policy" in what:
    env = Joystick("flat_terrain_backlash", device=dev)
    env.randomize(jr.split(jr.PRNGKey(1), n))
    st = env.reset(jr.split(jr.PRNGKey(0), n))
    if "flat" in what:
        for k in range(2):
            st = env.step(st, torch.rand(n, 14, device=dev) * 2 - 1)
        env.physics_substeps(None, 2)
        torch.cuda.synchronize()
        print("flat: k_randomize, k_reset, k_step x2, k_physics ok", float(st.reward.mean()))
    if "policy" in what:
        torch.manual_seed(0)
        w = ppo.PolicyWeights(ppo.MLP([101, 512, 256, 128, 28]).to(dev), 101, torch.device(dev))
        keys = torch.from_numpy(jr.split(jr.PRNGKey(3), n).view(np.int32).copy()).to(dev)
        act, raw, logp = ppo.policy_forward(env, w, keys, deterministic=False)
        T = 2
        buf = {"obs_p": torch.zeros(T + 1, n, 101, device=dev), "obs_v": torch.zeros(T + 1, n, 212, device=dev), "raw": torch.zeros(T, n, 14, device=dev),
               "logp": torch.zeros(T, n, device=dev), "reward": torch.zeros(T, n, device=dev), "done": torch.zeros(T, n, device=dev), "trunc": torch.zeros(T, n, device=dev)}
        ppo.attach_rollout_sink(env, buf)
        for t in range(T):
            ppo.rollout_step(env, w, keys, t)
        torch.cuda.synchronize()
        print("policy: k_pack_weights, k_pack_obs, k_gemm_tc x3, k_dense_tc, k_sink_obs0, k_step -> sink ok", float(logp.mean()), float(buf["reward"].mean()))
if "hf" in what:
    env = Joystick("rough_terrain_backlash", device=dev)
    env.randomize(jr.split(jr.PRNGKey(1), n))
    st = env.reset(jr.split(jr.PRNGKey(0), n))
    st = env.step(st, torch.rand(n, 14, device=dev) * 2 - 1)
    torch.cuda.synchronize()
    print("hf: k_reset<HF>, k_step<HF> ok", float(st.reward.mean()))
if "ppo" in what:
    cfg = ppo.PPOConfig(num_envs=64, unroll_length=3, num_minibatches=2, num_updates_per_batch=1, learner="device", cuda_graph=False, num_eval_envs=0)
    tr = ppo.PPOTrainer(Joystick("flat_terrain_backlash", device=dev), cfg)
    m = tr.training_step()
    torch.cuda.synchronize()
    print("ppo: device learner minibatch x2 ok", m["loss"])
