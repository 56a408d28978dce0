"""Debug aid: where do the eager and the graphed PPO paths part?  Rollout buffers and parameters after every training step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_duck_playground_b200 import ppo
from open_duck_playground_b200.joystick import Joystick

out = {}
for graph in (False, True):
    env = Joystick("flat_terrain_backlash", device="cuda:0")
    cfg = ppo.PPOConfig(num_envs=256, unroll_length=4, num_minibatches=2, num_updates_per_batch=1, learner="device", cuda_graph=graph)
    tr = ppo.PPOTrainer(env, cfg)
    rec = []
    for it in range(3):
        local = tr.rollout()
        torch.cuda.synchronize()
        snap = {k: v.clone() for k, v in local.items()}
        tr.update(local, sharded=False)
        torch.cuda.synchronize()
        rec.append((snap, tr.dev_learner.params.clone(), tr.dev_learner.grads.clone()))
    out[graph] = rec
for it in range(3):
    a, b = out[False][it], out[True][it]
    print(f"step {it}: rollout fields equal:", {k: bool(torch.equal(a[0][k], b[0][k])) for k in a[0]},
          "max|dparam| %.3e frac>1e-6 %.4f max|dgrad| %.3e" % ((a[1] - b[1]).abs().max().item(), ((a[1] - b[1]).abs() > 1e-6).float().mean().item(), (a[2] - b[2]).abs().max().item()))
