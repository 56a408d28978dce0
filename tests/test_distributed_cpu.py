"""N > 1 host logic on CPU: two gloo ranks own disjoint env shards of the oracle, gather the rollout once per training
step, and must reproduce the single-process rollout of the full batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from open_duck_playground_b200 import ppo
    from open_duck_playground_b200.joystick import Joystick
    from oracle import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    env = Joystick("flat_terrain_backlash", library=oracle_lib.load())
    cfg = ppo.PPOConfig(num_envs=n_total, unroll_length=3, num_minibatches=2, num_updates_per_batch=1)
    tr = ppo.PPOTrainer(env, cfg, rank=rank, world=world)
    batch = ppo.all_gather_rollout(tr.rollout(), world)
    m = tr.update(batch)
    if rank == 0:
        torch.save({"reward": batch["reward"], "raw": batch["raw"], "obs": batch["obs_p"], "loss": m["loss"],
                    "w": [p.detach().clone() for p in tr.policy.parameters()]}, out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    n = 8
    single, multi = str(tmp_path / "single.pt"), str(tmp_path / "multi.pt")
    _worker(0, 1, 0, n, single)
    mp.spawn(_worker, args=(2, _free_port(), n, multi), nprocs=2, join=True)
    a, b = torch.load(single), torch.load(multi)
    assert a["reward"].shape == b["reward"].shape == (3, n)
    assert torch.equal(a["reward"], b["reward"]) and torch.equal(a["raw"], b["raw"]) and torch.equal(a["obs"], b["obs"])
    assert abs(a["loss"] - b["loss"]) < 1e-6
    assert all(torch.allclose(x, y, atol=1e-7) for x, y in zip(a["w"], b["w"]))
