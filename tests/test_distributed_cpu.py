"""N > 1 host logic on CPU: two gloo ranks own disjoint env shards of the oracle.  Replicated mode: they gather the rollout once per
training step and must reproduce the single-process rollout and update of the full batch.  Sharded mode (Brax's
pmean scheme): averaged gradients and merged normaliser moments keep the ranks in step."""
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


def _sharded_worker(rank, world, port, n_total, out_prefix):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from open_duck_playground_b200 import ppo
    from open_duck_playground_b200.joystick import Joystick
    from oracle import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    env = Joystick("flat_terrain_backlash", library=oracle_lib.load())
    cfg = ppo.PPOConfig(num_envs=n_total, unroll_length=8, num_minibatches=2, num_updates_per_batch=2, update_mode="sharded", num_eval_envs=0)
    tr = ppo.PPOTrainer(env, cfg, rank=rank, world=world)
    ms = [tr.training_step() for _ in range(2)]
    torch.save({"w": [p.detach().clone() for p in tr.policy.parameters()] + [p.detach().clone() for p in tr.value.parameters()],
                "count": float(tr.stats["state"].count), "mean": tr.stats["state"].mean.clone(), "std": tr.stats["state"].std.clone(),
                "mode": tr.last_update_mode, "env_steps": tr.env_steps, "loss": [m["loss"] for m in ms]}, f"{out_prefix}{rank}.pt")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_sharded_update_keeps_the_ranks_in_step(tmp_path):
    """update_mode = "sharded" (Brax's pmean scheme, the CUDA default at N > 1) on two gloo ranks: every rank updates on its own env
    shard, gradients are averaged per minibatch and the observation-normaliser moments are summed over the ranks -- so both
    ranks hold the same weights and the same normaliser afterwards (its merge rule itself: test_running_stats_reduce_merges_the_shards)."""
    n = 8
    mp.spawn(_sharded_worker, args=(2, _free_port(), n, str(tmp_path / "r")), nprocs=2, join=True)
    _sharded_worker(0, 1, 0, n, str(tmp_path / "s"))
    r0, r1, s = (torch.load(str(tmp_path / f)) for f in ("r0.pt", "r1.pt", "s0.pt"))
    assert r0["mode"] == r1["mode"] == "sharded" and s["mode"] == "single"
    assert r0["env_steps"] == s["env_steps"] == 2 * n * 8
    assert all(np.isfinite(v) for v in r0["loss"] + r1["loss"])
    assert all(torch.equal(a, b) for a, b in zip(r0["w"], r1["w"]))              # same averaged gradients, same Adam: the ranks stay in step
    assert any(not torch.equal(a, b) for a, b in zip(r0["w"], s["w"]))           # (the shards' own minibatches: not the single-process update)
    # first training step: same transitions as the single process (keys sliced per rank), so the merged moments are the whole batch's;
    # after the first update the policies differ, so only the count is comparable from then on
    assert r0["count"] == r1["count"] == s["count"] == 2 * n * 8
    assert torch.equal(r0["mean"], r1["mean"]) and torch.equal(r0["std"], r1["std"])


def _stats_worker(rank, world, port, out_prefix):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from open_duck_playground_b200 import ppo
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(6, 10, 5, generator=g) * torch.tensor([1.0, 2.0, 0.5, 3.0, 1e-3]) + torch.tensor([0.0, 1.0, -2.0, 10.0, 0.0])
    st = ppo.RunningStats(5, torch.device("cpu"))
    for t in range(3):                                                       # three updates, each rank sees its half of the env axis
        st.update(full[2 * t:2 * t + 2, rank * 5:(rank + 1) * 5], reduce=True)
    torch.save({"mean": st.mean, "std": st.std, "count": float(st.count)}, f"{out_prefix}{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_running_stats_reduce_merges_the_shards(tmp_path):
    """RunningStats.update(reduce=True) (Brax: psum of the batch moments over the pmap axis) == the single-process statistics of the
    concatenated batch, update by update."""
    from open_duck_playground_b200 import ppo
    mp.spawn(_stats_worker, args=(2, _free_port(), str(tmp_path / "st")), nprocs=2, join=True)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(6, 10, 5, generator=g) * torch.tensor([1.0, 2.0, 0.5, 3.0, 1e-3]) + torch.tensor([0.0, 1.0, -2.0, 10.0, 0.0])
    ref = ppo.RunningStats(5, torch.device("cpu"))
    for t in range(3):
        ref.update(full[2 * t:2 * t + 2])
    a, b = torch.load(str(tmp_path / "st0.pt")), torch.load(str(tmp_path / "st1.pt"))
    assert a["count"] == b["count"] == float(ref.count) == 60.0
    assert torch.equal(a["mean"], b["mean"]) and torch.equal(a["std"], b["std"])
    assert torch.allclose(a["mean"], ref.mean, rtol=1e-5, atol=1e-6) and torch.allclose(a["std"], ref.std, rtol=1e-4, atol=1e-7)
    flat = full.reshape(-1, 5)
    assert torch.allclose(a["mean"], flat.mean(0), rtol=1e-5, atol=1e-6) and torch.allclose(a["std"], flat.std(0, unbiased=False), rtol=1e-4, atol=1e-7)
