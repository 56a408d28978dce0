#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the batched Open Duck joystick step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one rollout step over the whole batch = ``oduck_policy_forward`` (actor MLP + NormalTanh sampling) followed by one
``oduck_step`` launch: action delay / push / motor-target logic, 10 x (forward dynamics + contact solve + Euler), obs (101 + 212),
7 reward terms, episode + auto-reset bookkeeping.
Workload at N = 1: BASELINE.json configs[1] -- ``flat_terrain_backlash`` (the task the metric names), 4096 envs per GPU,
domain randomisation on, no PPO update.  Envs are independent, so ranks take disjoint env shards (weak scaling: per-GPU
work fixed) and there is no data-path collective in this config.

``value``  : inputs (actions) already resident in HBM, K steps timed back to back with CUDA events, max over ranks.
``e2e``    : the same step through the C-ABI with HOST buffers: pinned actions H2D, oduck_step, D2H of obs["state"],
             reward and done -- copies inside the timed region.
``roofline``: HBM roofline the metric asks for (algorithmic 3400 B / env-step) plus the fp32 fraction that
             actually binds.
``cpu_baseline`` / ``--impl reference``: the CPU oracle port (liboduck_oracle_f32.so, fp32, std::thread over envs) on this
             box's host cores; stand-in for the reference's mujoco.mj_step path, which cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK = "flat_terrain_backlash"
BYTES_PER_ENV_STEP = 3400          # 309 words read + 541 written
FLOP_PER_ENV_STEP = 1.0e6          # planning figure, SURVEY.md 8d (physics only)
L2_BYTES = 126e6
STATE_BYTES_PER_ENV = 4 * (128 + 144 + 224 + 256 + 101 + 212 + 16)   # records one step touches (csrc/oduck_device.cuh)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(self.samples)}


def cpu_port_rate(n_envs, steps, f32=True):
    """env-steps/s of the CPU oracle port on this box (all host threads)."""
    import torch
    from open_duck_playground_b200 import rng as jr
    from open_duck_playground_b200.joystick import Joystick
    from oracle import oracle_lib

    env = Joystick(TASK, library=oracle_lib.load(f32=f32))
    env.randomize(jr.split(jr.PRNGKey(2), n_envs))
    st = env.reset(jr.split(jr.PRNGKey(0), n_envs))
    rs = np.random.default_rng(1)
    acts = [torch.from_numpy(rs.uniform(-1, 1, (n_envs, 14)).astype(np.float32)) for _ in range(steps + 1)]
    env.step(st, acts[0])
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(st, acts[k + 1])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt / steps * 1e3


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    n = 512
    cores = os.cpu_count()
    rate, ms = cpu_port_rate(n, max(1, args.steps // 10) if args.steps > 20 else max(1, args.steps))
    line = {
        "impl": "reference", "metric": "env-steps/sec (batched physics+rollout)", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{TASK} joystick env.step, {n}-env bounded sample of the 4096-env config, CPU", "task": TASK, "envs": n},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n} envs x timed control steps, oracle/liboduck_oracle_f32.so (the reference's mujoco.mj_step / MJX cannot be installed: no wheel, no network)"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ppo(args, rank, world, dev):
    """BASELINE configs[2]: full PPO (8192 envs x unroll 20 per training step), timed end to end with the rollout / gather / update split."""
    import torch
    import torch.distributed as dist
    from open_duck_playground_b200 import ppo
    from open_duck_playground_b200.joystick import Joystick
    n_total = 8192 if args.envs_per_gpu == 4096 else args.envs_per_gpu * world
    cfg = ppo.PPOConfig(num_envs=n_total)
    tr = ppo.PPOTrainer(Joystick(TASK, device=dev), cfg, rank=rank, world=world)
    for _ in range(max(1, min(args.warmup, 3))):
        tr.training_step()
    steps = max(1, args.steps // 20)
    split = {"rollout_ms": 0.0, "gather_ms": 0.0, "update_ms": 0.0}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.training_step()
        for k in split:
            split[k] += tr.timing[k]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"metric": "env-steps/sec (full PPO: rollout + gather + update)", "value": steps * cfg.num_envs * cfg.unroll_length / dt, "unit": "env-steps/s",
                          "n_gpus": world, "steps": steps, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{TASK} full PPO, {cfg.num_envs} envs x unroll {cfg.unroll_length}, 4 epochs x 32 minibatches (BASELINE configs[2])"},
                          "split_ms_per_training_step": {k: v / steps for k, v in split.items()}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "ppo"], help="ppo = BASELINE configs[2]: full PPO training steps (rollout / gather / update split)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from open_duck_playground_b200 import rng as jr
    from open_duck_playground_b200.joystick import Joystick

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    if args.mode == "ppo":
        return run_ppo(args, rank, world, dev)
    # env sets rotated so that the working set exceeds L2 (timing rule: inputs larger than L2)
    n_sets = max(3, int(np.ceil(1.3 * L2_BYTES / (n * STATE_BYTES_PER_ENV))))
    # per-rank keys: split(seed, world*n) then sliced, so results do not depend on the GPU count
    all_dr = jr.split(jr.PRNGKey(2), world * n)
    envs, states = [], []
    for s in range(n_sets):
        e = Joystick(TASK, device=dev)
        sl = slice(rank * n, (rank + 1) * n)
        e.randomize(all_dr[sl])
        states.append(e.reset(jr.split(jr.PRNGKey(100 + s), world * n)[sl]))
        envs.append(e)
    # rollout step = actor-MLP forward (A15, random-init weights of the reference architecture 101-512-256-128-28) + env.step
    from open_duck_playground_b200 import ppo
    torch.manual_seed(0)
    policy = ppo.MLP([101, 512, 256, 128, 28]).to(dev)
    weights = ppo.PolicyWeights(policy, 101, dev)
    n_keys = 8
    keys = [torch.from_numpy(jr.split(jr.PRNGKey(1000 + k), world * n)[rank * n:(rank + 1) * n].view(np.int32).copy()).to(dev) for k in range(n_keys)]   # resident in HBM
    host_keys = [k.cpu().pin_memory() for k in keys]
    host_obs = torch.empty(n, 101).pin_memory()
    host_out = torch.empty(n, 14 + 3).pin_memory()
    dev_out = torch.empty(n, 14 + 3, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kstep_events = []                                                   # (start, end) around every k_step launch of the timed region

    def step_resident(k):
        e = envs[k % n_sets]
        act, raw, logp = ppo.policy_forward(e, weights, keys[k % n_keys], deterministic=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.step(None, act)
        b.record()
        kstep_events.append((a, b))

    def step_e2e(k):
        e = envs[k % n_sets]
        kd = host_keys[k % n_keys].to(dev, non_blocking=True)           # H2D: this step's sampling keys
        act, raw, logp = ppo.policy_forward(e, weights, kd, deterministic=False)
        st = e.step(None, act)
        dev_out[:, :14] = raw; dev_out[:, 14] = logp; dev_out[:, 15] = st.reward; dev_out[:, 16] = st.done
        host_obs.copy_(st.obs["state"], non_blocking=True)              # D2H: what a host-side learner stores per transition
        host_out.copy_(dev_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()                       # the host consumes the result every step

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(steps):
            fn(k)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for k in range(max(3, args.warmup)):
        step_resident(k)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sum(e.handle.launch_count() for e in envs)
    kstep_events.clear()
    ms_total = timed(step_resident, args.steps)
    ms_kstep = sum(a.elapsed_time(b) for a, b in kstep_events) / len(kstep_events)   # average k_step launch duration, on its stream
    launches = sum(e.handle.launch_count() for e in envs) - l0
    for k in range(3):
        step_e2e(k)
    e2e_steps = max(10, args.steps // 2)
    ms_e2e = timed(step_e2e, e2e_steps)
    sampler.stop_flag = True
    ms_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3)
    e2e_value = world * n * e2e_steps / (ms_e2e * 1e-3)
    if rank == 0:
        hbm, peak_kind, sm_max = _peaks()
        achieved = BYTES_PER_ENV_STEP * n / (ms_kstep * 1e-3) / 1e9
        clocks = sampler.summary()
        fp32_peak = 148 * 128 * 2 * (clocks.get("sm_mhz") or sm_max) * 1e6
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"k_step_{n}")
        line = {
            "metric": "env-steps/sec (batched physics+rollout)", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{TASK} joystick rollout step = actor-MLP forward + env.step (10 substeps + obs/reward/auto-reset), {n} envs per GPU, domain randomisation on, no PPO update (BASELINE configs[1])",
                       "task": TASK, "envs_per_gpu": n, "global_envs": world * n, "substeps_per_step": 10, "parallelism": f"env-shard x{world}",
                       "l2": f"{n_sets} env sets rotated, {n_sets * n * STATE_BYTES_PER_ENV / 1e6:.0f} MB working set > L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                         "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                         "note": "the path is fp32-latency/compute bound (~300-400 FLOP/B), so the HBM fraction is small by construction; see fp32_frac",
                         "fp32_frac": FLOP_PER_ENV_STEP * value / world / fp32_peak, "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
                         "kernel": "k_step", "kernel_ms": ms_kstep, "kernel_share_of_step": ms_kstep / ms_step,
                         "kernel_ms_note": "average k_step launch duration from CUDA events around every launch of the timed region; rollout step = policy kernels + k_step"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * 2 * 4, "d2h_bytes_per_step": n * (101 + 17) * 4, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "physics_substeps_per_s": value * 10,
        }
        if not args.no_cpu_baseline and world == 1:
            n_cpu = 512
            rate, _ = cpu_port_rate(n_cpu, 6)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{n_cpu} envs x 6 control steps of the same workload, oracle/liboduck_oracle_f32.so"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
