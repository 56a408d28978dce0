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
FLOP_PER_ENV_STEP = 9.39e5         # counted: op-counter build of the oracle (tools/count_flops.py): 938 666 flop per env-step of the backlash model
L2_BYTES = 126e6
STATE_BYTES_PER_ENV = 4 * (128 + 144 + 224 + 256 + 101 + 212 + 16)   # records one step touches (csrc/oduck_device.cuh)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in-process every 10 ms (an
    nvidia-smi subprocess takes longer than the timed region of a short run), nvidia-smi as the fallback.  Only samples whose
    host timestamp falls inside a window marked with ``window()`` count."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.samples, self.stop_flag, self.windows, self.source = index, uuid, [], False, [], None
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.source = pynvml, "nvml"
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown, n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        return (time.perf_counter(), sm, self.sm_max, [bool(r & b) for b in bits])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        return (time.perf_counter(), float(f[0]), float(f[1]), [x.lower().startswith("active") for x in f[2:6]])

    def run(self):
        while not self.stop_flag:
            try:
                self.samples.append(self._sample_nvml() if self.nvml else self._sample_smi())
                self.source = self.source or "nvidia-smi"
            except Exception:
                pass
            time.sleep(0.01 if self.nvml else 0.1)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def summary(self):
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)] if self.windows else self.samples
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sample inside the timed region"], "samples": 0, "source": self.source}
        sm = sorted(s[1] for s in inside)
        reasons = [n for k, n in enumerate(self.NAMES) if any(s[3][k] for s in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": inside[0][2], "reasons": reasons, "samples": len(inside), "source": self.source}


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


def run_physics(args, rank, world, dev):
    """SURVEY 8d config 2 read literally: ``mjx_env.step(model, data, ctrl, 10)`` alone -- ``oduck_physics_substeps(n = 10)`` with
    ctrl = home + 0.25 U(-1, 1) redrawn every control step, no env logic, no policy (algorithmic bytes: 1 096 B / env-step)."""
    import torch
    import torch.distributed as dist
    from open_duck_playground_b200 import rng as jr
    from open_duck_playground_b200.joystick import Joystick
    n = args.envs_per_gpu
    n_sets = max(3, int(np.ceil(1.3 * L2_BYTES / (n * 4 * (128 + 144 + 224)))))
    envs = []
    for s in range(n_sets):
        e = Joystick(TASK, device=dev)
        e.randomize(jr.split(jr.PRNGKey(2), world * n)[rank * n:(rank + 1) * n])
        e.reset(jr.split(jr.PRNGKey(100 + s), world * n)[rank * n:(rank + 1) * n])
        envs.append(e)
    home = torch.tensor(envs[0]._mj_model.key_ctrl[:14], dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    ctrls = [(home + 0.25 * (2 * torch.rand(n, 14, device=dev, generator=g) - 1)).contiguous() for _ in range(8)]
    def step(k):
        envs[k % n_sets].physics_substeps(ctrls[k % 8], 10)
    for k in range(max(3, args.warmup, n_sets)):
        step(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        step(k)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        hbm, kind, _ = _peaks()
        ach = 1096 * n / (ms / args.steps * 1e-3) / 1e9
        print(json.dumps({"metric": "env-steps/sec (physics only: 10 x mjx.step per env-step)", "value": world * n * args.steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world,
                          "steps": args.steps, "warmup": max(3, args.warmup, n_sets), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{TASK} oduck_physics_substeps(n=10), {n} envs per GPU, domain randomisation on (SURVEY 8d config 2)", "envs_per_gpu": n,
                                     "l2": f"{n_sets} env sets rotated"},
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "algorithmic_bytes_per_env_step": 1096,
                                       "kernel": "k_physics", "peak_source": kind},
                          "physics_substeps_per_s": world * n * args.steps * 10 / (ms * 1e-3), "gpu_launches": args.steps}))
    if world > 1:
        dist.destroy_process_group()


def run_rollout_pipelined(args, rank, world, dev):
    """Experiment for the round-2 A/B (--pipeline P, off by default): the rollout step of one env batch as P independent sub-batches,
    each with its own library handle, CUDA graph and stream.  Envs are independent and the actor only reads its own sub-batch's
    observations, so sub-batch 0 may start step k + 1 while the tail CTAs of sub-batch P - 1's step k still run.  At 4096 envs
    k_step is 512 CTAs over 296 resident slots = 1.73 waves, and a latency-bound wave costs the same full or not: one stream pays
    for two waves per step, P streams keep the slots full (the per-env rate of the 64 k-env line).  Every env still takes one
    actor forward + one env.step per step, in order; results per env are those of the single-stream run (same keys, same slices)."""
    import torch
    import torch.distributed as dist
    from open_duck_playground_b200 import ppo, rng as jr
    from open_duck_playground_b200.joystick import Joystick
    P, n = args.pipeline, args.envs_per_gpu
    if n % P:
        raise SystemExit("--envs-per-gpu must be a multiple of --pipeline")
    m = n // P
    n_sets = max(3, int(np.ceil(1.3 * L2_BYTES / (n * STATE_BYTES_PER_ENV))))
    all_dr = jr.split(jr.PRNGKey(2), world * n)
    envs = []                                                           # envs[set][sub-batch]
    for s_ in range(n_sets):
        rk = jr.split(jr.PRNGKey(100 + s_), world * n)
        row = []
        for q in range(P):
            sl = slice(rank * n + q * m, rank * n + (q + 1) * m)
            e = Joystick(TASK, device=dev)
            e.randomize(all_dr[sl])
            e.reset(rk[sl])
            row.append(e)
        envs.append(row)
    torch.manual_seed(0)
    policy = ppo.MLP([101, 512, 256, 128, 28]).to(dev)
    weights = ppo.PolicyWeights(policy, 101, dev)
    n_keys = 8
    keys = [torch.from_numpy(jr.split(jr.PRNGKey(1000 + k), world * n)[rank * n:(rank + 1) * n].view(np.int32).copy()).to(dev) for k in range(n_keys)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    key_static = [torch.empty(m, 2, dtype=torch.int32, device=dev) for _ in range(P)]
    for k in range(max(3, args.warmup, n_sets)):                        # eager warm-up of every handle (sizes its actor scratch)
        for q in range(P):
            e = envs[k % n_sets][q]
            act, _, _ = ppo.policy_forward(e, weights, keys[k % n_keys][q * m:(q + 1) * m].contiguous(), deterministic=False)
            e.step(None, act)
    torch.cuda.synchronize()
    graphs = []
    for s_ in range(n_sets):
        row = []
        for q in range(P):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[q]):
                act, raw, logp = ppo.policy_forward(envs[s_][q], weights, key_static[q], deterministic=False)
                envs[s_][q].step(None, act)
            row.append((g, act, raw, logp))
        graphs.append(row)
    torch.cuda.synchronize()

    def step(k):
        for q in range(P):
            with torch.cuda.stream(streams[q]):
                key_static[q].copy_(keys[k % n_keys][q * m:(q + 1) * m], non_blocking=True)
                graphs[k % n_sets][q][0].replay()

    def timed(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        main = torch.cuda.current_stream(dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(main)
        for st in streams:
            st.wait_event(t0)
        for k in range(steps):
            step(k)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        t1.record(main)
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    timed(n_sets)
    ms_total = timed(args.steps)
    if rank == 0:
        value = world * n * args.steps / (ms_total * 1e-3)
        print(json.dumps({"metric": "env-steps/sec (batched physics+rollout)", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                          "warmup": max(3, args.warmup, n_sets), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{TASK} joystick rollout step = actor-MLP forward + env.step, {n} envs per GPU as {P} sub-batches of {m} on {P} streams (experiment)",
                                     "task": TASK, "envs_per_gpu": n, "pipeline": P, "l2": f"{n_sets} env sets rotated",
                                     "launch": "one CUDA graph per (env set, sub-batch), each sub-batch replayed on its own stream"},
                          "gpu_launches": 6 * P * args.steps, "physics_substeps_per_s": value * 10}))
    if world > 1:
        dist.destroy_process_group()


def run_ppo(args, rank, world, dev):
    """BASELINE configs[2]: full PPO (8192 envs x unroll 20 per training step), timed end to end with the rollout / gather / update split."""
    import torch
    import torch.distributed as dist
    from open_duck_playground_b200 import ppo
    from open_duck_playground_b200.joystick import Joystick
    n_total = 8192 if args.envs_per_gpu == 4096 else args.envs_per_gpu * world
    cfg = ppo.PPOConfig(num_envs=n_total, rollout_pipeline=args.pipeline)
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
                          "config": {"workload": f"{TASK} full PPO, {cfg.num_envs} envs x unroll {cfg.unroll_length}, 4 epochs x 32 minibatches (BASELINE configs[2])", "rollout_pipeline": cfg.rollout_pipeline},
                          "split_ms_per_training_step": {k: v / steps for k, v in split.items()}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    global TASK
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--task", default=TASK, help="scene: flat_terrain_backlash (the metric's config), flat_terrain, rough_terrain_backlash (height-field floor, BASELINE configs[3])")
    ap.add_argument("--pipeline", type=int, default=1, help="experiment: split the env batch into P sub-batches, one CUDA graph chain and stream each (rollout and ppo modes)")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "ppo", "physics"],
                    help="ppo = BASELINE configs[2]: full PPO training steps (rollout / gather / update split); physics = oduck_physics_substeps(10) alone")
    args = ap.parse_args()
    TASK = args.task
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
        # NCCL prints its version banner to stdout when NCCL_DEBUG is set on the box: park fd 1 on stderr while the
        # communicator comes up (device_id => eager init), so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n = args.envs_per_gpu
    if args.mode == "ppo":
        return run_ppo(args, rank, world, dev)
    if args.mode == "physics":
        return run_physics(args, rank, world, dev)
    if args.pipeline > 1:
        return run_rollout_pipelined(args, rank, world, dev)
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
    # e2e staging: per step the host receives obs["state"] (what a host-side learner stores per transition) and
    # [raw action 14 | log-prob | reward | done]; two device staging slots + two pinned host slots so that the D2H of step k
    # (copy stream) overlaps the compute of step k+1 -- the host still consumes every step's result inside the timed region.
    host_out = [torch.empty(n, 101 + 17).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    host_sink = torch.zeros(())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kstep_events = []                                                   # (start, end) around every k_step launch of the eager pass

    def step_eager(k):
        e = envs[k % n_sets]
        act, raw, logp = ppo.policy_forward(e, weights, keys[k % n_keys], deterministic=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.step(None, act)
        b.record()
        kstep_events.append((a, b))

    # The rollout step (5 actor kernels + k_step) of every env set is captured once into a CUDA graph; a timed step is one
    # key copy + one graph replay, so the loop stays kernel-bound whatever the host does (8 ranks share the box's cores).
    key_static = torch.empty_like(keys[0])
    graphs = []

    def capture_graphs():
        for s_ in range(n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                act, raw, logp = ppo.policy_forward(envs[s_], weights, key_static, deterministic=False)
                envs[s_].step(None, act)
            graphs.append((g, act, raw, logp))

    def step_resident(k):
        key_static.copy_(keys[k % n_keys], non_blocking=True)
        graphs[k % n_sets][0].replay()

    def consume(slot):
        nonlocal host_sink
        ev_done[slot].synchronize()
        host_sink = host_sink + host_out[slot][0, 101 + 15]              # the host reads the delivered result (a reward)

    # e2e step: the device part (actor + env.step + packing of the outgoing record) is a CUDA graph per env set as well; around it
    # the step uploads this step's keys from pinned memory and downloads the record of the step on the copy stream
    stage_set = [torch.empty(n, 101 + 17, device=dev) for _ in range(n_sets)]
    graphs_e2e = []

    def capture_graphs_e2e():
        for s_ in range(n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                act, raw, logp = ppo.policy_forward(envs[s_], weights, key_static, deterministic=False)
                st = envs[s_].step(None, act)
                sg = stage_set[s_]
                sg[:, :101] = st.obs["state"]; sg[:, 101:115] = raw; sg[:, 115] = logp; sg[:, 116] = st.reward; sg[:, 117] = st.done
            graphs_e2e.append(g)

    def step_e2e(k):
        slot = k & 1
        main = torch.cuda.current_stream(dev)
        key_static.copy_(host_keys[k % n_keys], non_blocking=True)      # H2D: this step's sampling keys (pinned)
        graphs_e2e[k % n_sets].replay()
        ev_ready[slot].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_ready[slot])
            host_out[slot].copy_(stage_set[k % n_sets], non_blocking=True)   # D2H on the copy stream (pinned)
            ev_done[slot].record(copy_stream)
        if k > 0:
            consume(slot ^ 1)                                           # host waits for (and reads) step k-1 while step k runs

    def timed(fn, steps, after=None):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        t0.record()
        for k in range(steps):
            fn(k)
        if after:
            after(steps)
        t1.record()
        barrier()
        sampler.window(h0, time.perf_counter())
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        pass
    sampler = ClockSampler(local, uuid)
    for k in range(max(3, args.warmup, n_sets)):                        # every env set runs eagerly at least once before the capture
        step_eager(k)                                                   # (first use of a handle sizes its actor scratch: an allocation)
    torch.cuda.synchronize()
    capture_graphs()
    for k in range(n_sets):
        step_resident(k)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    # eager pass with events around every k_step launch (events cannot bracket a kernel inside a graph): the kernel's average
    # duration for the roofline, and the library's own launch count per step
    eager_steps = min(args.steps, 60)
    l0 = sum(e.handle.launch_count() for e in envs)
    kstep_events.clear()
    timed(step_eager, eager_steps)
    ms_kstep = sum(a.elapsed_time(b) for a, b in kstep_events) / len(kstep_events)   # average k_step launch duration, on its stream
    launches = (sum(e.handle.launch_count() for e in envs) - l0) * args.steps // eager_steps
    capture_graphs_e2e()
    for k in range(4):
        step_e2e(k)
    torch.cuda.synchronize()
    e2e_steps = max(10, args.steps // 2)
    ms_e2e = timed(step_e2e, e2e_steps, after=lambda steps: consume((steps - 1) & 1))   # the last step's result is consumed too
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
            traffic = json.load(open(tpath)).get(f"k_step_{n}" if TASK.startswith("flat") else f"k_step_hf_{n}")
        line = {
            "metric": "env-steps/sec (batched physics+rollout)", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{TASK} joystick rollout step = actor-MLP forward + env.step (10 substeps + obs/reward/auto-reset), {n} envs per GPU, domain randomisation on, no PPO update (BASELINE configs[{1 if TASK.startswith('flat') else 3}])",
                       "task": TASK, "envs_per_gpu": n, "global_envs": world * n, "substeps_per_step": 10, "parallelism": f"env-shard x{world}",
                       "l2": f"{n_sets} env sets rotated, {n_sets * n * STATE_BYTES_PER_ENV / 1e6:.0f} MB working set > L2",
                       "launch": "one CUDA graph per env set (actor kernels + k_step), replayed per step"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                         "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                         "note": "the path is fp32-latency/compute bound (~300-400 FLOP/B), so the HBM fraction is small by construction; see fp32_frac",
                         "fp32_frac": FLOP_PER_ENV_STEP * value / world / fp32_peak, "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
                         "kernel": "k_step", "kernel_ms": ms_kstep, "kernel_share_of_step": ms_kstep / ms_step,
                         "kernel_ms_note": "average k_step launch duration from CUDA events around every launch of an eager pass right after the timed region (the timed steps replay a CUDA graph of the same launches); rollout step = policy kernels + k_step"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * 2 * 4, "d2h_bytes_per_step": n * (101 + 17) * 4, "steps": e2e_steps,
                    "note": "host keys H2D + policy + env.step + D2H of obs/raw/logp/reward/done every step; the D2H of step k runs on a copy stream under step k+1 and the host reads step k-1's result before it issues step k+1"},
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
