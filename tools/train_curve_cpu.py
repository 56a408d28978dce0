"""A short PPO training run that records the evaluation curve.  Default: the CPU CHECKER stack (oracle fp32 port of the env + the
PyTorch fp32 twin of the learner) -- evidence that the restated env / rewards / PPO semantics learn, on the arithmetic the CUDA
path is parity-tested against; not a product path (the product has no CPU fallback) and not a benchmark.  ``--device cuda:0``: the
product itself (CUDA env + device learner, the reference PPO table unchanged), e.g. under a wall-clock ``timeout``: the curve file
is rewritten at every evaluation.

    python tools/train_curve_cpu.py --num_envs 512 --num_timesteps 4000000 --out profiles/r02_cpu_learning_curve.json
    timeout 16 python tools/train_curve_cpu.py --device cuda:0 --num_envs 8192 --num_minibatches 32 --num_eval_envs 128 \
        --num_timesteps 60000000 --num_evals 14 --learner_matmul tf32 --out gpurun_out/r02an_gpu_learning_curve.json

Reference loop: common/runner.py:86-118 (Brax ppo.train with progress_fn).  The PPO table is the reference's except num_envs /
batch shape, which are scaled down to what host cores finish in minutes.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="flat_terrain_backlash")
    ap.add_argument("--num_envs", type=int, default=512)
    ap.add_argument("--num_minibatches", type=int, default=8)
    ap.add_argument("--num_timesteps", type=int, default=4_000_000)
    ap.add_argument("--num_evals", type=int, default=10)
    ap.add_argument("--num_eval_envs", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", default="cpu", help="cpu = the checker stack (oracle library); cuda:N = the product (liboduck_cuda.so, device learner)")
    ap.add_argument("--learner_matmul", default="fp32", choices=["fp32", "tf32"])
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_cpu_learning_curve.json"))
    ap.add_argument("--save_prefix", default=None, help="write <prefix>.pt (trainer.params()) and <prefix>.onnx (deterministic policy, "
                                                        "common/runner.py:68-84) at the end, then run the ONNX file headless (infer.run_policy) and record it")
    ap.add_argument("--infer_steps", type=int, default=1000)
    args = ap.parse_args()

    import torch
    from open_duck_playground_b200 import ppo
    from open_duck_playground_b200.joystick import Joystick

    on_gpu = args.device.startswith("cuda")
    if on_gpu:
        make_env = lambda: Joystick(args.task, device=args.device)                                  # noqa: E731  (the product: no oracle import on this path)
        cfg = ppo.PPOConfig(num_envs=args.num_envs, num_minibatches=args.num_minibatches, num_timesteps=args.num_timesteps, num_evals=args.num_evals,
                            num_eval_envs=args.num_eval_envs, seed=args.seed, learner_matmul=args.learner_matmul)
        stack = f"the product on {torch.cuda.get_device_name(torch.device(args.device))}: CUDA env kernels + device learner ({args.learner_matmul} GEMMs), CUDA graphs"
    else:
        from oracle import oracle_lib
        torch.set_num_threads(os.cpu_count())
        make_env = lambda: Joystick(args.task, library=oracle_lib.load(f32=True, native=True))      # noqa: E731
        cfg = ppo.PPOConfig(num_envs=args.num_envs, num_minibatches=args.num_minibatches, num_timesteps=args.num_timesteps, num_evals=args.num_evals,
                            num_eval_envs=args.num_eval_envs, seed=args.seed, learner="torch", cuda_graph=False)
        stack = "the CPU checker stack (oracle fp32 env + PyTorch twin learner)"
    env = make_env()
    curve, t0 = [], time.time()

    def progress(steps, metrics):
        row = {"env_steps": int(steps), "wall_s": round(time.time() - t0, 1),
               **{k: float(v) for k, v in metrics.items() if k.startswith(("eval/episode_reward", "eval/avg_episode_length", "training/", "time/"))}}
        curve.append(row)
        print(json.dumps(row), flush=True)
        with open(args.out, "w") as f:
            json.dump({"what": f"PPO on {stack}; eval = first-episode sums over {args.num_eval_envs} eval envs (Brax EvalWrapper semantics); "
                               "wall_s includes construction, CUDA-graph capture and the evaluations",
                       "command": " ".join(sys.argv), "config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.__dict__.items()}, "curve": curve}, f, indent=1)

    tr = ppo.PPOTrainer(env, cfg, progress_fn=progress)
    params = tr.train()
    if args.save_prefix:
        # checkpoint -> deployment format -> sim-to-sim, the reference's own validation chain (common/runner.py:68-84, mujoco_infer.py)
        from open_duck_playground_b200 import infer
        from open_duck_playground_b200.export_onnx import brax_param_tree, export_onnx
        torch.save(params, args.save_prefix + ".pt")
        hidden = list(cfg.policy_hidden_layer_sizes)
        export_onnx(brax_param_tree(params, "state"), env.action_size, hidden, int(env.observation_size["state"][0]), output_path=args.save_prefix + ".onnx")
        sim = {}
        for name, command in (("forward_0.1", [0.1, 0, 0, 0, 0, 0, 0]), ("stand", [0, 0, 0, 0, 0, 0, 0])):
            e2 = make_env()
            sim[name] = infer.run_policy(e2, args.save_prefix + ".onnx", steps=args.infer_steps, seed=3, command=command, num_envs=8)
            print("sim2sim", name, json.dumps(sim[name]), flush=True)
        d = json.load(open(args.out))
        d["onnx_sim2sim"] = {"what": f"{args.save_prefix}.onnx (deterministic policy of the last checkpoint) evaluated with NumPy, 8 envs x {args.infer_steps} control steps "
                                     "under a fixed command (infer.run_policy; episode_ends counts falls + the 1000-step truncations)", **sim}
        with open(args.out, "w") as f:
            json.dump(d, f, indent=1)


if __name__ == "__main__":
    main()
