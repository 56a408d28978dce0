"""A short PPO training run on the CPU CHECKER stack (oracle fp32 port of the env + the PyTorch fp32 twin of the learner) that
records the evaluation curve: evidence that the restated env / rewards / PPO semantics learn, on the arithmetic the CUDA path
is parity-tested against.  Not a product path (the product has no CPU fallback) and not a benchmark.

    python tools/train_curve_cpu.py --num_envs 512 --num_timesteps 4000000 --out profiles/r02_cpu_learning_curve.json

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
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_cpu_learning_curve.json"))
    args = ap.parse_args()

    import torch
    from open_duck_playground_b200 import ppo
    from open_duck_playground_b200.joystick import Joystick
    from oracle import oracle_lib

    torch.set_num_threads(os.cpu_count())
    env = Joystick(args.task, library=oracle_lib.load(f32=True, native=True))
    cfg = ppo.PPOConfig(num_envs=args.num_envs, num_minibatches=args.num_minibatches, num_timesteps=args.num_timesteps, num_evals=args.num_evals,
                        num_eval_envs=args.num_eval_envs, seed=args.seed, learner="torch", cuda_graph=False)
    curve, t0 = [], time.time()

    def progress(steps, metrics):
        row = {"env_steps": int(steps), "wall_s": round(time.time() - t0, 1),
               **{k: float(v) for k, v in metrics.items() if k.startswith(("eval/episode_reward", "eval/avg_episode_length", "training/"))}}
        curve.append(row)
        print(json.dumps(row), flush=True)
        with open(args.out, "w") as f:
            json.dump({"what": "PPO on the CPU checker stack (oracle fp32 env + PyTorch twin learner); eval = first-episode sums over "
                               f"{args.num_eval_envs} eval envs (Brax EvalWrapper semantics)",
                       "command": " ".join(sys.argv), "config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.__dict__.items()}, "curve": curve}, f, indent=1)

    tr = ppo.PPOTrainer(env, cfg, progress_fn=progress)
    tr.train()


if __name__ == "__main__":
    main()
