"""Runs the training loop for Open Duck Mini V2 -- same CLI as the reference (open_duck_mini_v2/runner.py:35-64,
common/runner.py:24-118): ``--output_dir --num_timesteps --env --task --restore_checkpoint_path``.

    python -m open_duck_playground_b200.runner --task flat_terrain_backlash --num_timesteps 300000000
    torchrun --nproc-per-node 8 -m open_duck_playground_b200.runner --task flat_terrain_backlash   # env shards + NCCL gather
"""
import argparse
import os
from datetime import datetime
from pathlib import Path

import torch
import torch.distributed as dist

from . import joystick, randomize, standing
from .ppo import PPOConfig, PPOTrainer


class BaseRunner:
    def __init__(self, args: argparse.Namespace) -> None:
        self.args = args
        self.output_dir = Path.cwd() / Path(args.output_dir)
        self.num_timesteps = args.num_timesteps
        self.restore_checkpoint_path = None
        self.env = self.eval_env = self.randomizer = None
        self.action_size = self.obs_size = None
        self.rank, self.world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        self.writer = None
        if self.rank == 0:
            os.makedirs(self.output_dir, exist_ok=True)
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=str(self.output_dir))
            except Exception:   # tensorboard is optional in this image
                self.writer = None

    def progress_callback(self, num_steps: int, metrics: dict) -> None:
        if self.writer:
            for k, v in metrics.items():
                self.writer.add_scalar(k, v, num_steps)
        print("-----------")
        print(f'STEP: {num_steps} reward: {metrics["eval/episode_reward"]} reward_std: {metrics["eval/episode_reward_std"]}')
        print("-----------")

    def policy_params_fn(self, current_step, make_policy, params):
        d = datetime.now().strftime("%Y_%m_%d_%H%M%S")
        path = f"{self.output_dir}/{d}_{current_step}.pt"
        print(f"Saving checkpoint (step: {current_step}): {path}")
        torch.save(params, path)
        # deployment format like common/runner.py:76-84: an ONNX policy next to every checkpoint (input obs[1, obs_size])
        from .export_onnx import brax_param_tree, export_onnx
        hidden = [int(v.shape[0]) for k, v in params["policy"].items() if k.endswith(".weight")][:-1]
        export_onnx(brax_param_tree(params, params.get("policy_obs_key", "state")), self.action_size, hidden, self.obs_size,
                    output_path=f"{self.output_dir}/{d}_{current_step}.onnx", obs_key=params.get("policy_obs_key", "state"))

    def train(self) -> None:
        cfg = PPOConfig(num_timesteps=self.num_timesteps)      # BerkeleyHumanoidJoystickFlatTerrain table (common/runner.py:87-89)
        if getattr(self.args, "num_envs", None):
            cfg.num_envs = self.args.num_envs
        print(f"PPO params: {cfg}")
        trainer = PPOTrainer(self.env, cfg, rank=self.rank, world=self.world, progress_fn=self.progress_callback if self.rank == 0 else None,
                             policy_params_fn=self.policy_params_fn if self.rank == 0 else None)
        if self.restore_checkpoint_path:
            trainer.load(torch.load(self.restore_checkpoint_path, weights_only=False))
        trainer.train()


class OpenDuckMiniV2Runner(BaseRunner):
    def __init__(self, args):
        super().__init__(args)
        available_envs = {"joystick": (joystick, joystick.Joystick), "standing": (standing, standing.Standing)}   # open_duck_mini_v2/runner.py:14-17
        if args.env not in available_envs:
            raise ValueError(f"Unknown env {args.env}")
        self.env_file = available_envs[args.env]
        self.env_config = self.env_file[0].default_config()
        local = int(os.environ.get("LOCAL_RANK", 0))
        self.env = self.env_file[1](task=args.task, device=f"cuda:{local}")
        self.eval_env = self.env
        self.randomizer = randomize.domain_randomize
        self.action_size = self.env.action_size
        self.obs_size = int(self.env.observation_size["state"][0])
        self.restore_checkpoint_path = args.restore_checkpoint_path
        print(f"Observation size: {self.obs_size}")


def main() -> None:
    parser = argparse.ArgumentParser(description="Open Duck Mini Runner Script")
    parser.add_argument("--output_dir", type=str, default="checkpoints", help="Where to save the checkpoints")
    parser.add_argument("--num_timesteps", type=int, default=150000000)
    parser.add_argument("--env", type=str, default="joystick", help="env")
    parser.add_argument("--task", type=str, default="flat_terrain", help="Task to run")
    parser.add_argument("--restore_checkpoint_path", type=str, default=None, help="Resume training from this checkpoint")
    parser.add_argument("--num_envs", type=int, default=None, help="override the PPO table's num_envs (8192)")
    args = parser.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    runner = OpenDuckMiniV2Runner(args)
    runner.train()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
