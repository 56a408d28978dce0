"""Sim-to-sim check of an exported policy: the headless counterpart of the reference's ``mujoco_infer.py``
(playground/open_duck_mini_v2/mujoco_infer.py:16-266 -- load an ONNX policy, step the duck in plain MuJoCo at 50 Hz, build the
observation by hand, apply ``default + 0.25 * action`` with the motor speed limit).

Here the env itself is the simulator (1 env by default, same physics kernels as training), so the observation is the env's own
``obs["state"]`` and the action path (delay buffer, speed limit, position servo) is ``env.step``; the policy is the ``.onnx`` file
written by ``export_onnx`` evaluated with NumPy (onnxruntime is not in this image).  MuJoCo's viewer / keyboard control are out
of scope; ``--command`` fixes the joystick command instead.

    python -m open_duck_playground_b200.infer -o policy.onnx [--task flat_terrain_backlash] [--standing] [--steps 500]
                                              [--command 0.1 0 0 0 0 0 0]
"""
from __future__ import annotations

import argparse
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import rng as jr
from .export_onnx import run_onnx


def run_policy(env, onnx_model_path: str, steps: int = 500, seed: int = 0, command: Optional[Sequence[float]] = None, num_envs: int = 1) -> Dict[str, float]:
    """Roll ``steps`` control steps of ``env`` under the ONNX policy; returns summary statistics (and leaves ``env`` advanced)."""
    st = env.reset(jr.split(jr.PRNGKey(seed), num_envs))
    cmd = None if command is None else torch.tensor(list(command), dtype=st.info["command"].dtype, device=st.info["command"].device)
    total, falls, height = 0.0, 0, []
    for _ in range(steps):
        if cmd is not None:
            st.info["command"][:] = cmd                                  # views of the library's info record: the next step reads it
        obs = st.obs["state"].float().cpu().numpy()
        act = np.concatenate([run_onnx(onnx_model_path, obs[i:i + 1]) for i in range(num_envs)])
        st = env.step(st, torch.from_numpy(act).to(env.device))
        total += float(st.reward.float().mean())
        falls += int(st.done.float().sum())
        height.append(float(st.data.qpos[:, 2].float().mean()))
    return {"steps": steps, "mean_step_reward": total / steps, "episode_ends": falls, "mean_base_height": float(np.mean(height)), "final_base_height": height[-1]}


def main(argv=None) -> Dict[str, float]:
    ap = argparse.ArgumentParser(description="headless sim-to-sim run of an exported ONNX policy")
    ap.add_argument("-o", "--onnx_model_path", type=str, required=True)
    ap.add_argument("--task", type=str, default="flat_terrain_backlash")
    ap.add_argument("--standing", action="store_true", default=False)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=str, default="cuda:0")
    ap.add_argument("--command", type=float, nargs=7, default=None, help="lin_vel_x lin_vel_y ang_vel_yaw neck_pitch head_pitch head_yaw head_roll")
    args = ap.parse_args(argv)
    if args.standing:
        from .standing import Standing as Env
    else:
        from .joystick import Joystick as Env
    env = Env(task=args.task, device=args.device)
    out = run_policy(env, args.onnx_model_path, args.steps, args.seed, args.command)
    print(out)
    return out


if __name__ == "__main__":
    main()
