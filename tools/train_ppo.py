#!/usr/bin/env python
"""Train PPO on a hot-path task with the B200 env step and dump the learning curve as JSON.
    python tools/train_ppo.py --task Anymal --num-envs 4096 --epochs 300 --out gpurun_out/ppo_anymal.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="Anymal")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    env = isaacgymenv_b200.make(seed=args.seed, task=args.task, num_envs=args.num_envs, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    cfg = PPOConfig()
    if args.task != "Anymal" and args.task != "Hound":
        cfg = PPOConfig(units=(512, 256, 128), minibatch_size=16384, entropy_coef=0.001)
    if args.task == "Cartpole":
        cfg = PPOConfig(units=(32, 32), horizon_length=16, minibatch_size=8192, mini_epochs=8)
    ppo = PPO(env, cfg, seed=args.seed)
    log = ppo.train(max_epochs=args.epochs, log_every=10, verbose=True)
    out = {"task": args.task, "num_envs": args.num_envs, "epochs": log.epochs, "env_steps": log.env_steps, "mean_episode_reward": log.mean_episode_reward,
           "mean_episode_length": log.mean_episode_length, "wall_s": log.wall_s, "gpu": torch.cuda.get_device_name(0)}
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh)
    print(json.dumps({k: (v[-1] if isinstance(v, list) and v else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
