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
    ap.add_argument("--tf32", action="store_true", help="TF32 matmuls in the update")
    ap.add_argument("--fused-adam", action="store_true")
    ap.add_argument("--cuda-graphs", action="store_true", help="capture the rollout and the minibatch update as CUDA graphs")
    ap.add_argument("--fused-rollout", action="store_true", help="evaluate the policy in the rollout with the fused tcgen05 kernel")
    ap.add_argument("--fused-update", action="store_true", help="loss head / bias+ELU / clip+Adam / rollout bookkeeping through the library's kernels")
    ap.add_argument("--yaml", action="store_true", help="hyper-parameters and network from cfg/train/<Task>PPO.yaml (separate towers, mixed precision, ...)")
    ap.add_argument("--self-collision", type=int, default=-1, help="override sim.physx.self_collision (1 / 0); default: what the task's create_actor asks for")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import isaacgymenv_b200
    from isaacgymenv_b200.distributed import rank_info
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    info = rank_info()
    multi = info.world_size > 1
    if multi:      # one process per GPU (torchrun); envs sharded by rank, gradients all-reduced over NCCL
        torch.cuda.set_device(info.local_rank)
        dist.init_process_group("nccl", device_id=torch.device(info.device))
    overrides = {"sim": {"physx": {"self_collision": bool(args.self_collision)}}} if args.self_collision >= 0 else None
    env = isaacgymenv_b200.make(seed=args.seed, task=args.task, num_envs=args.num_envs, sim_device=info.device, rl_device=info.device, headless=True,
                                multi_gpu=multi, overrides=overrides)
    cfg = PPOConfig()
    if args.task != "Anymal" and args.task != "Hound":
        cfg = PPOConfig(units=(512, 256, 128), minibatch_size=16384, entropy_coef=0.001)
    if args.task in ("Houndarm", "Manipulator"):      # cfg/train/HoundarmPPO.yaml / ManipulatorPPO.yaml:24,65-67
        cfg = PPOConfig(units=(256, 128, 64), horizon_length=32, minibatch_size=16384, mini_epochs=5)
    if args.task == "Cartpole":
        cfg = PPOConfig(units=(32, 32), horizon_length=16, minibatch_size=8192, mini_epochs=8)
    if args.yaml:
        from isaacgymenv_b200.train import load_train_config, ppo_config_from_train_cfg

        cfg = ppo_config_from_train_cfg(load_train_config(f"{args.task}PPO"))
    cfg.tf32, cfg.fused_adam = args.tf32, args.fused_adam
    ppo = PPO(env, cfg, multi_gpu=multi, seed=args.seed + info.rank, fused_rollout=args.fused_rollout, cuda_graphs=args.cuda_graphs,
              fused_update=args.fused_update and env.num_acts <= 24)
    log = ppo.train(max_epochs=args.epochs, log_every=10, verbose=info.rank == 0)
    if multi:
        dist.barrier()
    if info.rank != 0:
        if multi:
            from isaacgymenv_b200.distributed import shutdown

            shutdown(ppo)
        return
    out = {"task": args.task, "self_collision_override": args.self_collision, "fused_rollout": bool(args.fused_rollout), "cuda_graphs": bool(args.cuda_graphs), "fused_update": bool(args.fused_update), "yaml": bool(args.yaml),
           "separate": bool(cfg.separate), "mixed_precision": bool(cfg.mixed_precision), "num_envs_per_gpu": args.num_envs, "n_gpus": info.world_size, "env_steps_all_gpus": [s * info.world_size for s in log.env_steps], "epochs": log.epochs, "env_steps": log.env_steps, "mean_episode_reward": log.mean_episode_reward,
           "mean_episode_length": log.mean_episode_length, "wall_s": log.wall_s, "gpu": torch.cuda.get_device_name(0),
           "env_steps_per_sec_incl_learner": (log.env_steps[-1] * info.world_size / log.wall_s[-1]) if log.wall_s else None}
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh)
    print(json.dumps({k: (v[-1] if isinstance(v, list) and v else v) for k, v in out.items()}))
    if multi:
        from isaacgymenv_b200.distributed import shutdown

        shutdown(ppo)


if __name__ == "__main__":
    main()
