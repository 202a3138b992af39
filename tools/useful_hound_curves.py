#!/usr/bin/env python
"""Config 4 (UsefulHound) as a simulation one can train on: (a) how high robots get under random leg actions with the arm actions zero /
random, (b) PPO curves with the reference's UsefulHoundPPO.yaml for both end-effector-state settings -- `refreshEefState: false` replicates
the reference (the OSC law's velocity damping sees a never-refreshed, all-zero end-effector row), `true` feeds it the live row.
    python tools/useful_hound_curves.py --epochs 300 --out gpurun_out/useful_hound_curves.json"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO
    from isaacgymenv_b200.train import load_train_config, ppo_config_from_train_cfg

    out = {"task": "UsefulHound", "num_envs": args.num_envs, "gpu": torch.cuda.get_device_name(0)}
    for refresh in (False, True):
        key = f"refreshEefState={str(refresh).lower()}"
        res = {}
        for arm in ("zero", "random"):
            env = isaacgymenv_b200.make(seed=7, task="UsefulHound", num_envs=args.num_envs, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                        overrides={"env": {"refreshEefState": refresh}})
            env.reset()
            g = torch.Generator(device="cuda").manual_seed(3)
            zmax, vmax = 0.0, 0.0
            for _ in range(300):
                a = 2 * torch.rand(args.num_envs, 18, device="cuda", generator=g) - 1
                if arm == "zero":
                    a[:, 12:] = 0
                env.step(a)
                zmax = max(zmax, float(env.root_states[:, 2].max()))
                vmax = max(vmax, float(env.root_states[:, 7:10].norm(dim=-1).max()))
            res[f"arm_actions_{arm}"] = {"max_root_z_m": zmax, "max_root_speed_mps": vmax, "steps": 300}
            del env
        env = isaacgymenv_b200.make(seed=42, task="UsefulHound", num_envs=args.num_envs, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                    overrides={"env": {"refreshEefState": refresh}})
        cfg = ppo_config_from_train_cfg(load_train_config("UsefulHoundPPO"))
        cfg.tf32 = True
        ppo = PPO(env, cfg, seed=42, fused_rollout=True, cuda_graphs=True, fused_update=True)
        log = ppo.train(max_epochs=args.epochs, log_every=10, verbose=False)
        res["ppo"] = {"epochs": log.epochs, "env_steps": log.env_steps, "mean_episode_reward": log.mean_episode_reward, "mean_episode_length": log.mean_episode_length,
                      "wall_s": log.wall_s, "separate_towers": bool(cfg.separate), "units": list(cfg.units)}
        out[key] = res
        print(key, json.dumps({k: v for k, v in res.items() if k != "ppo"}), "reward first/last", log.mean_episode_reward[:1], log.mean_episode_reward[-1:],
              "episode length last", log.mean_episode_length[-1:], flush=True)
        ppo.release_graphs()
        del ppo, env
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh)


if __name__ == "__main__":
    main()
