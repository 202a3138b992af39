"""Command-line entry in the style of the reference's Hydra ``train.py`` (``isaacgymenvs/train.py:83-218``; options of
``cfg/config.yaml``):

    python -m isaacgymenv_b200.train task=Anymal headless=True
    python -m isaacgymenv_b200.train task=AnymalTerrain num_envs=4096 max_iterations=300 seed=7
    python -m isaacgymenv_b200.train task=Anymal test=True checkpoint=runs/Anymal/nn/Anymal.pth num_envs=64
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 -m isaacgymenv_b200.train task=Anymal multi_gpu=True

``key=value`` overrides as in Hydra; dotted keys reach into the task / train yaml (``task.env.learn.pushInterval_s=8``,
``train.params.config.horizon_length=16``).  Hyper-parameters come from ``cfg/train/<Task>PPO.yaml`` (the reference's files with
the interpolations resolved); the learner is the in-repo PPO (``learning/ppo.py``) instead of rl_games, which is not installed.
Extra switches of this repo: ``cuda_graphs=True`` (default), ``fused_update=True`` (loss head + clip + Adam kernels), ``fused_rollout=True`` (tcgen05 policy kernel: resident weights for [256,128,64], streamed weights for [512,256,128]).
Checkpoints go to ``runs/<experiment or task>/nn/<name>.pth`` (rl_games' directory layout, ``docs/rl_examples.md``; the file holds this
learner's own state dict -- network, normalisers, optimiser -- not rl_games' key names), every ``save_frequency`` epochs and at the end."""
from __future__ import annotations

import json
import os
import sys
from typing import Any, Dict

import yaml

DEFAULTS = {"task": "Anymal", "train": "", "experiment": "", "num_envs": "", "seed": 42, "max_iterations": "", "sim_device": "cuda:0",
            "rl_device": "cuda:0", "graphics_device_id": 0, "test": False, "checkpoint": "", "multi_gpu": False, "headless": True,
            "cuda_graphs": True, "fused_rollout": False, "fused_update": False, "tf32": True, "save_frequency": "", "output": ""}


def _parse_value(v: str):
    try:
        return yaml.safe_load(v)
    except yaml.YAMLError:
        return v


def parse_overrides(argv):
    top: Dict[str, Any] = dict(DEFAULTS)
    task_over: Dict[str, Any] = {}
    train_over: Dict[str, Any] = {}
    for a in argv:
        if "=" not in a:
            raise SystemExit(f"expected key=value, got {a!r}")
        k, v = a.split("=", 1)
        val = _parse_value(v)
        if k.startswith("task.") or k.startswith("train."):
            root, rest = k.split(".", 1)
            d = task_over if root == "task" else train_over
            parts = rest.split(".")
            for p in parts[:-1]:
                d = d.setdefault(p, {})
            d[parts[-1]] = val
        elif k in top:
            top[k] = val
        else:
            raise SystemExit(f"unknown option {k!r}; known: {sorted(top)} or task.* / train.*")
    return top, task_over, train_over


def _deep_update(d, u):
    for k, v in u.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            _deep_update(d[k], v)
        else:
            d[k] = v


def load_train_config(name: str, overrides=None) -> Dict[str, Any]:
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg", "train", f"{name}.yaml")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"no train config {path}")
    cfg = yaml.safe_load(open(path))
    if overrides:
        _deep_update(cfg, overrides)
    return cfg


def ppo_config_from_train_cfg(tc: Dict[str, Any], max_iterations=None):
    from .learning.ppo import PPOConfig

    c = tc["params"]["config"]
    units = tuple(tc["params"]["network"]["mlp"]["units"])
    separate = bool(tc["params"]["network"].get("separate", False))
    shaper = c.get("reward_shaper") or {}
    return PPOConfig(mixed_precision=bool(c.get("mixed_precision", False)), separate=separate, reward_scale=float(shaper.get("scale_value", 1.0)), save_frequency=int(c.get("save_frequency", 0) or 0),
                     horizon_length=int(c["horizon_length"]), minibatch_size=int(c["minibatch_size"]), mini_epochs=int(c["mini_epochs"]),
                     gamma=float(c["gamma"]), tau=float(c["tau"]), e_clip=float(c["e_clip"]), entropy_coef=float(c.get("entropy_coef", 0.0)),
                     learning_rate=float(c["learning_rate"]), kl_threshold=float(c.get("kl_threshold", 0.008)), grad_norm=float(c.get("grad_norm", 1.0)),
                     critic_coef=float(c.get("critic_coef", 2.0)), bounds_loss_coef=float(c.get("bounds_loss_coef", 0.0) or 0.0), units=units,
                     max_epochs=int(max_iterations or c["max_epochs"]))


def main(argv=None):
    top, task_over, train_over = parse_overrides(sys.argv[1:] if argv is None else argv)
    import torch
    import torch.distributed as dist

    import isaacgymenv_b200
    from .distributed import rank_info
    from .learning.ppo import PPO

    info = rank_info()
    multi = bool(top["multi_gpu"]) and info.world_size > 1
    sim_device, rl_device = top["sim_device"], top["rl_device"]
    if multi:       # utils/rlgames_utils.py:89-107: one process per GPU, device = local rank
        sim_device = rl_device = info.device
        torch.cuda.set_device(info.local_rank)
        dist.init_process_group("nccl", device_id=torch.device(info.device))
    task = top["task"]
    tc = load_train_config(top["train"] or f"{task}PPO", train_over)
    num_envs = int(top["num_envs"]) if top["num_envs"] != "" else None
    env = isaacgymenv_b200.make(seed=int(top["seed"]), task=task, num_envs=num_envs, sim_device=sim_device, rl_device=rl_device,
                                graphics_device_id=int(top["graphics_device_id"]), headless=bool(top["headless"]), multi_gpu=multi,
                                overrides=task_over or None)
    cfg = ppo_config_from_train_cfg(tc, top["max_iterations"] if top["max_iterations"] != "" else None)
    cfg.tf32 = bool(top["tf32"])
    fused = bool(top["fused_rollout"]) and len(cfg.units) == 3 and cfg.units[0] <= 512 and max(cfg.units[1:]) <= 256
    # a task whose step() synchronises with the host (generic hook path: reset_buf.nonzero(); domain randomisation: frame-count
    # schedules and host-side generators) cannot be replayed from a CUDA graph -- it would freeze at its warm-up values
    graphs = bool(top["cuda_graphs"]) and not getattr(env, "needs_host_sync", False)
    if top["save_frequency"] != "":
        cfg.save_frequency = int(top["save_frequency"])
    ppo = PPO(env, cfg, multi_gpu=multi, seed=int(top["seed"]) + info.rank, fused_rollout=fused, cuda_graphs=graphs,
              fused_update=bool(top["fused_update"]) and env.num_acts <= 24)
    name = top["experiment"] or tc["params"]["config"]["name"]
    ckpt = top["output"] or os.path.join("runs", name, "nn", f"{name}.pth")
    ppo.checkpoint_path = ckpt if info.rank == 0 else None
    if top["checkpoint"]:
        ppo.load(top["checkpoint"], load_optimizer=not top["test"])
    if top["test"]:
        rew, length = ppo.play(steps=int(env.max_episode_length) + 10 if hasattr(env, "max_episode_length") else 1000)
        if info.rank == 0:
            print(json.dumps({"task": task, "mode": "test", "mean_episode_reward": rew, "mean_episode_length": length}))
    else:
        log = ppo.train(max_epochs=cfg.max_epochs, log_every=10, verbose=info.rank == 0)
        if info.rank == 0:
            ppo.save(ckpt)
            print(json.dumps({"task": task, "mode": "train", "epochs": log.epochs[-1] if log.epochs else 0, "env_steps": log.env_steps[-1] if log.env_steps else 0,
                              "mean_episode_reward": log.mean_episode_reward[-1] if log.mean_episode_reward else None,
                              "wall_s": log.wall_s[-1] if log.wall_s else None, "checkpoint": ckpt}))
    if multi:
        from .distributed import shutdown

        shutdown(ppo)


if __name__ == "__main__":
    main()
