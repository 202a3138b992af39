"""One PPO iteration of BASELINE config 5 (Anymal, 8192 envs) for an ncu launch list: which kernels the minibatch update spends its time in."""
import sys
sys.path.insert(0, ".")
import torch
import isaacgymenv_b200
from isaacgymenv_b200.learning.ppo import PPO
from isaacgymenv_b200.train import load_train_config, ppo_config_from_train_cfg

env = isaacgymenv_b200.make(seed=42, task="Anymal", num_envs=8192, sim_device="cuda:0", rl_device="cuda:0", headless=True)
cfg = ppo_config_from_train_cfg(load_train_config("AnymalPPO"))
cfg.tf32 = True
cfg.mini_epochs = 1
ppo = PPO(env, cfg, seed=42, fused_rollout=True, cuda_graphs=False, fused_update=True)
ppo.train(max_epochs=2, log_every=10 ** 9)
torch.cuda.synchronize()
print("done")
