"""Minimal PPO learner for the hot-path tasks -- a stand-in for rl_games' ``a2c_continuous`` agent, which the reference
delegates training to (``train.py:200-218``; hyper-parameters ``cfg/train/AnymalPPO.yaml``; loss structure as stated in-tree by
``learning/common_agent.py:361-509``): shared-trunk actor-critic MLP (ELU), state-independent log-std, GAE(gamma, tau) with
value bootstrap on time-outs (``extras["time_outs"]``), clipped surrogate + clipped value loss + bound loss, advantage /
observation / value normalisation, adaptive-KL learning rate, gradient-norm clipping.

This is SURVEY.md 8(f) row 1 ("next"): it exists to show that the B200 environment step trains (reward curve on Anymal) and
to drive the multi-GPU path (``torch.distributed`` all-reduce of gradients, one process per GPU, no collective in the env
step).  The network maths runs through torch/cuBLAS here; a hand-written tcgen05 forward is future work.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn


class RunningMeanStd(nn.Module):
    def __init__(self, shape, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("mean", torch.zeros(shape, dtype=torch.float64))
        self.register_buffer("var", torch.ones(shape, dtype=torch.float64))
        self.register_buffer("count", torch.ones((), dtype=torch.float64))

    @torch.no_grad()
    def update(self, x):
        x = x.reshape(-1, *self.mean.shape).double()
        bm, bv, bc = x.mean(0), x.var(0, unbiased=False), x.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        self.mean += delta * bc / tot
        self.var.copy_((self.var * self.count + bv * bc + delta * delta * self.count * bc / tot) / tot)
        self.count.copy_(tot)

    def normalize(self, x, clip=5.0):
        return torch.clamp((x - self.mean.float()) / torch.sqrt(self.var.float() + self.eps), -clip, clip)

    def denormalize(self, x):
        return x * torch.sqrt(self.var.float() + self.eps) + self.mean.float()


class ActorCritic(nn.Module):
    def __init__(self, num_obs, num_actions, units=(256, 128, 64)):
        super().__init__()
        layers, last = [], num_obs
        for u in units:
            layers += [nn.Linear(last, u), nn.ELU()]
            last = u
        self.trunk = nn.Sequential(*layers)
        self.mu = nn.Linear(last, num_actions)
        self.value = nn.Linear(last, 1)
        self.log_std = nn.Parameter(torch.zeros(num_actions))      # sigma_init const 0 -> std 1, fixed_sigma (state independent)

    def forward(self, obs):
        h = self.trunk(obs)
        return self.mu(h), self.log_std.expand(obs.shape[0], -1), self.value(h).squeeze(-1)


def neglogp(x, mu, log_std):
    return 0.5 * (((x - mu) / log_std.exp()) ** 2).sum(-1) + log_std.sum(-1) + 0.5 * x.shape[-1] * 1.8378770664093453


@dataclass
class PPOConfig:
    horizon_length: int = 24
    minibatch_size: int = 32768
    mini_epochs: int = 5
    gamma: float = 0.99
    tau: float = 0.95
    e_clip: float = 0.2
    entropy_coef: float = 0.0
    learning_rate: float = 3e-4
    kl_threshold: float = 0.008
    grad_norm: float = 1.0
    critic_coef: float = 2.0
    bounds_loss_coef: float = 0.001
    units: tuple = (256, 128, 64)
    max_epochs: int = 1000


@dataclass
class TrainLog:
    epochs: List[int] = field(default_factory=list)
    env_steps: List[int] = field(default_factory=list)
    mean_episode_reward: List[float] = field(default_factory=list)
    mean_episode_length: List[float] = field(default_factory=list)
    wall_s: List[float] = field(default_factory=list)


class PPO:
    def __init__(self, env, cfg: PPOConfig = PPOConfig(), multi_gpu: bool = False, seed: int = 42):
        self.env, self.cfg, self.multi_gpu = env, cfg, multi_gpu
        self.device = env.rl_device
        torch.manual_seed(seed)
        self.model = ActorCritic(env.num_obs, env.num_acts, cfg.units).to(self.device)
        self.obs_rms = RunningMeanStd((env.num_obs,)).to(self.device)
        self.val_rms = RunningMeanStd(()).to(self.device)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=cfg.learning_rate, eps=1e-8)
        self.lr = cfg.learning_rate
        if multi_gpu:
            import torch.distributed as dist

            for p in self.model.parameters():
                dist.broadcast(p.data, 0)
        n = env.num_envs
        self.ep_rew = torch.zeros(n, device=self.device)
        self.ep_len = torch.zeros(n, device=self.device)
        self.done_rew: List[float] = []
        self.done_len: List[float] = []

    def _allreduce_grads(self):
        import torch.distributed as dist

        flat = torch.cat([p.grad.reshape(-1) for p in self.model.parameters() if p.grad is not None])
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
        off = 0
        for p in self.model.parameters():
            if p.grad is not None:
                k = p.grad.numel()
                p.grad.copy_(flat[off:off + k].view_as(p.grad))
                off += k

    def train(self, max_epochs=None, log_every=10, verbose=False) -> TrainLog:
        cfg, env = self.cfg, self.env
        T, N = cfg.horizon_length, env.num_envs
        log = TrainLog()
        obs = env.reset()["obs"].clone()
        t0 = time.time()
        steps = 0
        for epoch in range(max_epochs or cfg.max_epochs):
            b_obs = torch.zeros(T, N, env.num_obs, device=self.device)
            b_act = torch.zeros(T, N, env.num_acts, device=self.device)
            b_nlp, b_val, b_rew, b_done = (torch.zeros(T, N, device=self.device) for _ in range(4))
            b_mu = torch.zeros(T, N, env.num_acts, device=self.device)
            with torch.no_grad():
                for t in range(T):
                    self.obs_rms.update(obs)
                    nobs = self.obs_rms.normalize(obs)
                    mu, log_std, v = self.model(nobs)
                    act = mu + log_std.exp() * torch.randn_like(mu)
                    b_obs[t], b_act[t], b_mu[t], b_nlp[t] = nobs, act, mu, neglogp(act, mu, log_std)
                    b_val[t] = self.val_rms.denormalize(v)
                    o, rew, done, extras = env.step(torch.clamp(act, -1.0, 1.0))
                    obs = o["obs"].clone()
                    rew = rew.clone()
                    # value bootstrap on time-outs (rl_games value_bootstrap, docs/release_notes.md:67)
                    rew += cfg.gamma * b_val[t] * extras["time_outs"].float()
                    b_rew[t], b_done[t] = rew, done.float()
                    self.ep_rew += rew
                    self.ep_len += 1
                    ids = done.nonzero(as_tuple=False).flatten()
                    if len(ids) > 0:
                        self.done_rew += self.ep_rew[ids].tolist()
                        self.done_len += self.ep_len[ids].tolist()
                        self.ep_rew[ids] = 0
                        self.ep_len[ids] = 0
                steps += T * N
                _, _, v_last = self.model(self.obs_rms.normalize(obs))
                v_last = self.val_rms.denormalize(v_last)
                adv = torch.zeros(T, N, device=self.device)
                last = torch.zeros(N, device=self.device)
                for t in reversed(range(T)):
                    nv = v_last if t == T - 1 else b_val[t + 1]
                    nonterm = 1.0 - b_done[t]
                    delta = b_rew[t] + cfg.gamma * nv * nonterm - b_val[t]
                    last = delta + cfg.gamma * cfg.tau * nonterm * last
                    adv[t] = last
                ret = adv + b_val
                self.val_rms.update(ret)
                f_ret = (ret.reshape(-1) - self.val_rms.mean.float()) / torch.sqrt(self.val_rms.var.float() + 1e-5)
                f_val = (b_val.reshape(-1) - self.val_rms.mean.float()) / torch.sqrt(self.val_rms.var.float() + 1e-5)
                f_adv = adv.reshape(-1)
                f_adv = (f_adv - f_adv.mean()) / (f_adv.std() + 1e-8)
                f_obs, f_act, f_nlp, f_mu = b_obs.reshape(T * N, -1), b_act.reshape(T * N, -1), b_nlp.reshape(-1), b_mu.reshape(T * N, -1)
            mb = min(cfg.minibatch_size, T * N)
            for _ in range(cfg.mini_epochs):
                perm = torch.randperm(T * N, device=self.device)
                kls = []
                for s in range(0, T * N, mb):
                    idx = perm[s:s + mb]
                    mu, log_std, v = self.model(f_obs[idx])
                    nlp = neglogp(f_act[idx], mu, log_std)
                    ratio = torch.exp(f_nlp[idx] - nlp)
                    a = f_adv[idx]
                    a_loss = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - cfg.e_clip, 1.0 + cfg.e_clip)).mean()
                    v_clip = f_val[idx] + (v - f_val[idx]).clamp(-cfg.e_clip, cfg.e_clip)
                    c_loss = torch.max((v - f_ret[idx]) ** 2, (v_clip - f_ret[idx]) ** 2).mean()
                    b_loss = (torch.clamp(mu - 1.1, min=0.0) ** 2 + torch.clamp(-1.1 - mu, min=0.0) ** 2).sum(-1).mean()
                    entropy = (log_std + 0.5 + 0.9189385332046727).sum(-1).mean()
                    loss = a_loss + 0.5 * cfg.critic_coef * c_loss - cfg.entropy_coef * entropy + cfg.bounds_loss_coef * b_loss
                    self.opt.zero_grad(set_to_none=True)
                    loss.backward()
                    if self.multi_gpu:
                        self._allreduce_grads()
                    nn.utils.clip_grad_norm_(self.model.parameters(), cfg.grad_norm)
                    self.opt.step()
                    with torch.no_grad():
                        # KL between the old and the new diagonal Gaussians (same fixed sigma family)
                        kl = (((mu - f_mu[idx]) ** 2) / (2.0 * torch.exp(2.0 * log_std))).sum(-1).mean()
                        kls.append(kl)
                kl = torch.stack(kls).mean()
                if self.multi_gpu:
                    import torch.distributed as dist

                    dist.all_reduce(kl)
                    kl /= dist.get_world_size()
                kl = float(kl)
                if kl > 2.0 * cfg.kl_threshold:
                    self.lr = max(self.lr / 1.5, 1e-6)
                elif kl < 0.5 * cfg.kl_threshold:
                    self.lr = min(self.lr * 1.5, 1e-2)
                for g in self.opt.param_groups:
                    g["lr"] = self.lr
            if (epoch + 1) % log_every == 0 or epoch == 0:
                r = sum(self.done_rew[-2000:]) / max(len(self.done_rew[-2000:]), 1)
                l = sum(self.done_len[-2000:]) / max(len(self.done_len[-2000:]), 1)
                log.epochs.append(epoch + 1)
                log.env_steps.append(steps)
                log.mean_episode_reward.append(r)
                log.mean_episode_length.append(l)
                log.wall_s.append(time.time() - t0)
                if verbose:
                    print(f"epoch {epoch + 1:5d} env_steps {steps:10d} ep_rew {r:8.3f} ep_len {l:7.1f} lr {self.lr:.2e} wall {time.time() - t0:6.1f}s", flush=True)
                self.done_rew, self.done_len = self.done_rew[-4000:], self.done_len[-4000:]
        return log
