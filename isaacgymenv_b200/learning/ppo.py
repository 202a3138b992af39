"""Minimal PPO learner for the hot-path tasks -- a stand-in for rl_games' ``a2c_continuous`` agent, which the reference
delegates training to (``train.py:200-218``; hyper-parameters ``cfg/train/AnymalPPO.yaml``; loss structure as stated in-tree by
``learning/common_agent.py:361-509``): actor-critic MLP (ELU; shared trunk or, with ``network.separate: true`` as in the rough-terrain
train configs, separate actor / critic towers), state-independent log-std, GAE(gamma, tau) with
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
    """rl_games' ``actor_critic`` network builder as the train yamls configure it (``cfg/train/AnymalPPO.yaml:5-22``): ``separate:
    false`` = one MLP trunk feeding both heads (flat tasks), ``separate: true`` = an actor tower and a critic tower of the same shape
    (``cfg/train/AnymalTerrainPPO.yaml:8``, ``UsefulHoundPPO.yaml:8``).  ``trunk`` is the actor tower (and the shared trunk)."""

    def __init__(self, num_obs, num_actions, units=(256, 128, 64), separate=False):
        super().__init__()

        def mlp():
            layers, last = [], num_obs
            for u in units:
                layers += [nn.Linear(last, u), nn.ELU()]
                last = u
            return nn.Sequential(*layers), last

        self.separate = bool(separate)
        self.fused_layers = False      # set by PPO(fused_update=True): hidden layers through learning/fused_update.linear_elu in training mode
        self.bf16_wgrad = False        # weight-gradient GEMMs on bf16 operands (fp32 accumulate / output): PPOConfig.mixed_precision
        self.direct_grad = False       # fused layers store parameter gradients into the existing .grad views (PPO fused update: one use per backward)
        self.trunk, last = mlp()
        self.critic_trunk = mlp()[0] if self.separate else None
        self.mu = nn.Linear(last, num_actions)
        self.value = nn.Linear(last, 1)
        self.log_std = nn.Parameter(torch.zeros(num_actions))      # sigma_init const 0 -> std 1, fixed_sigma (state independent)

    def _tower(self, seq, x, x16):
        from .fused_update import linear_elu

        for m in seq:
            if isinstance(m, nn.Linear):
                x, x16 = linear_elu(x, m.weight, m.bias, x16, self.direct_grad)
        return x

    def forward(self, obs, obs16=None):
        if self.fused_layers and obs.is_cuda and torch.is_grad_enabled():
            # training pass through the library's kernels around the cuBLAS GEMMs (learning/fused_update.py)
            from .fused_update import heads

            x16 = (obs16 if obs16 is not None else obs.to(torch.bfloat16)) if self.bf16_wgrad else None
            h = self._tower(self.trunk, obs, x16)
            if not self.separate:
                mu, v = heads(h, self.mu.weight, self.mu.bias, self.value.weight, self.value.bias, self.direct_grad)
                return mu, self.log_std.expand(obs.shape[0], -1), v
            hv = self._tower(self.critic_trunk, obs, x16)
            return self.mu(h), self.log_std.expand(obs.shape[0], -1), self.value(hv).squeeze(-1)
        h = self.trunk(obs)
        hv = self.critic_trunk(obs) if self.separate else h
        return self.mu(h), self.log_std.expand(obs.shape[0], -1), self.value(hv).squeeze(-1)


def compute_gae(rew, val, done, v_last, gamma, tau):
    """Generalised advantage estimation exactly as rl_games' ``a2c_common.discount_values`` runs it for the hot-path tasks (the same
    recursion the fork states in-tree for its AMP agent, ``learning/common_agent.py:406-418``): with ``done[t]`` the flag returned by
    step t, delta_t = r_t + gamma V_{t+1} (1 - done_t) - V_t and A_t = delta_t + gamma tau (1 - done_t) A_{t+1}.
    ``rew``, ``val``, ``done``: (T, N); ``v_last``: (N) value of the observation after the last step.  Returns the advantages (T, N)."""
    T = rew.shape[0]
    adv = torch.zeros_like(rew)
    last = torch.zeros_like(v_last)
    for t in reversed(range(T)):
        nv = v_last if t == T - 1 else val[t + 1]
        nonterm = 1.0 - done[t]
        delta = rew[t] + gamma * nv * nonterm - val[t]
        last = delta + gamma * tau * nonterm * last
        adv[t] = last
    return adv


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
    separate: bool = False      # network.separate of the train yaml: separate actor / critic towers
    mixed_precision: bool = False   # config.mixed_precision of the train yaml (True in every hot-path yaml): with fused_update, the K = minibatch
                                    # weight-gradient GEMMs take bf16 operands (fp32 accumulation and output); everything else stays fp32 / TF32
    reward_scale: float = 1.0   # config.reward_shaper.scale_value (Cartpole: 0.1)
    max_epochs: int = 1000
    save_frequency: int = 0     # config.save_frequency: checkpoint every so many epochs (0 = only at the end)
    tf32: bool = False          # TF32 tensor-core matmuls in the update (the reference trains with mixed_precision: True)
    fused_adam: bool = False    # single-kernel Adam (torch fused implementation)


@dataclass
class TrainLog:
    epochs: List[int] = field(default_factory=list)
    env_steps: List[int] = field(default_factory=list)
    mean_episode_reward: List[float] = field(default_factory=list)
    mean_episode_length: List[float] = field(default_factory=list)
    wall_s: List[float] = field(default_factory=list)


class PPO:
    def __init__(self, env, cfg: PPOConfig = PPOConfig(), multi_gpu: bool = False, seed: int = 42, fused_rollout: bool = False,
                 cuda_graphs: bool = False, graph_allreduce: bool = True, fused_update: bool = False):
        """``fused_rollout``: evaluate the policy during the rollout with the library's fused tcgen05 kernel
        (``learning/fused_policy.py``; bf16 operands, fp32 accumulation) instead of the torch modules. The update still
        differentiates the fp32 torch network; the behaviour policy's (mu, neglogp, value) are the kernel's.

        ``cuda_graphs``: capture the whole rollout (``horizon_length`` x [normalise, policy, sample, ``env.step``, bookkeeping] +
        GAE) as ONE CUDA graph and each minibatch update (forward, backward, clip, Adam) as another, so that an epoch is a
        handful of graph launches instead of ~2000 kernel launches from Python.  Needs a task whose ``step()`` is free of host
        synchronisation and host-side per-step state (the fused flat tasks and Cartpole).  With several GPUs the gradient
        all-reduce (one flat NCCL all-reduce per minibatch) is captured inside the update graph (``graph_allreduce``); if the
        capture is refused the update falls back to eager launches with the same all-reduce.

        ``fused_update``: the loss head (clipped surrogate, clipped value loss, bound loss, entropy, KL and their gradients) as ONE
        kernel of the library, parameters and gradients as one flat vector each, gradient clipping + Adam as two launches
        (``learning/fused_update.py``); the linear layers stay with torch / cuBLAS."""
        self.env, self.cfg, self.multi_gpu = env, cfg, multi_gpu
        self.fused = self.fused_critic = None
        # a step() that synchronises with the host cannot be captured: such tasks roll out eagerly whatever the caller asked for
        self.cuda_graphs = bool(cuda_graphs) and not getattr(env, "needs_host_sync", False)
        self.device = env.rl_device
        torch.manual_seed(seed)
        self.model = ActorCritic(env.num_obs, env.num_acts, cfg.units, separate=cfg.separate).to(self.device)
        self.obs_rms = RunningMeanStd((env.num_obs,)).to(self.device)
        self.val_rms = RunningMeanStd(()).to(self.device)
        self.lr_t = torch.tensor(cfg.learning_rate, device=self.device, dtype=torch.float32)
        # the minibatch update is captured on one GPU, and on several when the NCCL all-reduce of the gradient can be captured with it
        graph_update = self.cuda_graphs and (not multi_gpu or graph_allreduce)
        self.graph_allreduce = bool(graph_allreduce) and multi_gpu
        self.checkpoint_path = None
        self.update_capture_error = None
        if cfg.tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        self.fused_update = bool(fused_update)
        self.flatp = self.head = None
        if self.fused_update:
            from .fused_update import FlatParameters, FusedClipAdam

            world = 1
            if multi_gpu:
                import torch.distributed as dist

                world = dist.get_world_size()
            self.flatp = FlatParameters(self.model)
            self.model.fused_layers = all(u % 4 == 0 and u <= 1024 for u in cfg.units) and cfg.units[-1] <= 256
            self.model.bf16_wgrad = bool(cfg.mixed_precision)
            self.opt = FusedClipAdam(self.flatp, lr=self.lr_t, eps=1e-8, max_grad_norm=cfg.grad_norm, grad_scale=1.0 / world)
        else:
            self.opt = torch.optim.Adam(self.model.parameters(), lr=self.lr_t if graph_update else cfg.learning_rate, eps=1e-8,
                                        capturable=graph_update, fused=True if cfg.fused_adam else None)
        self.lr = cfg.learning_rate
        if multi_gpu:
            import torch.distributed as dist

            for p in self.model.parameters():
                dist.broadcast(p.data, 0)
        if fused_rollout:
            from .fused_policy import FusedPolicy

            self.fused = FusedPolicy(env.num_obs, env.num_acts, cfg.units, self.device)
            # separate towers: a second instance of the kernel evaluates the critic tower (its action head is unused)
            self.fused_critic = FusedPolicy(env.num_obs, env.num_acts, cfg.units, self.device) if cfg.separate else None
        n, T, dev = env.num_envs, cfg.horizon_length, self.device
        self.ep_rew = torch.zeros(n, device=dev)
        self.ep_len = torch.zeros(n, device=dev)
        # finished-episode accumulators (device side: no host sync inside the rollout)
        self.fin = torch.zeros(3, device=dev, dtype=torch.float64)         # sum of rewards, sum of lengths, count
        self._last_stats = (0.0, 0.0)
        # rollout storage (static: the graphs replay into these)
        self.obs = torch.zeros(n, env.num_obs, device=dev)
        self.b_obs = torch.zeros(T, n, env.num_obs, device=dev)
        self.b_act = torch.zeros(T, n, env.num_acts, device=dev)
        self.b_mu = torch.zeros(T, n, env.num_acts, device=dev)
        self.b_nlp, self.b_val, self.b_rew, self.b_done = (torch.zeros(T, n, device=dev) for _ in range(4))
        self.f_ret = torch.zeros(T * n, device=dev)
        self.f_val = torch.zeros(T * n, device=dev)
        self.f_adv = torch.zeros(T * n, device=dev)
        self.mb = min(cfg.minibatch_size, T * n)
        self.idx = torch.zeros(self.mb, dtype=torch.long, device=dev)
        self.kl_acc = torch.zeros((), device=dev)
        if self.fused_update:
            from .fused_update import PpoHead

            from .fused_update import RolloutKernels

            self.rk = RolloutKernels(self, seed)
            self.head = PpoHead(self.mb, env.num_acts, dev, cfg.e_clip, cfg.critic_coef, cfg.entropy_coef, cfg.bounds_loss_coef)
            self.head.bind(self.idx, self.b_act.view(T * n, -1), self.b_mu.view(T * n, -1), self.b_nlp.view(-1), self.f_adv, self.f_val, self.f_ret)
            self.mb_obs = torch.zeros(self.mb, env.num_obs, device=dev)
            self.mb_obs16 = torch.zeros(self.mb, env.num_obs, device=dev, dtype=torch.bfloat16)
            self.model.direct_grad = True
        self._g_rollout = None
        self._g_update = None

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

    # ------------------------------------------------------------------ rollout (static shapes, no host sync)
    @torch.no_grad()
    def _rollout_kernels(self):
        """The rollout with the library's bookkeeping kernels (``fused_update=True``): per step running statistics (2 launches + count),
        normalise-and-store, policy, sample-and-store, the environment step, reward / done / episode bookkeeping; then GAE, return
        statistics and the three normalisations in six launches.  Same mathematics as :meth:`_rollout`; the Gaussian noise comes
        from the kernel's own Philox stream instead of torch's generator."""
        cfg, env, rk = self.cfg, self.env, self.rk
        T, N = cfg.horizon_length, env.num_envs
        obs = self.obs
        if self.fused is not None:
            self.fused.sync(self.model)
            self.fused.set_obs_norm(None, None, 1e-5, 1e30)          # observations reach the policy kernel already normalised
            if self.fused_critic is not None:
                self.fused_critic.sync(self.model, critic=True)
                self.fused_critic.set_obs_norm(None, None, 1e-5, 1e30)
        log_std = self.model.log_std
        for t in range(T):
            rk.obs_stats(obs)
            rk.normalize(obs, self.b_obs[t])
            if self.fused is not None:
                self.fused.forward(self.b_obs[t], rk.mu, rk.value)
                if self.fused_critic is not None:
                    self.fused_critic.forward(self.b_obs[t], rk.mu_scratch, rk.value)
                mu, v = rk.mu, rk.value
            else:
                mu, _, v = self.model(self.b_obs[t])
                mu, v = mu.contiguous(), v.contiguous()
            rk.sample(t, mu, v, log_std)
            o, rew, done, extras = env.step(rk.env_actions)
            obs = o["obs"]
            rk.post(t, rew, done, extras["time_outs"])
        self.obs.copy_(obs)
        rk.normalize(self.obs, rk.nobs_last)
        _, _, v_last = self.model(rk.nobs_last)
        v_last = self.val_rms.denormalize(v_last).contiguous()
        rk.gae_finish(v_last)

    @torch.no_grad()
    def _rollout(self):
        if self.fused_update and self.device != "cpu":
            return self._rollout_kernels()
        cfg, env = self.cfg, self.env
        T, N = cfg.horizon_length, env.num_envs
        obs = self.obs
        if self.fused is not None:
            self.fused.sync(self.model)
            if self.fused_critic is not None:
                self.fused_critic.sync(self.model, critic=True)
        log_std = self.model.log_std.expand(N, -1)
        for t in range(T):
            self.obs_rms.update(obs)
            nobs = self.obs_rms.normalize(obs)
            if self.fused is not None:
                self.fused.set_obs_norm(self.obs_rms.mean, self.obs_rms.var, self.obs_rms.eps, 5.0)
                mu, v = self.fused.forward(obs)
                if self.fused_critic is not None:
                    self.fused_critic.set_obs_norm(self.obs_rms.mean, self.obs_rms.var, self.obs_rms.eps, 5.0)
                    _, v = self.fused_critic.forward(obs)
            else:
                mu, _, v = self.model(nobs)
            act = mu + log_std.exp() * torch.randn_like(mu)
            self.b_obs[t], self.b_act[t], self.b_mu[t], self.b_nlp[t] = nobs, act, mu, neglogp(act, mu, log_std)
            self.b_val[t] = self.val_rms.denormalize(v)
            o, rew, done, extras = env.step(torch.clamp(act, -1.0, 1.0))
            obs.copy_(o["obs"])
            # rl_games play_steps: shaped = reward_shaper(r) (+ gamma V time_outs: value_bootstrap, docs/release_notes.md:67) goes to the
            # buffer; the episode statistics accumulate the RAW reward
            shaped = cfg.reward_scale * rew + cfg.gamma * self.b_val[t] * extras["time_outs"].float()
            donef = (done != 0).float()
            self.b_rew[t], self.b_done[t] = shaped, donef
            self.ep_rew += rew
            self.ep_len += 1
            self.fin[0] += (self.ep_rew * donef).sum()
            self.fin[1] += (self.ep_len * donef).sum()
            self.fin[2] += donef.sum()
            self.ep_rew *= 1.0 - donef
            self.ep_len *= 1.0 - donef
        _, _, v_last = self.model(self.obs_rms.normalize(obs))
        v_last = self.val_rms.denormalize(v_last)
        adv = compute_gae(self.b_rew, self.b_val, self.b_done, v_last, cfg.gamma, cfg.tau)
        ret = adv + self.b_val
        self.val_rms.update(ret)
        self.f_ret.copy_((ret.reshape(-1) - self.val_rms.mean.float()) / torch.sqrt(self.val_rms.var.float() + 1e-5))
        self.f_val.copy_((self.b_val.reshape(-1) - self.val_rms.mean.float()) / torch.sqrt(self.val_rms.var.float() + 1e-5))
        a = adv.reshape(-1)
        self.f_adv.copy_((a - a.mean()) / (a.std() + 1e-8))

    # ------------------------------------------------------------------ one minibatch update on self.idx
    def _update(self):
        cfg = self.cfg
        T, N = cfg.horizon_length, self.env.num_envs
        idx = self.idx
        if self.fused_update:
            from .fused_update import gather_rows

            f_obs = self.b_obs.reshape(T * N, -1)
            if f_obs.shape[1] % 4 == 0:      # minibatch rows as float32 + bf16 in one pass
                gather_rows(f_obs, idx, self.mb_obs, self.mb_obs16 if self.model.bf16_wgrad else None)
                mu, _, v = self.model(self.mb_obs, self.mb_obs16 if self.model.bf16_wgrad else None)
            else:
                mu, _, v = self.model(f_obs[idx])
            self.flatp.grad.zero_()          # parameters outside the fused layers (separate-tower heads) still accumulate
            self.head.backward_direct(mu, v, self.model.log_std)
            if self.multi_gpu:
                import torch.distributed as dist

                dist.all_reduce(self.flatp.grad)          # SUM over ranks; the 1 / world factor is the optimiser's grad_scale
            self.opt.step()                                # global-norm clip + Adam
            self.kl_acc += self.head.out[4]
            return
        f_obs, f_act = self.b_obs.reshape(T * N, -1), self.b_act.reshape(T * N, -1)
        f_nlp, f_mu = self.b_nlp.reshape(-1), self.b_mu.reshape(T * N, -1)
        mu, log_std, v = self.model(f_obs[idx])
        nlp = neglogp(f_act[idx], mu, log_std)
        ratio = torch.exp(f_nlp[idx] - nlp)
        a = self.f_adv[idx]
        a_loss = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - cfg.e_clip, 1.0 + cfg.e_clip)).mean()
        fv, fr = self.f_val[idx], self.f_ret[idx]
        v_clip = fv + (v - fv).clamp(-cfg.e_clip, cfg.e_clip)
        c_loss = torch.max((v - fr) ** 2, (v_clip - fr) ** 2).mean()
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
            self.kl_acc += (((mu - f_mu[idx]) ** 2) / (2.0 * torch.exp(2.0 * log_std))).sum(-1).mean()

    @torch.no_grad()
    def _sync_normalisers(self):
        """Several GPUs: every rank has the same weights, so it must also normalise inputs and value targets the same way.  After each
        rollout the running moments are pooled over the ranks (equal sample counts per rank: mean of the means, mean of the second
        moments) -- one small all-reduce per epoch instead of one per step."""
        import torch.distributed as dist

        w = dist.get_world_size()
        for rms in (self.obs_rms, self.val_rms):
            m2 = rms.var + rms.mean * rms.mean
            buf = torch.cat([rms.mean.reshape(-1), m2.reshape(-1)])
            dist.all_reduce(buf)
            buf /= w
            k = rms.mean.numel()
            mean = buf[:k].view_as(rms.mean)
            rms.mean.copy_(mean)
            rms.var.copy_((buf[k:].view_as(rms.var) - mean * mean).clamp_min(1e-12))

    def _capture(self, fn, warmup=2):
        """Warm ``fn`` up on a side stream (cuBLAS handles, lazy state), then capture it."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:      # the parameters' accumulation nodes were created by the eager warm-up, on another stream: expected here
            quiet(False)
        with torch.cuda.graph(g):
            fn()
        return g

    def release_graphs(self):
        """Drop the captured CUDA graphs (call before ``torch.distributed.destroy_process_group``: tearing a NCCL communicator down
        while graphs that captured its all-reduces are alive has been seen to hang)."""
        import gc

        self._g_rollout = self._g_update = None
        gc.collect()
        if str(self.device) != "cpu":
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ checkpoints / evaluation
    def state_dict(self):
        """What rl_games keeps in its ``.pth`` files (``a2c_common.get_full_state_weights``): network, normalisers, optimiser."""
        return {"model": self.model.state_dict(), "obs_rms": self.obs_rms.state_dict(), "val_rms": self.val_rms.state_dict(),
                "optimizer": self.opt.state_dict(), "lr": float(self.lr_t) if torch.is_tensor(self.lr_t) else self.lr, "units": tuple(self.cfg.units)}

    def load_state_dict(self, sd, load_optimizer=True):
        self.model.load_state_dict(sd["model"])
        self.obs_rms.load_state_dict(sd["obs_rms"])
        self.val_rms.load_state_dict(sd["val_rms"])
        if load_optimizer and "optimizer" in sd:
            try:
                self.opt.load_state_dict(sd["optimizer"])
            except (ValueError, KeyError):
                pass        # optimiser flavour differs (capturable / fused): keep the fresh one
        self.lr = float(sd.get("lr", self.lr))
        self.lr_t.fill_(self.lr)

    def save(self, path):
        import os

        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(self.state_dict(), path)

    def load(self, path, load_optimizer=True):
        self.load_state_dict(torch.load(path, map_location=self.device), load_optimizer)

    @torch.no_grad()
    def play(self, steps=1000, deterministic=True):
        """``test=True`` of the reference's train.py: run the policy without learning; returns (mean episode reward, length)."""
        env = self.env
        obs = env.reset()["obs"].clone()
        ep_rew = torch.zeros(env.num_envs, device=self.device)
        ep_len = torch.zeros(env.num_envs, device=self.device)
        tot = torch.zeros(3, device=self.device, dtype=torch.float64)
        for _ in range(steps):
            mu, log_std, _ = self.model(self.obs_rms.normalize(obs))
            act = mu if deterministic else mu + log_std.exp() * torch.randn_like(mu)
            o, rew, done, _ = env.step(torch.clamp(act, -1.0, 1.0))
            obs = o["obs"].clone()
            donef = (done != 0).float()
            ep_rew += rew
            ep_len += 1
            tot[0] += (ep_rew * donef).sum()
            tot[1] += (ep_len * donef).sum()
            tot[2] += donef.sum()
            ep_rew *= 1.0 - donef
            ep_len *= 1.0 - donef
        t = tot.tolist()
        return (t[0] / t[2], t[1] / t[2]) if t[2] > 0 else (float(ep_rew.mean()), float(ep_len.mean()))

    def _episode_stats(self):
        s = self.fin.tolist()
        if s[2] > 0:
            self._last_stats = (s[0] / s[2], s[1] / s[2])
        self.fin.zero_()
        return self._last_stats

    def train(self, max_epochs=None, log_every=10, verbose=False) -> TrainLog:
        cfg, env = self.cfg, self.env
        T, N = cfg.horizon_length, env.num_envs
        log = TrainLog()
        self.obs.copy_(env.reset()["obs"])
        graph_update = self.cuda_graphs and (not self.multi_gpu or self.graph_allreduce)
        if self.multi_gpu:
            import torch.distributed as dist

            for rms in (self.obs_rms, self.val_rms):      # same starting point on every rank
                for b in (rms.mean, rms.var, rms.count):
                    dist.broadcast(b, 0)
        if self.cuda_graphs and self._g_rollout is None:
            if hasattr(env, "enable_device_step_counter"):      # rough-terrain tasks: no per-step host state inside the graph
                env.enable_device_step_counter(True)
            # the warm-up passes are real rollouts/updates on the live state (a few extra environment steps before epoch 0)
            self._g_rollout = self._capture(self._rollout, warmup=1)
            if graph_update:
                self.idx.copy_(torch.randperm(T * N, device=self.device)[:self.mb])
                try:
                    self._g_update = self._capture(self._update, warmup=2)
                except Exception as exc:      # NCCL refused to be captured (old library, watchdog settings): eager update, same math
                    if not self.multi_gpu:
                        raise
                    self._g_update = None
                    self.update_capture_error = f"{type(exc).__name__}: {exc}"[:200]
                    torch.cuda.synchronize(self.device)
                    graph_update = False
                    # the capturable Adam keeps its learning rate in a device tensor either way
            self.fin.zero_()
        t0 = time.time()
        steps = 0
        kl_lo, kl_hi = 0.5 * cfg.kl_threshold, 2.0 * cfg.kl_threshold
        n_mb = max((T * N) // self.mb, 1)
        total_epochs = max_epochs or cfg.max_epochs
        for epoch in range(total_epochs):
            if self._g_rollout is not None:
                self._g_rollout.replay()
            else:
                self._rollout()
            if self.multi_gpu:
                self._sync_normalisers()
            steps += T * N
            for _ in range(cfg.mini_epochs):
                perm = self.rk.permutation(T * N) if self.fused_update else torch.randperm(T * N, device=self.device)
                self.kl_acc.zero_()
                for k in range(n_mb):
                    self.idx.copy_(perm[k * self.mb:(k + 1) * self.mb])
                    if self._g_update is not None:
                        self._g_update.replay()
                    else:
                        self._update()
                kl = self.kl_acc / n_mb
                if self.multi_gpu:
                    import torch.distributed as dist

                    dist.all_reduce(kl)
                    kl /= dist.get_world_size()
                if torch.is_tensor(self.opt.param_groups[0]["lr"]):       # adaptive learning rate on the device (the capturable Adam reads self.lr_t)
                    lr = self.lr_t
                    self.lr_t.copy_(torch.where(kl > kl_hi, torch.clamp(lr / 1.5, min=1e-6), torch.where(kl < kl_lo, torch.clamp(lr * 1.5, max=1e-2), lr)))
                else:
                    kl = float(kl)
                    if kl > kl_hi:
                        self.lr = max(self.lr / 1.5, 1e-6)
                    elif kl < kl_lo:
                        self.lr = min(self.lr * 1.5, 1e-2)
                    for g in self.opt.param_groups:
                        g["lr"] = self.lr
            if (epoch + 1) % log_every == 0 or epoch == 0 or epoch + 1 == total_epochs:
                if torch.is_tensor(self.opt.param_groups[0]["lr"]):
                    self.lr = float(self.lr_t)
                r, l = self._episode_stats()
                log.epochs.append(epoch + 1)
                log.env_steps.append(steps)
                log.mean_episode_reward.append(r)
                log.mean_episode_length.append(l)
                log.wall_s.append(time.time() - t0)
                if verbose:
                    print(f"epoch {epoch + 1:5d} env_steps {steps:10d} ep_rew {r:8.3f} ep_len {l:7.1f} lr {self.lr:.2e} wall {time.time() - t0:6.1f}s", flush=True)
            if cfg.save_frequency > 0 and self.checkpoint_path and (epoch + 1) % cfg.save_frequency == 0 and epoch + 1 < total_epochs:
                self.save(self.checkpoint_path)       # config.save_frequency of the train yaml: something survives a crash
        return log
