"""The in-repo PPO learner (rl_games a2c_continuous stand-in): host logic on a toy CPU environment, and on
the GPU the eager / fused-policy / CUDA-graph variants against each other on a small Anymal batch."""
import pytest
import torch


class PointEnv:
    """N independent 2-D points pushed by the action; reward peaks at the origin; 20-step episodes."""

    def __init__(self, n=128, device="cpu"):
        self.num_envs, self.num_obs, self.num_acts = n, 4, 2
        self.rl_device = device
        g = torch.Generator().manual_seed(0)
        self._g = g
        self.pos = torch.randn(n, 2, generator=g)
        self.progress = torch.zeros(n)

    def _obs(self):
        return {"obs": torch.cat([self.pos, self.pos ** 2], dim=1)}

    def reset(self):
        return self._obs()

    def step(self, act):
        self.pos = self.pos + 0.2 * act
        self.progress += 1
        rew = torch.exp(-(self.pos ** 2).sum(-1))
        done = self.progress >= 20
        fresh = torch.randn(self.num_envs, 2, generator=self._g)
        self.pos = torch.where(done[:, None], fresh, self.pos)
        self.progress = torch.where(done, torch.zeros_like(self.progress), self.progress)
        return self._obs(), rew, done.long(), {"time_outs": done}


def test_ppo_learns_toy_env_on_cpu():
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    env = PointEnv()
    cfg = PPOConfig(horizon_length=20, minibatch_size=1280, mini_epochs=4, units=(32, 32, 16), learning_rate=1e-3)
    ppo = PPO(env, cfg, seed=1)
    log = ppo.train(max_epochs=40, log_every=10)
    assert all(torch.isfinite(p).all() for p in ppo.model.parameters())
    assert log.mean_episode_length[-1] == pytest.approx(20.0)
    # episode return of a random policy is ~4.5; a policy that walks to the origin collects > 8
    assert log.mean_episode_reward[-1] > log.mean_episode_reward[0] + 2.0, log.mean_episode_reward


@pytest.mark.gpu
@pytest.mark.parametrize("fused,graphs", [(False, False), (True, False), (False, True), (True, True)])
def test_ppo_variants_run_on_anymal(fused, graphs):
    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    env = isaacgymenv_b200.make(seed=3, task="Anymal", num_envs=256, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    cfg = PPOConfig(horizon_length=8, minibatch_size=1024, mini_epochs=2)
    ppo = PPO(env, cfg, seed=3, fused_rollout=fused, cuda_graphs=graphs)
    before = [p.detach().clone() for p in ppo.model.parameters()]
    log = ppo.train(max_epochs=6, log_every=2)
    torch.cuda.synchronize()
    assert all(torch.isfinite(p).all() for p in ppo.model.parameters())
    assert any((a - b).abs().max() > 0 for a, b in zip(before, ppo.model.parameters()))
    assert torch.isfinite(ppo.b_obs).all() and torch.isfinite(ppo.f_adv).all() and torch.isfinite(ppo.f_ret).all()
    assert log.env_steps[-1] == 6 * 8 * 256
    assert 1e-6 <= ppo.lr <= 1e-2
    if fused:
        # the behaviour policy's mean recorded by the kernel matches the fp32 network on the stored observations (bf16 tolerance)
        with torch.no_grad():
            mu, _, _ = ppo.model(ppo.b_obs[-1])
        # parameters moved during the update, so only a loose bound holds
        assert (mu - ppo.b_mu[-1]).abs().max().item() < 0.5


# ---------------------------------------------------------------------------------------------------------------------
# numpy restatements the learner is checked against
# ---------------------------------------------------------------------------------------------------------------------
def _gae_rlgames_numpy(rew, val, dones_after, v_last, gamma, tau):
    """rl_games' a2c_common.discount_values (the stock agent the reference trains Anymal with, train.py:200-218) in plain loops:
    fdones = dones after the last step, mb_fdones[t + 1] = dones returned by step t."""
    import numpy as np

    T, N = rew.shape
    adv = np.zeros((T, N))
    lastgaelam = np.zeros(N)
    for t in reversed(range(T)):
        nextnonterminal = 1.0 - dones_after[t]
        nextvalues = v_last if t == T - 1 else val[t + 1]
        delta = rew[t] + gamma * nextvalues * nextnonterminal - val[t]
        lastgaelam = delta + gamma * tau * nextnonterminal * lastgaelam
        adv[t] = lastgaelam
    return adv


def _gae_common_agent_numpy(mb_fdones, mb_values, mb_rewards, mb_next_values, gamma, tau):
    """learning/common_agent.py:406-418 as written in the fork (its AMP agent): the terminal mask sits in next_values (:281-284)."""
    import numpy as np

    T = mb_rewards.shape[0]
    lastgaelam = 0
    mb_advs = np.zeros_like(mb_rewards)
    for t in reversed(range(T)):
        not_done = 1.0 - mb_fdones[t]
        delta = mb_rewards[t] + gamma * mb_next_values[t] - mb_values[t]
        lastgaelam = delta + gamma * tau * not_done * lastgaelam
        mb_advs[t] = lastgaelam
    return mb_advs


def test_gae_matches_rlgames_and_common_agent_restatements():
    import numpy as np

    from isaacgymenv_b200.learning.ppo import compute_gae

    rng = np.random.default_rng(0)
    T, N, gamma, tau = 24, 64, 0.99, 0.95
    rew = rng.normal(size=(T, N))
    val = rng.normal(size=(T, N))
    v_last = rng.normal(size=N)
    done = (rng.random((T, N)) < 0.1).astype(np.float64)
    adv = compute_gae(torch.tensor(rew), torch.tensor(val), torch.tensor(done), torch.tensor(v_last), gamma, tau).numpy()
    np.testing.assert_allclose(adv, _gae_rlgames_numpy(rew, val, done, v_last, gamma, tau), rtol=1e-12, atol=1e-12)
    # the fork's own statement of the recursion: next_values[t] = V(obs after step t) * (1 - terminated), with terminated = done here
    next_values = np.concatenate([val[1:], v_last[None]], axis=0) * (1.0 - done)
    np.testing.assert_allclose(adv, _gae_common_agent_numpy(done, val, rew, next_values, gamma, tau), rtol=1e-12, atol=1e-12)
    # hand-checked two-step case: no terminal, gamma = tau = 1 -> A_0 = r_0 + r_1 + V_last - V_0
    a = compute_gae(torch.tensor([[1.0], [2.0]]), torch.tensor([[0.5], [0.25]]), torch.zeros(2, 1), torch.tensor([4.0]), 1.0, 1.0)
    assert a[0, 0].item() == pytest.approx(1.0 + 2.0 + 4.0 - 0.5) and a[1, 0].item() == pytest.approx(2.0 + 4.0 - 0.25)
    # a terminal at step 0 cuts both the bootstrap and the recursion
    a = compute_gae(torch.tensor([[1.0], [2.0]]), torch.tensor([[0.5], [0.25]]), torch.tensor([[1.0], [0.0]]), torch.tensor([4.0]), 1.0, 1.0)
    assert a[0, 0].item() == pytest.approx(1.0 - 0.5)


class ScriptedEnv:
    """Deterministic rewards / dones / time-outs so the rollout buffers can be recomputed by hand."""

    def __init__(self, n=8, T=6, device="cpu"):
        self.num_envs, self.num_obs, self.num_acts = n, 3, 2
        self.rl_device = device
        g = torch.Generator().manual_seed(5)
        self.rews = torch.rand(T + 4, n, generator=g)
        self.dones = (torch.rand(T + 4, n, generator=g) < 0.3)
        self.touts = self.dones & (torch.rand(T + 4, n, generator=g) < 0.5)
        self.obs_seq = torch.randn(T + 5, n, 3, generator=g)
        self.t = 0

    def reset(self):
        return {"obs": self.obs_seq[0]}

    def step(self, act):
        t = self.t
        self.t += 1
        return {"obs": self.obs_seq[t + 1]}, self.rews[t], self.dones[t].long(), {"time_outs": self.touts[t]}


def test_rollout_buffers_follow_rlgames_play_steps():
    """Rewards in the buffer are reward_shaper.scale_value * r + gamma * V * time_outs (rl_games play_steps with value_bootstrap),
    returns = GAE + values, while the logged episode reward accumulates the RAW reward."""
    import numpy as np

    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    T, n = 6, 8
    env = ScriptedEnv(n, T)
    cfg = PPOConfig(horizon_length=T, minibatch_size=T * n, mini_epochs=1, units=(8, 8, 8), reward_scale=0.1, gamma=0.97, tau=0.9)
    ppo = PPO(env, cfg, seed=2)
    ppo.obs.copy_(env.reset()["obs"])
    ppo._rollout()
    val = ppo.b_val.numpy().astype(np.float64)
    shaped = 0.1 * env.rews[:T].numpy() + 0.97 * val * env.touts[:T].numpy()
    np.testing.assert_allclose(ppo.b_rew.numpy(), shaped, rtol=1e-6, atol=1e-7)
    # raw episode statistics: sum of raw rewards of the finished episodes
    fin = ppo.fin.numpy()
    ep = np.zeros(n)
    tot = cnt = 0.0
    for t in range(T):
        ep += env.rews[t].numpy()
        d = env.dones[t].numpy()
        tot += (ep * d).sum()
        cnt += d.sum()
        ep *= 1 - d
    assert fin[0] == pytest.approx(tot, rel=1e-6) and fin[2] == cnt and cnt > 0


@pytest.mark.parametrize("separate", [False, True])
def test_actor_critic_towers(separate):
    """network.separate of the train yaml: one shared trunk (flat tasks) or separate actor / critic towers of the same shape
    (cfg/train/AnymalTerrainPPO.yaml:8, UsefulHoundPPO.yaml:8)."""
    from isaacgymenv_b200.learning.ppo import ActorCritic

    m = ActorCritic(188, 12, (512, 256, 128), separate=separate)
    n_lin = sum(1 for x in m.modules() if isinstance(x, torch.nn.Linear))
    assert n_lin == (8 if separate else 5)
    obs = torch.randn(4, 188)
    mu, ls, v = m(obs)
    assert mu.shape == (4, 12) and v.shape == (4,) and ls.shape == (4, 12)
    if separate:
        # the value must not depend on the actor tower and vice versa
        for p in m.trunk.parameters():
            p.data.zero_()
        mu2, _, v2 = m(obs)
        assert torch.equal(v, v2) and not torch.equal(mu, mu2)
    params = sum(p.numel() for p in m.parameters())
    one = 188 * 512 + 512 + 512 * 256 + 256 + 256 * 128 + 128
    assert params == (2 * one if separate else one) + 128 * 12 + 12 + 128 + 1 + 12


def test_train_cfg_plumbs_network_and_reward_shaper():
    from isaacgymenv_b200.train import load_train_config, ppo_config_from_train_cfg

    c = ppo_config_from_train_cfg(load_train_config("AnymalTerrainPPO"))
    assert c.separate and c.units == (512, 256, 128) and c.reward_scale == 1.0 and c.save_frequency == 50
    c = ppo_config_from_train_cfg(load_train_config("CartpolePPO"))
    assert not c.separate and c.reward_scale == pytest.approx(0.1)
    assert not ppo_config_from_train_cfg(load_train_config("AnymalPPO")).separate


@pytest.mark.gpu
def test_ppo_separate_towers_fused_rollout_on_terrain():
    """[512, 256, 128] separate towers on AnymalTerrain (plane): the fused tcgen05 rollout policy (two kernel instances) records the same
    means and values as the fp32 network at rollout time (bf16 tolerance), inside CUDA graphs."""
    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    env = isaacgymenv_b200.make(seed=3, task="AnymalTerrain", num_envs=256, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    cfg = PPOConfig(horizon_length=4, minibatch_size=1024, mini_epochs=1, units=(512, 256, 128), separate=True, learning_rate=0.0)
    ppo = PPO(env, cfg, seed=3, fused_rollout=True, cuda_graphs=True)
    ppo.train(max_epochs=2, log_every=1)
    torch.cuda.synchronize()
    with torch.no_grad():
        mu, _, v = ppo.model(ppo.b_obs[-1])       # learning rate 0: the parameters are those of the rollout
        v = ppo.val_rms.denormalize(v)
    assert (mu - ppo.b_mu[-1]).abs().max().item() < 5e-2
    assert (v - ppo.b_val[-1]).abs().max().item() < 5e-2 * (1 + v.abs().max().item())
