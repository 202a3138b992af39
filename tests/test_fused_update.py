"""Learner-side kernels (b2g_ppo_head, b2g_adam_clip_step; isaacgymenv_b200/learning/fused_update.py).

CPU: the closed-form gradients the loss-head kernel writes, restated in numpy, against torch autograd of the plain-torch loss (ties of
torch.max inside the clip range included).  GPU: the kernels against the same torch reference and against torch.optim.Adam +
clip_grad_norm_; the PPO learner with the fused update against the torch update."""
import numpy as np
import pytest
import torch

from isaacgymenv_b200.learning.fused_update import ppo_head_reference


def _case(B=257, A=12, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    mu, value, log_std = 1.5 * r(B, A), r(B), 0.3 * r(A)
    actions, old_mu = mu + 0.7 * r(B, A), mu + 0.05 * r(B, A)
    nlp = 0.5 * (((actions - mu) / log_std.exp()) ** 2).sum(-1) + log_std.sum() + 0.5 * A * 1.8378770664093453
    old_nlp = nlp + 0.3 * r(B)                     # ratios on both sides of the clip range
    old_nlp[: B // 4] = nlp[: B // 4]              # ... and exactly 1 (the tie of torch.max)
    adv, old_val = r(B), value + 0.3 * r(B)
    old_val[: B // 5] = value[: B // 5]
    ret = r(B)
    t = [mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret]
    return [x.to(device) for x in t]


HYPER = dict(e_clip=0.2, critic_coef=2.0, entropy_coef=0.001, bounds_loss_coef=0.001)


def _closed_form(mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret, e_clip, critic_coef, entropy_coef, bounds_loss_coef, mu_bound=1.1):
    """The arithmetic of k_ppo_head (csrc/b2g_learner.cu), vectorised in numpy float64."""
    mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret = (np.asarray(x, np.float64) for x in (mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret))
    B, A = mu.shape
    isg = np.exp(-log_std)
    z = (actions - mu) * isg
    nlp = 0.5 * (z ** 2).sum(1) + log_std.sum() + 0.5 * A * 1.8378770664093453
    ratio = np.exp(old_nlp - nlp)
    s1, s2 = -adv * ratio, -adv * np.clip(ratio, 1 - e_clip, 1 + e_clip)
    g_nlp = np.where(s1 >= s2, adv * ratio, 0.0) / B
    hi, lo = np.maximum(mu - mu_bound, 0), np.maximum(-mu_bound - mu, 0)
    g_mu = g_nlp[:, None] * (-z * isg) + bounds_loss_coef * 2 * (hi - lo) / B
    g_ls = (g_nlp[:, None] * (1 - z ** 2)).sum(0) - entropy_coef
    dv = value - old_val
    vc = old_val + np.clip(dv, -e_clip, e_clip)
    l1, l2 = (value - ret) ** 2, (vc - ret) ** 2
    inside = (dv >= -e_clip) & (dv <= e_clip)
    g_v = np.where(l1 > l2, 2 * (value - ret), np.where(l2 > l1, np.where(inside, 2 * (vc - ret), 0.0), (value - ret) + np.where(inside, vc - ret, 0.0)))
    g_v = 0.5 * critic_coef * g_v / B
    return g_mu, g_v, g_ls


def test_closed_form_gradients_equal_autograd():
    mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret = [x.double() for x in _case()]
    mu.requires_grad_(), value.requires_grad_(), log_std.requires_grad_()
    loss = ppo_head_reference(mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret, **HYPER)[0]
    loss.backward()
    g_mu, g_v, g_ls = _closed_form(mu.detach(), value.detach(), log_std.detach(), actions, old_mu, old_nlp, adv, old_val, ret, **HYPER)
    np.testing.assert_allclose(g_mu, mu.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(g_v, value.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(g_ls, log_std.grad.numpy(), rtol=1e-10, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("B,A,gather", [(257, 12, False), (32768, 12, True), (1000, 18, True), (64, 1, False)])
def test_ppo_head_kernel_matches_torch(B, A, gather):
    from isaacgymenv_b200.learning.fused_update import PpoHead

    dev = "cuda:0"
    mu, value, log_std, actions, old_mu, old_nlp, adv, old_val, ret = _case(B, A, seed=B, device=dev)
    idx = None
    full = [actions, old_mu, old_nlp, adv, old_val, ret]
    if gather:            # the rollout buffers are 3x larger and shuffled: the kernel gathers through the index vector
        n_full = 3 * B
        idx = torch.randperm(n_full, device=dev)[:B].contiguous()
        big = []
        for t in full:
            b = torch.randn(n_full, *t.shape[1:], device=dev)
            b[idx] = t
            big.append(b.contiguous())
        full = big
    head = PpoHead(B, A, dev, **HYPER)
    head.bind(idx, *full)
    mu_k, v_k, ls_k = mu.clone().requires_grad_(), value.clone().requires_grad_(), log_std.clone().requires_grad_()
    loss_k = head.loss(mu_k, v_k, ls_k)
    (3.0 * loss_k).backward()                       # an upstream factor must scale the gradients
    mu_t, v_t, ls_t = mu.clone().requires_grad_(), value.clone().requires_grad_(), log_std.clone().requires_grad_()
    ref = ppo_head_reference(mu_t, v_t, ls_t, actions, old_mu, old_nlp, adv, old_val, ret, **HYPER)
    (3.0 * ref[0]).backward()
    torch.cuda.synchronize()
    out = head.out.cpu().numpy()
    np.testing.assert_allclose(out, np.array([float(x) for x in ref]), rtol=2e-4, atol=1e-6)
    assert float(loss_k) == pytest.approx(float(ref[0]), rel=2e-4)
    scale = float(mu_t.grad.abs().max())
    np.testing.assert_allclose(mu_k.grad.cpu().numpy(), mu_t.grad.cpu().numpy(), rtol=1e-3, atol=2e-5 * scale)
    np.testing.assert_allclose(v_k.grad.cpu().numpy(), v_t.grad.cpu().numpy(), rtol=1e-3, atol=2e-5 * float(v_t.grad.abs().max()))
    np.testing.assert_allclose(ls_k.grad.cpu().numpy(), ls_t.grad.cpu().numpy(), rtol=1e-3, atol=1e-5)
    # deterministic: a second launch reproduces the first bit for bit
    o1, g1 = head.out.clone(), head.grad_mu.clone()
    head.launch(mu, value, log_std)
    torch.cuda.synchronize()
    assert torch.equal(o1, head.out) and torch.equal(g1, head.grad_mu)


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm", [1.0, 0.0])
def test_fused_clip_adam_matches_torch(max_norm):
    from isaacgymenv_b200.learning.fused_update import FlatParameters, FusedClipAdam

    dev = "cuda:0"
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(48, 256), torch.nn.ELU(), torch.nn.Linear(256, 13)).to(dev)
    net_b = torch.nn.Sequential(torch.nn.Linear(48, 256), torch.nn.ELU(), torch.nn.Linear(256, 13)).to(dev)
    net_b.load_state_dict(net_a.state_dict())
    flat = FlatParameters(net_a)
    lr = torch.tensor(3e-4, device=dev)
    opt_a = FusedClipAdam(flat, lr=lr, eps=1e-8, max_grad_norm=max_norm, grad_scale=0.5)
    opt_b = torch.optim.Adam(net_b.parameters(), lr=3e-4, eps=1e-8)
    x = torch.randn(512, 48, device=dev)
    for step in range(25):
        y = torch.randn(512, 13, device=dev)
        opt_a.zero_grad()
        (2.0 * 40.0 * ((net_a(x) - y) ** 2).mean()).backward()        # twice the gradient, halved by grad_scale
        opt_a.step()
        opt_b.zero_grad()
        (40.0 * ((net_b(x) - y) ** 2).mean()).backward()
        if max_norm > 0:
            norm_b = torch.nn.utils.clip_grad_norm_(net_b.parameters(), max_norm)
            torch.cuda.synchronize()
            assert float(opt_a.norm[0]) == pytest.approx(float(norm_b), rel=1e-4)
        opt_b.step()
        if step == 10:
            lr.fill_(1e-3)                                       # the adaptive schedule rewrites the device scalar
            for g in opt_b.param_groups:
                g["lr"] = 1e-3
    torch.cuda.synchronize()
    assert int(opt_a.step_count) == 25
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        np.testing.assert_allclose(pa.detach().cpu().numpy(), pb.detach().cpu().numpy(), rtol=2e-4, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [False, True])
def test_ppo_fused_update_tracks_torch_update(graphs):
    """Same seed, same environment: the learner with the fused loss head / clip / Adam kernels follows the torch update (same losses,
    same parameters to float32 rounding) over a few epochs."""
    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    res = []
    for fused in (False, True):
        env = isaacgymenv_b200.make(seed=3, task="Cartpole", num_envs=256, sim_device="cuda:0", rl_device="cuda:0", headless=True)
        cfg = PPOConfig(horizon_length=8, minibatch_size=1024, mini_epochs=2, units=(32, 32, 16), learning_rate=3e-4, kl_threshold=1e9)
        torch.manual_seed(11)
        ppo = PPO(env, cfg, seed=5, cuda_graphs=graphs, fused_update=fused)
        ppo.train(max_epochs=1, log_every=1)
        torch.cuda.synchronize()
        res.append(torch.cat([p.detach().reshape(-1) for p in ppo.model.parameters()]).cpu())
        assert all(torch.isfinite(p).all() for p in ppo.model.parameters())
    # one epoch = 4 optimiser steps from identical data: parameters agree to rounding (later epochs diverge chaotically through the env)
    assert (res[0] - res[1]).abs().max().item() < 2e-5, (res[0] - res[1]).abs().max().item()
