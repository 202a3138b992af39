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
@pytest.mark.parametrize("task,units,separate", [("Cartpole", (32, 32, 16), False), ("Anymal", (256, 128, 64), False), ("AnymalTerrain", (512, 256, 128), True)])
def test_ppo_fused_minibatch_update_equals_torch_update(task, units, separate):
    """One rollout, then the SAME minibatches through both update paths from the same weights: the fused one (kernel loss head with
    closed-form gradients, fused bias+ELU layers, flat clip + Adam) and the torch one (autograd through plain torch ops,
    clip_grad_norm_, torch.optim.Adam) end with the same parameters to float32 rounding."""
    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    torch.backends.cuda.matmul.allow_tf32 = False
    n, T = 256, 8
    env = isaacgymenv_b200.make(seed=3, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    cfg = PPOConfig(horizon_length=T, minibatch_size=512, mini_epochs=1, units=units, separate=separate, learning_rate=3e-4, entropy_coef=0.001,
                    bounds_loss_coef=0.001)
    a = PPO(env, cfg, seed=5, fused_update=True)
    b = PPO(env, cfg, seed=5, fused_update=False)
    b.model.load_state_dict(a.model.state_dict())
    a.obs.copy_(env.reset()["obs"])
    a._rollout()
    for name in ("b_obs", "b_act", "b_mu", "b_nlp", "b_val", "f_ret", "f_val", "f_adv"):
        getattr(b, name).copy_(getattr(a, name))
    perm = torch.randperm(T * n, device="cuda:0")
    for k in range(4):
        idx = perm[k * a.mb:(k + 1) * a.mb]
        a.idx.copy_(idx); b.idx.copy_(idx)
        a._update(); b._update()
    torch.cuda.synchronize()
    pa = torch.cat([p.detach().reshape(-1) for p in a.model.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in b.model.parameters()])
    assert torch.isfinite(pa).all()
    # Adam's first steps move every weight by about lr * sign(g): a weight whose gradient is at rounding level may step the other way in
    # one of the two paths, so a handful of the ~500 k weights may differ by up to a few learning rates; everything else agrees to rounding
    diff = (pa - pb).abs()
    assert (diff > 3e-5).float().mean().item() < 1e-3 and diff.max().item() < 4 * 4 * cfg.learning_rate, (diff.max().item(), (diff > 3e-5).float().mean().item())
    assert float(a.kl_acc) == pytest.approx(float(b.kl_acc), rel=1e-3, abs=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,k,cols", [(32768, 48, 256), (1000, 256, 128), (777, 188, 512), (64, 128, 64)])
def test_linear_elu_matches_torch(rows, k, cols):
    from isaacgymenv_b200.learning.fused_update import linear_elu

    dev = "cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(rows)
    x = torch.randn(rows, k, device=dev, generator=g)
    w = (torch.randn(cols, k, device=dev, generator=g) / k ** 0.5)
    b = 0.1 * torch.randn(cols, device=dev, generator=g)
    up = torch.randn(rows, cols, device=dev, generator=g)
    xa, wa, ba = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    xb, wb, bb = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ha, none16 = linear_elu(xa, wa, ba)
    assert none16 is None
    hb = torch.nn.functional.elu(torch.nn.functional.linear(xb, wb, bb))
    (ha * up).sum().backward()
    (hb * up).sum().backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(ha.detach().cpu().numpy(), hb.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)
    for a, bref, name in ((xa.grad, xb.grad, "dx"), (wa.grad, wb.grad, "dw"), (ba.grad, bb.grad, "db")):
        scale = float(bref.abs().max())
        np.testing.assert_allclose(a.cpu().numpy(), bref.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale, err_msg=name)
    # the first layer's input needs no gradient: none is computed
    xc = x.clone()
    hc, _ = linear_elu(xc, wa.detach().requires_grad_(), ba.detach().requires_grad_())
    hc.sum().backward()
    assert xc.grad is None


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(4096, 48), (8192, 188), (1000, 300), (98304, 1), (7, 4)])
def test_running_stat_kernel_matches_running_mean_std(rows, cols):
    import ctypes as C

    from isaacgymenv_b200 import _lib
    from isaacgymenv_b200.learning.ppo import RunningMeanStd

    dev = "cuda:0"
    lib = _lib.load()
    vp = C.c_void_p
    lib.b2g_running_stat_update.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, vp]
    lib.b2g_stat_workspace_doubles.argtypes = [C.c_int, C.c_int]
    a, b = RunningMeanStd((cols,)).to(dev), RunningMeanStd((cols,)).to(dev)
    ws = torch.zeros(int(lib.b2g_stat_workspace_doubles(rows, cols)), device=dev, dtype=torch.float64)
    mean_f, inv_f = torch.zeros(cols, device=dev), torch.zeros(cols, device=dev)
    g = torch.Generator(device=dev).manual_seed(rows + cols)
    p = lambda t: C.c_void_p(t.data_ptr())
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for k in range(4):
        x = (3.0 * torch.randn(rows, cols, device=dev, generator=g) + 10.0 * k).contiguous()
        _lib.check(lib.b2g_running_stat_update(p(x), rows, cols, p(a.mean), p(a.var), p(a.count), p(ws), p(mean_f), p(inv_f), C.c_float(1e-5), s))
        b.update(x)
    torch.cuda.synchronize()
    np.testing.assert_allclose(a.mean.cpu().numpy(), b.mean.cpu().numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(a.var.cpu().numpy(), b.var.cpu().numpy(), rtol=1e-8, atol=1e-9)
    assert float(a.count) == float(b.count) == 1 + 4 * rows
    np.testing.assert_allclose(inv_f.cpu().numpy(), (1.0 / torch.sqrt(b.var.float() + 1e-5)).cpu().numpy(), rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("task,separate,fused_policy", [("Anymal", False, True), ("Cartpole", False, False), ("AnymalTerrain", True, True)])
def test_rollout_kernels_fill_the_buffers_like_the_torch_rollout(task, separate, fused_policy):
    """After one rollout through the bookkeeping kernels every buffer satisfies the relations the torch rollout establishes: stored
    neglogp = neglogp(stored action | stored mean, log_std); stored observations are the normalised observations; shaped rewards, GAE,
    returns and the three normalisations recomputed in torch from the raw buffers; unit-variance Gaussian noise."""
    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig, compute_gae, neglogp

    n, T = 512, 8
    env = isaacgymenv_b200.make(seed=3, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    units = (512, 256, 128) if separate else (256, 128, 64)
    cfg = PPOConfig(horizon_length=T, minibatch_size=n * T, mini_epochs=1, units=units, separate=separate, reward_scale=0.5, gamma=0.97, tau=0.9)
    ppo = PPO(env, cfg, seed=7, fused_rollout=fused_policy, fused_update=True)
    ppo.obs.copy_(env.reset()["obs"])
    for _ in range(3):          # a few rollouts so that episodes end, time-outs aside
        vm0, vv0, vc0 = ppo.val_rms.mean.clone(), ppo.val_rms.var.clone(), ppo.val_rms.count.clone()
        ppo._rollout()
    torch.cuda.synchronize()
    ls = ppo.model.log_std.detach()
    np.testing.assert_allclose(ppo.b_nlp.cpu().numpy(), neglogp(ppo.b_act, ppo.b_mu, ls.expand_as(ppo.b_mu)).cpu().numpy(), rtol=1e-4, atol=1e-4)
    z = ((ppo.b_act - ppo.b_mu) / ls.exp()).flatten()
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.05 and float(z.abs().max()) < 7.0
    assert float(ppo.b_obs.abs().max()) <= 5.0 + 1e-6
    # GAE / returns / normalisations from the raw buffers
    with torch.no_grad():
        nobs = ppo.obs_rms.normalize(ppo.obs)
        _, _, v_last = ppo.model(nobs)
        # val_rms was updated by the rollout: v_last was de-normalised with the statistics BEFORE that update
        v_last = v_last * torch.sqrt(vv0.float() + 1e-5) + vm0.float()
    adv = compute_gae(ppo.b_rew, ppo.b_val, ppo.b_done, v_last, cfg.gamma, cfg.tau)
    ret = adv + ppo.b_val
    vm, vs = ppo.val_rms.mean.float(), torch.sqrt(ppo.val_rms.var.float() + 1e-5)
    np.testing.assert_allclose(ppo.f_ret.cpu().numpy(), ((ret.reshape(-1) - vm) / vs).cpu().numpy(), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(ppo.f_val.cpu().numpy(), ((ppo.b_val.reshape(-1) - vm) / vs).cpu().numpy(), rtol=1e-4, atol=1e-4)
    a = adv.reshape(-1)
    np.testing.assert_allclose(ppo.f_adv.cpu().numpy(), ((a - a.mean()) / (a.std() + 1e-8)).cpu().numpy(), rtol=2e-3, atol=2e-3)
    # the value normaliser saw exactly the returns
    from isaacgymenv_b200.learning.ppo import RunningMeanStd

    chk = RunningMeanStd(()).to("cuda:0")
    chk.mean.copy_(vm0); chk.var.copy_(vv0); chk.count.copy_(vc0)
    chk.update(ret)
    assert float(ppo.val_rms.mean) == pytest.approx(float(chk.mean), rel=1e-3, abs=1e-4) and float(ppo.val_rms.var) == pytest.approx(float(chk.var), rel=1e-3)
    assert set(torch.unique(ppo.b_done).tolist()) <= {0.0, 1.0} and float(ppo.fin[2]) >= 0
    assert torch.isfinite(ppo.f_adv).all() and torch.isfinite(ppo.b_rew).all()


@pytest.mark.gpu
def test_linear_elu_bf16_weight_gradient():
    """mixed_precision: the weight-gradient GEMM on bf16 copies written by the two kernels (no cast pass): float32 accumulation over the
    32768 rows, so the result is within bf16 input rounding of the float32 one; h, dx and db are untouched float32."""
    from isaacgymenv_b200.learning.fused_update import linear_elu

    dev = "cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(32768, 48, device=dev, generator=g)
    w1, b1 = torch.randn(256, 48, device=dev, generator=g) / 7, 0.1 * torch.randn(256, device=dev, generator=g)
    w2, b2 = torch.randn(128, 256, device=dev, generator=g) / 16, 0.1 * torch.randn(128, device=dev, generator=g)
    up = torch.randn(32768, 128, device=dev, generator=g)
    outs = []
    for use16 in (False, True):
        p = [t.clone().requires_grad_() for t in (w1, b1, w2, b2)]
        x16 = x.to(torch.bfloat16) if use16 else None
        h1, h16 = linear_elu(x, p[0], p[1], x16)
        h2, h216 = linear_elu(h1, p[2], p[3], h16)
        assert (h16 is not None) == use16 and (h216 is not None) == use16
        if use16:
            assert h16.dtype == torch.bfloat16 and torch.allclose(h16.float(), h1, rtol=1e-2, atol=1e-2)
        (h2 * up).sum().backward()
        outs.append([h2.detach()] + [t.grad for t in p])
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0])                                  # forward identical
    assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-4) and torch.allclose(outs[0][4], outs[1][4], rtol=1e-5, atol=1e-4)   # db float32
    for i in (1, 3):                                                            # dW1, dW2
        ref, got = outs[0][i], outs[1][i]
        assert (got - ref).norm().item() < 1e-2 * ref.norm().item(), ((got - ref).norm().item(), ref.norm().item())


@pytest.mark.gpu
@pytest.mark.parametrize("rows,hid,A", [(32768, 64, 12), (1000, 128, 18), (70, 64, 1), (130, 52, 24), (200, 50, 5), (300, 256, 3)])
def test_heads_backward_kernel_matches_torch(rows, hid, A):
    from isaacgymenv_b200.learning.fused_update import heads

    dev = "cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(rows)
    h = torch.randn(rows, hid, device=dev, generator=g)
    wm, bm = torch.randn(A, hid, device=dev, generator=g) / 8, torch.randn(A, device=dev, generator=g)
    wv, bv = torch.randn(1, hid, device=dev, generator=g) / 8, torch.randn(1, device=dev, generator=g)
    um, uv = torch.randn(rows, A, device=dev, generator=g), torch.randn(rows, device=dev, generator=g)
    a = [t.clone().requires_grad_() for t in (h, wm, bm, wv, bv)]
    b = [t.clone().requires_grad_() for t in (h, wm, bm, wv, bv)]
    mu_a, v_a = heads(*a)
    mu_b, v_b = torch.nn.functional.linear(b[0], b[1], b[2]), torch.nn.functional.linear(b[0], b[3], b[4]).squeeze(-1)
    ((mu_a * um).sum() + (v_a * uv).sum()).backward()
    ((mu_b * um).sum() + (v_b * uv).sum()).backward()
    torch.cuda.synchronize()
    assert torch.allclose(mu_a, mu_b, rtol=1e-5, atol=1e-5) and torch.allclose(v_a, v_b, rtol=1e-5, atol=1e-5)
    for x, y, name in zip(a, b, ("dh", "dw_mu", "db_mu", "dw_v", "db_v")):
        scale = float(y.grad.abs().max())
        assert x.grad.shape == y.grad.shape, name
        assert torch.allclose(x.grad, y.grad, rtol=1e-4, atol=2e-5 * scale), (name, float((x.grad - y.grad).abs().max()), scale)


@pytest.mark.gpu
@pytest.mark.parametrize("use16", [False, True])
def test_direct_gradient_stores_equal_autograd_accumulation(use16):
    """direct_grad=True: the fused layers and heads STORE the parameter gradients into pre-existing .grad tensors (pre-filled with junk
    here: overwrite, not accumulate) and hand autograd nothing -- same numbers as the autograd path."""
    from isaacgymenv_b200.learning.fused_update import heads, linear_elu

    dev = "cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(11)
    rows = 4096
    x = torch.randn(rows, 48, device=dev, generator=g)
    shapes = [(256, 48), (256,), (64, 256), (64,), (12, 64), (12,), (1, 64), (1,)]
    init = [torch.randn(*s, device=dev, generator=g) / 8 for s in shapes]
    um, uv = torch.randn(rows, 12, device=dev, generator=g), torch.randn(rows, device=dev, generator=g)
    res = []
    for direct in (False, True):
        p = [t.clone().requires_grad_() for t in init]
        if direct:
            for i, t in enumerate(p):      # junk everywhere the backward pass stores; the value bias (one element) still accumulates
                t.grad = torch.full_like(t, 123.0) if i != 7 else torch.zeros_like(t)
        x16 = x.to(torch.bfloat16) if use16 else None
        h1, h16 = linear_elu(x, p[0], p[1], x16, direct)
        h2, _ = linear_elu(h1, p[2], p[3], h16, direct)
        mu, v = heads(h2, p[4], p[5], p[6], p[7], direct)
        torch.autograd.backward([mu, v], [um, uv])
        res.append([t.grad.clone() for t in p])
    torch.cuda.synchronize()
    for a, b, s in zip(res[0], res[1], shapes):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(a.abs().max())), (s, float((a - b).abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,n", [(32768, 48, 196608), (1000, 188, 5000), (5, 4, 9)])
def test_gather_rows_kernel(rows, cols, n):
    from isaacgymenv_b200.learning.fused_update import gather_rows

    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(rows)
    src = torch.randn(n, cols, device=dev, generator=g)
    idx = torch.randint(0, n, (rows,), device=dev, generator=g)
    out, out16 = torch.zeros(rows, cols, device=dev), torch.zeros(rows, cols, device=dev, dtype=torch.bfloat16)
    gather_rows(src, idx, out, out16)
    torch.cuda.synchronize()
    assert torch.equal(out, src[idx]) and torch.equal(out16, src[idx].to(torch.bfloat16))
    only16 = torch.zeros_like(out16)
    gather_rows(src, idx, None, only16)
    assert torch.equal(only16, out16)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4097, 196608])
def test_random_permutation_kernel_is_a_bijection_and_moves_with_the_counter(n):
    """b2g_random_permutation (the mini-epoch shuffle of the update): every index exactly once, a different order per (seed, counter), and no
    visible structure -- the first half of the output draws about half of its entries from each half of the range."""
    import ctypes as C

    from isaacgymenv_b200 import _lib

    lib = _lib.load()
    lib.b2g_random_permutation.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.b2g_random_permutation.restype = C.c_int
    dev = "cuda:0"
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    outs = []
    for seed, ctr in ((7, 1), (7, 2), (8, 1)):
        out = torch.full((n,), -1, dtype=torch.int64, device=dev)
        _lib.check(lib.b2g_random_permutation(C.c_void_p(out.data_ptr()), n, seed, ctr, st), "perm")
        torch.cuda.synchronize()
        assert torch.equal(torch.sort(out).values, torch.arange(n, device=dev))
        outs.append(out)
    if n >= 1000:
        assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
        for o in outs:
            low = float((o[: n // 2] < n // 2).float().mean())
            assert 0.45 < low < 0.55, low
            assert float((o == torch.arange(n, device=dev)).float().mean()) < 0.01      # fixed points: ~1 expected
