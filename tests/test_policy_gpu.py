"""Fused tcgen05 policy forward (b2g_policy_*) against a plain torch fp32 evaluation of the same network.

Tolerances: the kernel multiplies bf16-rounded weights and activations with fp32 accumulation, so it is compared (a) tightly
(2e-3 absolute) with a torch fp32 evaluation that rounds operands to bf16 at the same places, and (b) loosely (5e-2) with the
un-rounded fp32 network.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(num_obs, num_act, units, seed):
    from isaacgymenv_b200.learning.ppo import ActorCritic

    torch.manual_seed(seed)
    m = ActorCritic(num_obs, num_act, units).cuda()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    return m


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _ref(model, obs, mean, var, eps, clip, emulate_bf16):
    x = torch.clamp((obs - mean) * torch.rsqrt(var + eps), -clip, clip)
    r = _bf if emulate_bf16 else (lambda t: t)
    h = r(x)
    for mod in model.trunk:
        if isinstance(mod, torch.nn.Linear):
            h = r(torch.nn.functional.elu(h.double() @ r(mod.weight).double().T + mod.bias.double()).float())
    mu = (h.double() @ r(model.mu.weight).double().T + model.mu.bias.double()).float()
    v = (h.double() @ r(model.value.weight).double().T + model.value.bias.double()).float().squeeze(-1)
    return mu, v


@pytest.mark.parametrize("num_obs,num_act,units,rows", [
    (48, 12, (256, 128, 64), 4096),      # Anymal / Hound (cfg/train/AnymalPPO.yaml)
    (48, 12, (256, 128, 64), 1),
    (48, 12, (256, 128, 64), 129),       # ragged last tile
    (48, 12, (256, 128, 64), 148 * 128 * 2 + 77),   # more tiles than SMs: the persistent tile loop
    (4, 1, (32, 32, 16), 512),           # narrow network, padded K
    (37, 15, (64, 256, 32), 1000),       # odd observation width, widest middle layer, full head
    # streamed-weights kernel (k_policy_forward_wide): the rough-terrain networks of cfg/train/AnymalTerrainPPO.yaml / UsefulHoundPPO.yaml
    (188, 12, (512, 256, 128), 4096),    # AnymalTerrain / HoundTerrain
    (204, 18, (512, 256, 128), 4096),    # UsefulHound: 19 head columns -> 32-column head
    (204, 18, (512, 256, 128), 1),
    (188, 12, (512, 256, 128), 148 * 128 + 131),   # a second tile on some CTAs, ragged last tile: the ring runs across tiles
    (203, 18, (512, 256, 128), 300),     # observation width not a multiple of 4: scalar loads
    (48, 20, (256, 128, 64), 700),       # narrow network with a 32-column head: streamed path, single column part
    (250, 31, (320, 48, 16), 257),       # 160-column halves, tiny tail layers, widest supported head
])
def test_policy_forward_matches_torch(num_obs, num_act, units, rows):
    from isaacgymenv_b200.learning.fused_policy import FusedPolicy

    model = _net(num_obs, num_act, units, seed=rows)
    g = torch.Generator(device="cuda").manual_seed(1234 + rows)
    obs = torch.randn(rows, num_obs, device="cuda", generator=g) * 2.0 + 0.3
    mean = torch.randn(num_obs, device="cuda", generator=g) * 0.5
    var = torch.rand(num_obs, device="cuda", generator=g) * 3.0 + 0.1
    pol = FusedPolicy(num_obs, num_act, units, "cuda:0")
    linears = [m for m in model.trunk if isinstance(m, torch.nn.Linear)]
    for i, lin in enumerate(linears):
        pol.set_layer(i, lin.weight, lin.bias)
    pol.set_layer(3, model.mu.weight, model.mu.bias)
    pol.set_layer(4, model.value.weight, model.value.bias)
    pol.set_obs_norm(mean, var, 1e-5, 5.0)
    mu, v = pol.forward(obs)
    torch.cuda.synchronize()
    mu_e, v_e = _ref(model, obs, mean, var, 1e-5, 5.0, True)
    mu_f, v_f = _ref(model, obs, mean, var, 1e-5, 5.0, False)
    assert torch.isfinite(mu).all() and torch.isfinite(v).all()
    assert (mu - mu_e).abs().max().item() < 2e-3, (mu - mu_e).abs().max().item()
    assert (v - v_e).abs().max().item() < 2e-3, (v - v_e).abs().max().item()
    assert (mu - mu_f).abs().max().item() < 5e-2
    assert (v - v_f).abs().max().item() < 5e-2
    assert pol.launches == 7
    # the same call again (weights and ring barriers start from scratch in every launch): bit-identical
    mu2, v2 = pol.forward(obs)
    assert torch.equal(mu, mu2) and torch.equal(v, v2)


def test_policy_identity_norm_and_sync():
    from isaacgymenv_b200.learning.fused_policy import FusedPolicy
    from isaacgymenv_b200.learning.ppo import RunningMeanStd

    model = _net(48, 12, (256, 128, 64), seed=7)
    pol = FusedPolicy(48, 12, (256, 128, 64), "cuda:0")
    pol.sync(model)                       # no normaliser: identity
    obs = torch.randn(300, 48, device="cuda")
    mu, v = pol.forward(obs)
    mu_e, v_e = _ref(model, obs, torch.zeros(48, device="cuda"), torch.ones(48, device="cuda"), 0.0, 5.0, True)
    assert (mu - mu_e).abs().max().item() < 2e-3 and (v - v_e).abs().max().item() < 2e-3
    rms = RunningMeanStd((48,)).cuda()
    rms.update(obs * 3 + 1)
    pol.sync(model, rms)
    mu, v = pol.forward(obs)
    mu_e, v_e = _ref(model, obs, rms.mean.float(), rms.var.float(), rms.eps, 5.0, True)
    assert (mu - mu_e).abs().max().item() < 2e-3 and (v - v_e).abs().max().item() < 2e-3


def test_policy_unaligned_observation_view():
    """An observation view that is not 16-byte aligned takes the scalar load path and gives the same result."""
    from isaacgymenv_b200.learning.fused_policy import FusedPolicy

    model = _net(48, 12, (256, 128, 64), seed=11)
    pol = FusedPolicy(48, 12, (256, 128, 64), "cuda:0")
    pol.sync(model)
    buf = torch.randn(700 * 48 + 1, device="cuda")
    obs_u = buf[1:].view(700, 48)
    assert obs_u.data_ptr() % 16 != 0 and obs_u.is_contiguous()
    mu_u, v_u = pol.forward(obs_u)
    mu_a, v_a = pol.forward(obs_u.clone())
    assert torch.equal(mu_u, mu_a) and torch.equal(v_u, v_a)


def test_policy_rejects_unsupported_shapes():
    from isaacgymenv_b200 import _lib
    from isaacgymenv_b200.learning.fused_policy import FusedPolicy

    with pytest.raises(_lib.B2GError):
        FusedPolicy(188, 12, (1024, 256, 128), "cuda:0")     # first layer beyond the 512 accumulator columns
    with pytest.raises(_lib.B2GError):
        FusedPolicy(188, 12, (512, 512, 128), "cuda:0")
    with pytest.raises(_lib.B2GError):
        FusedPolicy(48, 32, (256, 128, 64), "cuda:0")        # 33 head columns
    with pytest.raises(_lib.B2GError):
        FusedPolicy(300, 12, (256, 128, 64), "cuda:0")
    with pytest.raises(_lib.B2GError):
        FusedPolicy(48, 12, (250, 128, 64), "cuda:0")
