"""Houndarm: reward function and OSC law against golden vectors produced by the reference's own
``compute_houndarm_reward`` / ``Houndarm._compute_osc_torques`` (tests/golden/gen_golden.py --houndarm-only), and on the GPU
the task on the generic gym-tensor path: contract, reset quirk, and closed-loop reaching under the OSC controller (which
exercises dynamics, rigid-body state, Jacobian and mass-matrix tensors together)."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "houndarm.npz"))


def _t(k):
    return torch.from_numpy(G[k])


def test_reward_matches_reference():
    from isaacgymenv_b200.tasks.hound_arm import compute_houndarm_reward

    rew, reset = compute_houndarm_reward(_t("reset"), _t("progress"), _t("eef_pos"), _t("eef_vel")[:, :], _t("commands"), 0.1, 0.1, 150.0)
    np.testing.assert_allclose(rew.numpy(), G["rew"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(reset.numpy(), G["reset_out"])
    assert (G["rew"] > 0.05).any() and G["reset_out"].sum() > G["reset"].sum()


def test_osc_law_matches_reference():
    from isaacgymenv_b200.tasks.hound_arm import osc_torques

    kp = torch.full((6,), 150.0)
    kpn = torch.full((6,), 10.0)
    u = osc_torques(_t("mm"), _t("j_eef"), _t("dpose"), _t("eef_vel"), _t("q"), _t("qd"), kp, 2 * torch.sqrt(kp), kpn, 2 * torch.sqrt(kpn),
                    torch.zeros(6), torch.full((6,), 1000.0))
    np.testing.assert_allclose(u.numpy(), G["u"], rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
def test_houndarm_task_contract_and_reaching():
    import isaacgymenv_b200

    n = 128
    env = isaacgymenv_b200.make(seed=4, task="Houndarm", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    assert env.num_obs == 10 and env.num_acts == 6 and env.num_dofs == 6 and env.num_houndarm_bodies == 7
    assert env._j_eef.shape == (n, 6, 6) and env._mm.shape == (n, 6, 6)
    obs = env.reset()["obs"]
    assert obs.shape == (n, 10)
    # gravity is disabled for the asset: with zero torque the arm keeps its reset pose
    q0 = env._q.clone()
    o, r, d, ex = env.step(torch.zeros(n, 6, device="cuda"))
    assert o["obs"].shape == (n, 10) and r.shape == (n,) and d.shape == (n,) and "time_outs" in ex
    assert torch.isfinite(o["obs"]).all()
    # zero action = OSC holding pose with the null-space term pulling to the default: small motion only
    assert (env._q - q0).abs().max() < 0.05
    # quaternion part of the observation is unit
    assert torch.allclose(o["obs"][:, 3:7].norm(dim=-1), torch.ones(n, device="cuda"), atol=1e-4)
    # closed loop: command the end effector toward a reachable target near its current position
    env.commands[:] = env.states["eef_pos"] + torch.tensor([0.05, -0.04, 0.03], device="cuda")
    d0 = (env.states["eef_pos"] - env.commands).norm(dim=-1).mean().item()
    env.progress_buf[:] = 0
    for _ in range(60):
        err = env.commands - env.states["eef_pos"]
        act = torch.cat([torch.clamp(err / 0.1, -1, 1), torch.zeros(n, 3, device="cuda")], dim=1)
        o, r, d, ex = env.step(act)
    d1 = (env.states["eef_pos"] - env.commands).norm(dim=-1).mean().item()
    assert d1 < 0.35 * d0, (d0, d1)
    assert r.mean().item() > 0.1 * (1 - np.tanh(10 * d0)) and torch.isfinite(r).all()
    # episode end: reset_buf raised by the reward function at episodeLength-1, cleared again by reset_idx (quirk Q5)
    env.progress_buf[:] = env.max_episode_length - 2
    o, r, d, ex = env.step(torch.zeros(n, 6, device="cuda"))
    assert d.all() and ex["time_outs"].all()
    o, r, d, ex = env.step(torch.zeros(n, 6, device="cuda"))
    assert not d.any() and (env.progress_buf == 0).all()
    lo, hi = env.houndarm_dof_lower_limits, env.houndarm_dof_upper_limits
    assert ((env._q >= lo - 1e-3) & (env._q <= hi + 1e-3)).all()
