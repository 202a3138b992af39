"""Manipulator (reference ``tasks/manipulator.py``): the arm reach task on the 7-DOF Franka.  Reward function, the
7-joint operational-space law and the reset logic against golden vectors produced by the reference's own ``compute_franka_reward`` /
``Manipulator._compute_osc_torques`` / ``Manipulator.reset_idx`` (tests/golden/gen_golden.py --manipulator-only); the model compiler on the
reference's URDF (whose hand and fingers lie behind the end of the ``<robot>`` element); on the GPU the task on both execution paths."""
import os

import numpy as np
import pytest
import torch

from oracle import task_math as tm
from tests import kernel_checks as kc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "manipulator.npz"))


def _t(k):
    return torch.from_numpy(G[k])


def test_reward_matches_reference():
    from isaacgymenv_b200.tasks.hound_arm import compute_houndarm_reward

    rew, reset = compute_houndarm_reward(_t("reset"), _t("progress"), _t("eef_pos"), _t("eef_vel"), _t("commands"), 0.1, 0.1, 1000.0)
    np.testing.assert_allclose(rew.numpy(), G["rew"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(reset.numpy(), G["reset_out"])
    assert (G["rew"] > 0.05).any() and G["reset_out"].sum() > G["reset"].sum()


def test_osc_law_matches_reference():
    """Seven joints, six task dimensions: J is 6 x 7, the null-space projector 7 x 7, the posture term pulls towards franka_default_dof_pos,
    and the per-joint effort limits (87 / 12 N m) clamp a good part of the samples."""
    from isaacgymenv_b200.tasks.hound_arm import osc_torques

    kp, kpn = torch.full((6,), 150.0), torch.full((7,), 10.0)
    u = osc_torques(_t("mm"), _t("j_eef"), _t("dpose"), _t("eef_vel"), _t("q"), _t("qd"), kp, 2 * torch.sqrt(kp), kpn, 2 * torch.sqrt(kpn),
                    _t("default_q"), _t("effort"))
    np.testing.assert_allclose(u.numpy(), G["u"], rtol=1e-5, atol=2e-4)
    clamped = np.abs(G["u"]) >= G["effort"] - 1e-6
    assert 0.1 < clamped.mean() < 0.9
    # the oracle's numpy restatement (what the kernel tests compare against), float32 mode like the reference and float64 mode
    for exact, tol in ((False, 2e-3), (True, 2e-3)):
        uo = tm.osc_torques(G["mm"], G["j_eef"], G["dpose"], G["eef_vel"], G["q"], G["qd"], 150.0, 10.0, G["effort"], exact=exact, default_q=G["default_q"])
        np.testing.assert_allclose(uo, G["u"], rtol=1e-4, atol=tol)


def test_reset_logic_matches_reference():
    """reset_idx of the reference run on an attribute bag: commands from three draws, joint positions around the default posture from
    seven more, clamped to the limits, THEN the last two joints put back on their default without noise (:417), velocities zero,
    progress and reset_buf cleared -- the formula the kernel and emulator tests use (kernel_checks.arm_reset_positions)."""
    ids, d = G["reset_env_ids"], G["reset_draws"]
    q = kc.arm_reset_positions(G["default_q"], 0.25, d[:, 3:10], G["lower"], G["upper"], 2)
    np.testing.assert_array_equal(q, G["reset_q"][ids])
    assert np.array_equal(G["reset_q"][ids][:, -2:], np.tile(G["default_q"][-2:], (len(ids), 1)))
    rng = np.array([[-0.5, 0.5], [-0.5, 0.5], [0.2, 0.6]], np.float32)
    cmd = (rng[:, 1] - rng[:, 0]) * d[:, :3] + rng[:, 0]
    np.testing.assert_allclose(G["reset_commands"][ids], cmd, rtol=0, atol=1e-7)
    others = np.setdiff1d(np.arange(len(G["reset_q"])), ids)
    np.testing.assert_array_equal(G["reset_q"][others], G["reset_q_before"][others])
    assert (G["reset_qd"][ids] == 0).all() and (G["reset_progress"][ids] == 0).all() and (G["reset_reset"][ids] == 0).all()
    # joint 4's upper limit (-0.0698) lies below default + noise for part of the draws: the clamp is exercised
    assert (G["reset_q"][ids][:, 3] <= -0.0698 + 1e-6).all()


def test_model_compiler_on_the_reference_urdf():
    """8 links, 7 revolute joints (everything after </robot> -- the hand and the fingers inside a malformed comment -- is dropped), one
    fixed-base chain; no <inertial> in the file: masses from the convex hulls of the collision meshes at density 1000 (Isaac Gym's rule),
    which lands on the real robot's ~18 kg."""
    art = kc.load_robot("manipulator")
    assert art.fixed_base and art.num_dofs == 7 and art.num_bodies == 8 and list(art.chain_len) == [7]
    assert art.dof_names == [f"panda_joint{i}" for i in range(1, 8)] and art.body_names[-1] == "panda_link7"
    assert 17.0 < art.total_mass < 21.0 and (np.asarray(art.mass) > 0.3).all()
    for i in range(8):
        w = np.linalg.eigvalsh(np.asarray(art.inertia[i]))
        assert (w > 0).all() and w[2] < w[0] + w[1] + 1e-9          # a physical inertia tensor (triangle inequality)
    np.testing.assert_allclose(art.effort, [87, 87, 87, 87, 12, 12, 12])
    np.testing.assert_allclose(art.upper[3], -0.0698)
    ref = os.environ.get("B2G_REFERENCE_ROOT", "/root/reference")
    path = os.path.join(ref, "assets", "urdf", "franka_description", "robots", "franka_panda_manipulator.urdf")
    if os.path.isfile(path):      # the build container: the committed model is what the compiler makes of the reference's file
        from isaacgymenv_b200.model.urdf import AssetOptions, compile_urdf

        fresh = compile_urdf(path, AssetOptions(fix_base_link=True, collapse_fixed_joints=False, disable_gravity=True, thickness=0.001))
        np.testing.assert_allclose(fresh.mass, art.mass, rtol=1e-12)
        np.testing.assert_allclose(fresh.inertia, art.inertia, rtol=1e-9, atol=1e-12)
        assert fresh.dof_names == art.dof_names


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_manipulator_task_contract_and_reaching(fused):
    import isaacgymenv_b200

    n = 128
    torch.manual_seed(0)
    env = isaacgymenv_b200.make(seed=4, task="Manipulator", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                overrides={"env": {"fusedStep": fused}})
    assert env.num_obs == 10 and env.num_acts == 6 and env.num_dofs == 7 and env.num_franka_bodies == 8 and env.num_franka_dofs == 7
    assert env._j_eef.shape == (n, 6, 7) and env._mm.shape == (n, 7, 7) and env.max_episode_length == 1000
    obs = env.reset()["obs"]
    assert obs.shape == (n, 10)
    # reset posture: around franka_default_dof_pos, the last two joints exactly on it (manipulator.py:417), inside the limits
    q0 = env._q.clone()
    dflt = env.franka_default_dof_pos
    assert torch.equal(q0[:, -2:], dflt[-2:].expand(n, 2)) and (q0[:, :5] - dflt[:5]).abs().max() <= 0.25 + 1e-6
    assert ((q0 >= env.franka_dof_lower_limits - 1e-6) & (q0 <= env.franka_dof_upper_limits + 1e-6)).all()
    o, r, d, ex = env.step(torch.zeros(n, 6, device="cuda"))
    assert o["obs"].shape == (n, 10) and r.shape == (n,) and d.shape == (n,) and "time_outs" in ex and torch.isfinite(o["obs"]).all()
    assert torch.allclose(o["obs"][:, 3:7].norm(dim=-1), torch.ones(n, device="cuda"), atol=1e-4)
    # closed loop under the 7-joint OSC law: drive the end effector to a reachable target near its current position
    env._refresh()
    env.commands[:] = env.states["eef_pos"] + torch.tensor([0.05, -0.04, 0.03], device="cuda")
    d0 = (env.states["eef_pos"] - env.commands).norm(dim=-1).mean().item()
    env.progress_buf[:] = 0
    for _ in range(80):
        env._refresh()
        err = env.commands - env.states["eef_pos"]
        act = torch.cat([torch.clamp(err / 0.1, -1, 1), torch.zeros(n, 3, device="cuda")], dim=1)
        o, r, d, ex = env.step(act)
    env._refresh()
    d1 = (env.states["eef_pos"] - env.commands).norm(dim=-1).mean().item()
    assert d1 < 0.35 * d0, (d0, d1)
    assert torch.isfinite(r).all() and r.mean().item() > 0.1 * (1 - np.tanh(10 * d0))
    # the redundant joint: the null-space term keeps the posture near the default while the end effector tracks
    assert (env._q - dflt).abs().max().item() < 1.5
    # episode end (quirk Q5: reset_buf raised by the reward at episodeLength - 1, cleared by reset_idx on the next step)
    env.progress_buf[:] = env.max_episode_length - 2
    o, r, d, ex = env.step(torch.zeros(n, 6, device="cuda"))
    assert d.all() and ex["time_outs"].all()
    o, r, d, ex = env.step(torch.zeros(n, 6, device="cuda"))
    assert not d.any() and (env.progress_buf == 0).all()
    assert torch.equal(env._q[:, -2:], dflt[-2:].expand(n, 2)) or not fused      # (the generic path has integrated one more step)


@pytest.mark.gpu
def test_manipulator_fused_equals_generic_path():
    """One launch of k_houndarm_step<7> against the reference's hook structure on the gym-tensor API (torch OSC with torch.inverse,
    k_simulate, k_body_state, k_jacobian, k_mass_matrix) from the same start state and commands."""
    import isaacgymenv_b200

    n = 64
    envs = []
    for fused in (True, False):
        torch.manual_seed(5)
        envs.append(isaacgymenv_b200.make(seed=9, task="Manipulator", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                          overrides={"env": {"fusedStep": fused}}))
    f, g = envs
    g._dof_state.copy_(f._dof_state)
    g.commands.copy_(f.commands)
    g._refresh()
    gen = torch.Generator(device="cuda").manual_seed(2)
    for i in range(15):
        a = 0.6 * torch.rand(n, 6, device="cuda", generator=gen) - 0.3
        of, rf, df, _ = f.step(a)
        og, rg, dg, _ = g.step(a)
        assert not df.any() and not dg.any()
    dev = (of["obs"][:, :3] - og["obs"][:, :3]).abs().max(dim=1).values
    assert dev.median().item() < 2e-4 and dev.quantile(0.8).item() < 5e-3 and dev.max().item() < 0.2, (dev.median(), dev.quantile(0.8), dev.max())
    assert (of["obs"][:, 7:] - og["obs"][:, 7:]).abs().max().item() == 0.0
    assert (rf - rg).abs().median().item() < 1e-4


@pytest.mark.gpu
def test_manipulator_trains_with_graphs_and_fused_policy():
    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig

    env = isaacgymenv_b200.make(seed=1, task="Manipulator", num_envs=1024, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                overrides={"env": {"episodeLength": 150}})
    ppo = PPO(env, PPOConfig(horizon_length=16, minibatch_size=4096, mini_epochs=4), seed=1, fused_rollout=True, cuda_graphs=True)
    log = ppo.train(max_epochs=40, log_every=10)
    assert all(torch.isfinite(p).all() for p in ppo.model.parameters())
    assert log.mean_episode_reward[-1] > log.mean_episode_reward[0], log.mean_episode_reward
