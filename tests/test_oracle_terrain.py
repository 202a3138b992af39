"""numpy restatement of the rough-terrain tasks' post_physics_step vs golden vectors produced by executing the
reference's own eager methods (tests/golden/gen_golden.py --terrain): AnymalTerrain on a plane with a push step,
AnymalTerrain on a generated heightfield with curriculum, HoundTerrain."""
import os

import numpy as np
import pytest

from oracle import task_math as tm
from tests.kernel_checks import GOLDEN


def terrain_case(name):
    g = np.load(os.path.join(GOLDEN, name))
    st = dict(root=g["root"].copy(), dof_pos=g["dof_pos"].copy(), dof_vel=g["dof_vel"].copy(), contact=g["contact"].copy(), torques=g["torques"].copy(),
              commands=g["commands"].copy(), actions=g["actions"].copy(), last_actions=g["last_actions"].copy(), last_dof_vel=g["last_dof_vel"].copy(),
              feet_air_time=g["feet_air_time"].copy(), progress=g["progress"].copy(), timeout_prev=g["timeout_prev"].copy(),
              episode_sums=g["episode_sums"].copy())
    custom = bool(g["custom_origins"])
    terrain = None
    if custom:
        st.update(terrain_levels=g["terrain_levels"].copy(), terrain_types=g["terrain_types"].copy(), env_origins=g["env_origins"].copy())
        terrain = dict(height_samples=g["height_samples"], border_size=float(g["border_size"]), hscale=float(g["hscale"]), vscale=float(g["vscale"]),
                       env_length=float(g["env_length"]), env_rows=int(g["env_rows"]), terrain_origins=g["terrain_origins"])
    else:
        st["terrain_levels"] = np.zeros(len(g["root"]), np.int64)
    hound = len(g["base_indices"]) > 0
    cfg = dict(rew_scales=g["rew_scales"], knee=g["knee"], feet=g["feet"], base_indices=g["base_indices"], base_body=0, allow_knee=bool(g["allow_knee"]),
               hound=hound, base_height_target=0.48 if hound else 0.52, noise_scale_vec=g["noise_scale_vec"], dt=float(g["dt"]), max_len=int(g["max_len"]),
               push=int(g["common_step_counter"] + 1) % int(g["push_interval"]) == 0, default_dof_pos=g["default"][0],
               init_root=np.array([0, 0, 0.62, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], np.float32), cmd_x=[-1, 1], cmd_y=[-1, 1], cmd_yaw=[-3.14, 3.14],
               custom_origins=custom, curriculum=True, terrain=terrain, max_episode_length_s=20.0, lin_vel_scale=2.0, ang_vel_scale=0.25,
               dof_pos_scale=1.0, dof_vel_scale=0.05, height_meas_scale=5.0)
    draws = dict(reset=g["reset_draws"], noise=g["noise_draws"], push=g["push_draws"])
    return g, st, cfg, draws


def check_terrain_outputs(g, st, obs, rew, reset, timeout, measured, extras, rtol=1e-5):
    assert np.array_equal(reset, g["o_reset"]) and 0 < reset.sum() < len(reset)
    assert np.array_equal(st["progress"], g["o_progress"]) and np.array_equal(timeout, g["o_timeout"])
    np.testing.assert_allclose(rew, g["o_rew"], rtol=rtol, atol=2e-7)
    np.testing.assert_allclose(measured, g["o_measured_heights"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(obs, g["o_obs"], rtol=rtol, atol=2e-6)
    np.testing.assert_allclose(st["root"], g["o_root"], rtol=rtol, atol=1e-6)
    np.testing.assert_allclose(st["dof_pos"], g["o_dof_pos"], rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(st["dof_vel"], g["o_dof_vel"], rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(st["commands"], g["o_commands"], rtol=rtol, atol=1e-6)
    np.testing.assert_allclose(st["last_actions"], g["o_last_actions"], rtol=0, atol=0)
    np.testing.assert_allclose(st["last_dof_vel"], g["o_last_dof_vel"], rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(st["feet_air_time"], g["o_feet_air_time"], rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(st["episode_sums"], g["o_episode_sums"], rtol=rtol, atol=1e-6)
    if extras is not None:
        np.testing.assert_allclose(extras, g["o_extras"], rtol=1e-4, atol=1e-6)
    if "o_terrain_levels" in g:
        assert np.array_equal(st["terrain_levels"], g["o_terrain_levels"])
        np.testing.assert_allclose(st["env_origins"], g["o_env_origins"], rtol=0, atol=0)


@pytest.mark.parametrize("name", ["anymal_terrain_plane.npz", "anymal_terrain_trimesh.npz", "hound_terrain_plane.npz"])
def test_terrain_post_physics_matches_reference(name):
    g, st, cfg, draws = terrain_case(name)
    obs, rew, reset, timeout, measured, extras = tm.terrain_post_physics(st, cfg, draws)
    check_terrain_outputs(g, st, obs, rew, reset, timeout, measured, extras)
    if "trimesh" in name:
        assert np.abs(g["o_measured_heights"]).max() > 0.05          # the scan really sees terrain
        assert (g["o_terrain_levels"] != g["terrain_levels"]).any()  # the curriculum really moved someone


def useful_case():
    g = np.load(os.path.join(GOLDEN, "useful_hound_plane.npz"))
    ds = g["dof_state"]
    st = dict(root=g["root"].copy(), dof_pos=ds[:, :12, 0].copy(), dof_vel=ds[:, :12, 1].copy(), arm_q=ds[:, 12:, 0].copy(), arm_qd=ds[:, 12:, 1].copy(),
              contact=g["contact"].copy(), torques=g["torques"].copy(), commands=g["commands"].copy(), actions=g["actions"].copy(),
              last_actions=g["last_actions"].copy(), last_dof_vel=g["last_dof_vel"].copy(), feet_air_time=g["feet_air_time"].copy(),
              progress=g["progress"].copy(), timeout_prev=g["timeout_prev"].copy(), episode_sums=g["episode_sums"].copy(),
              eef_state=g["eef_state"].copy(), arm_commands=np.zeros((len(g["root"]), 3), np.float32), terrain_levels=np.zeros(len(g["root"]), np.int64))
    cfg = dict(rew_scales=g["rew_scales"], knee=g["knee"], feet=g["feet"], base_indices=g["base_indices"], base_body=0, allow_knee=True, hound=True,
               base_height_target=0.52, noise_scale_vec=g["noise_scale_vec"], dt=float(g["dt"]), max_len=int(g["max_len"]),
               push=int(g["common_step_counter"] + 1) % int(g["push_interval"]) == 0, default_dof_pos=g["default"][0],
               init_root=np.array([0, 0, 0.62, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], np.float32), cmd_x=[-1, 1], cmd_y=[-1, 1], cmd_yaw=[-3.14, 3.14],
               custom_origins=False, curriculum=True, terrain=None, max_episode_length_s=20.0, lin_vel_scale=2.0, ang_vel_scale=0.25, dof_pos_scale=1.0,
               dof_vel_scale=0.05, height_meas_scale=5.0,
               arm=dict(dof_noise=0.25, lower=np.full(6, -1.57, np.float32), upper=np.full(6, 1.57, np.float32)))
    draws = dict(reset=g["reset_draws"], noise=g["noise_draws"], push=g["push_draws"])
    return g, st, cfg, draws


def test_useful_hound_post_physics_matches_reference():
    g, st, cfg, draws = useful_case()
    obs, rew, reset, timeout, measured, extras = tm.terrain_post_physics(st, cfg, draws)
    assert obs.shape == (64, 204)
    assert np.array_equal(reset, g["o_reset"]) and np.array_equal(st["progress"], g["o_progress"]) and np.array_equal(timeout, g["o_timeout"])
    np.testing.assert_allclose(rew, g["o_rew"], rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(obs, g["o_obs"], rtol=1e-5, atol=2e-6)
    ods = g["o_dof_state"]
    np.testing.assert_allclose(st["dof_pos"], ods[:, :12, 0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(st["arm_q"], ods[:, 12:, 0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(st["arm_qd"], ods[:, 12:, 1], rtol=0, atol=0)
    np.testing.assert_allclose(st["root"], g["o_root"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st["commands"], g["o_commands"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st["last_actions"], g["o_last_actions"], rtol=0, atol=0)
    np.testing.assert_allclose(st["last_dof_vel"], g["o_last_dof_vel"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(st["episode_sums"], g["o_episode_sums"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(extras[:13], g["o_extras"][:13], rtol=1e-4, atol=1e-6)


def test_osc_torques_match_reference():
    g = np.load(os.path.join(GOLDEN, "useful_hound_plane.npz"))
    u = tm.osc_torques(g["osc_mm"], g["osc_j"], g["osc_dpose"], g["osc_eef_vel"], g["osc_q"], g["osc_qd"])
    np.testing.assert_allclose(u, g["osc_u"], rtol=2e-4, atol=2e-4)      # two float32 6x6 inversions in the reference
    assert np.abs(g["osc_u"]).max() > 10
