"""The numpy task-math oracle against golden vectors produced by the reference's own
@torch.jit.script functions (tests/golden/gen_golden.py).  Tolerance: 1e-5 relative (north_star);
boolean/integer masks bit-exact."""
import os

import numpy as np

from oracle import task_math as tm

RTOL, ATOL = 1e-5, 1e-6


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _check_flat(g):
    scales = {"lin_vel_xy": float(g["scale_lin"]), "ang_vel_z": float(g["scale_ang"]), "torque": float(g["scale_torque"])}
    rew, reset = tm.compute_anymal_reward(g["root"], g["commands"], g["torques"], g["contact"], g["knee"], g["progress"],
                                          scales, int(g["base"]), int(g["max_len"]))
    np.testing.assert_allclose(rew, g["rew"], rtol=RTOL, atol=1e-8)
    assert np.array_equal(reset, g["reset"])
    assert reset.any() and not reset.all()
    s = g["obs_scales"]
    grav = np.tile(np.array([[0, 0, -1]], dtype=np.float32), (g["root"].shape[0], 1))
    obs = tm.compute_anymal_observations(g["root"], g["commands"], g["dof_pos"], g["default"], g["dof_vel"], grav, g["actions"],
                                         s[0], s[1], s[2], s[3])
    assert obs.shape == g["obs"].shape and obs.dtype == np.float32
    np.testing.assert_allclose(obs, g["obs"], rtol=RTOL, atol=ATOL)


def test_anymal_reward_obs(golden_dir):
    _check_flat(_load(golden_dir, "anymal_flat.npz"))


def test_hound_reward_obs(golden_dir):
    _check_flat(_load(golden_dir, "hound_flat.npz"))


def test_cartpole_reward(golden_dir):
    g = _load(golden_dir, "cartpole.npz")
    rew, reset = tm.compute_cartpole_reward(g["pole_angle"], g["pole_vel"], g["cart_vel"], g["cart_pos"], 3.0, g["reset_buf"], g["progress"], 500.0)
    np.testing.assert_allclose(rew, g["rew"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(reset, g["reset"])


def test_jit_utils(golden_dir):
    g = _load(golden_dir, "jit_utils.npz")
    for name, val in (("quat_rotate", tm.quat_rotate(g["q"], g["v"])), ("quat_rotate_inverse", tm.quat_rotate_inverse(g["q"], g["v"])),
                      ("quat_apply", tm.quat_apply(g["q"], g["v"])), ("quat_mul", tm.quat_mul(g["q"], g["q2"])),
                      ("normalize", tm.normalize(g["v"])), ("quat_apply_yaw", tm.quat_apply_yaw(g["q"], g["v"])),
                      ("torch_rand_float", tm.torch_rand_float(np.float32(0.5), np.float32(1.5), g["rand_u"]))):
        np.testing.assert_allclose(val, g[name], rtol=RTOL, atol=ATOL, err_msg=name)
    # wrap_to_pi: C fmod semantics (SURVEY trap 4): -3.5 stays -3.5
    w = tm.wrap_to_pi(g["angles"])
    np.testing.assert_allclose(w, g["wrap_to_pi"], rtol=RTOL, atol=2e-5)
    assert abs(w[2] + 3.5) < 1e-6


def test_survey_known_answers():
    """SURVEY.md 8(c) values computed from the reference in the survey session."""
    n = 4

    def sinfill(shape, a, b):
        k = int(np.prod(shape))
        return np.sin(a * np.arange(k) + b).reshape(shape).astype(np.float32)

    root = sinfill((n, 13), .37, .1)
    root[:, 3:7] /= np.linalg.norm(root[:, 3:7], axis=1, keepdims=True)
    root[:, 7:13] *= 0.5
    cmd = sinfill((n, 3), .91, .2) * 0.5
    tq = sinfill((n, 12), .43, .6) * 40
    contact = sinfill((n * 13 * 3,), .67, .7).reshape(n, 13, 3) * np.array([0.3, 0.55, 1.2, 0.3], dtype=np.float32)[:, None, None]
    progress = np.array([0, 10, 2498, 2499])
    rew, reset = tm.compute_anymal_reward(root, cmd, tq, contact, np.array([2, 5, 8, 11]), progress,
                                          {"lin_vel_xy": 0.02, "ang_vel_z": 0.01, "torque": -5e-7}, 0, 2500)
    np.testing.assert_allclose(rew, [0.0010574, 0.0125927, 0.0142953, 0.0043275], rtol=2e-4)
    assert reset.tolist() == [False, False, True, True]
    q = np.array([[.1, .2, .3, .9]], dtype=np.float32)
    q /= np.linalg.norm(q)
    out = tm.quat_apply_yaw(np.repeat(q, 2, 0), np.array([[1, 0, 0], [0, 1, 0]], dtype=np.float32))
    np.testing.assert_allclose(out, [[.8, .6, 0], [-.6, .8, 0]], atol=1e-6)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def h(a):
        return ["%08x" % x for x in a]
    assert h(tm.philox4x32([0, 0, 0, 0], [0, 0])) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert h(tm.philox4x32([0xffffffff] * 4, [0xffffffff] * 2)) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert h(tm.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]
    u = tm.philox_uniform(42, np.arange(1000), np.zeros(1000), 27)
    assert u.shape == (1000, 27) and u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
