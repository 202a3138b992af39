"""ORACLE -- TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's per-task
reward / observation / reset math.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
Pinned: ``tests/golden/*.npz`` hold outputs of the reference's own ``@torch.jit.script`` functions
(generated in the build container by ``tests/golden/gen_golden.py``); ``tests/test_oracle_task_math.py``
checks every function here against them.

Each function cites the reference code it follows (paths relative to ``/root/reference/isaacgymenvs``).
Arithmetic runs in the dtype of the inputs (float32 = what the reference computes in).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------------------------
# utils/torch_jit_utils.py
# ----------------------------------------------------------------------------------------------
def quat_rotate(q, v):
    """utils/torch_jit_utils.py:80-90 (xyzw)."""
    q_w = q[:, 3:4]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0)
    b = np.cross(q_vec, v) * q_w * 2.0
    c = q_vec * np.sum(q_vec * v, axis=1, keepdims=True) * 2.0
    return (a + b + c).astype(v.dtype)


def quat_rotate_inverse(q, v):
    """utils/torch_jit_utils.py:93-103."""
    q_w = q[:, 3:4]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0)
    b = np.cross(q_vec, v) * q_w * 2.0
    c = q_vec * np.sum(q_vec * v, axis=1, keepdims=True) * 2.0
    return (a - b + c).astype(v.dtype)


def quat_apply(a, b):
    """utils/torch_jit_utils.py:70-77."""
    xyz = a[:, :3]
    t = np.cross(xyz, b) * 2
    return (b + a[:, 3:4] * t + np.cross(xyz, t)).astype(b.dtype)


def normalize(x, eps=1e-9):
    """utils/torch_jit_utils.py:65-67."""
    n = np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), eps)
    return (x / n).astype(x.dtype)


def quat_mul(a, b):
    """utils/torch_jit_utils.py:41-61."""
    x1, y1, z1, w1 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    x2, y2, z2, w2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return np.stack([x, y, z, w], axis=-1).astype(a.dtype)


def torch_rand_float(lower, upper, u):
    """utils/torch_jit_utils.py:215-218 with the uniform draw ``u`` in [0,1) injected."""
    return ((upper - lower) * u + lower).astype(u.dtype)


def wrap_to_pi(angles):
    """tasks/anymal_terrain.py:683-687 as TorchScript executes it: ``%`` compiles to aten::fmod
    (C semantics, sign of the dividend), SURVEY.md trap 4."""
    a = np.fmod(angles, np.asarray(2 * np.pi, dtype=angles.dtype))
    a = a - np.asarray(2 * np.pi, dtype=angles.dtype) * (a > np.pi)
    return a.astype(angles.dtype)


def quat_apply_yaw(quat, vec):
    """tasks/anymal_terrain.py:676-680."""
    q = quat.copy().reshape(-1, 4)
    q[:, :2] = 0.0
    q = normalize(q)
    return quat_apply(q, vec)


# ----------------------------------------------------------------------------------------------
# tasks/anymal.py (tasks/hound.py is the same math with other body indices)
# ----------------------------------------------------------------------------------------------
def compute_anymal_reward(root_states, commands, torques, contact_forces, knee_indices, episode_lengths,
                          rew_scales, base_index, max_episode_length):
    """tasks/anymal.py:311-351."""
    dt = root_states.dtype
    base_quat = root_states[:, 3:7]
    base_lin_vel = quat_rotate_inverse(base_quat, root_states[:, 7:10])
    base_ang_vel = quat_rotate_inverse(base_quat, root_states[:, 10:13])
    lin_vel_error = np.sum(np.square(commands[:, :2] - base_lin_vel[:, :2]), axis=1)
    ang_vel_error = np.square(commands[:, 2] - base_ang_vel[:, 2])
    rew_lin_vel_xy = np.exp(-lin_vel_error / dt.type(0.25)) * dt.type(rew_scales["lin_vel_xy"])
    rew_ang_vel_z = np.exp(-ang_vel_error / dt.type(0.25)) * dt.type(rew_scales["ang_vel_z"])
    rew_torque = np.sum(np.square(torques), axis=1) * dt.type(rew_scales["torque"])
    total = rew_lin_vel_xy + rew_ang_vel_z + rew_torque
    total = np.clip(total, 0.0, None).astype(dt)
    reset = np.linalg.norm(contact_forces[:, base_index, :], axis=1) > 1.0
    reset = reset | np.any(np.linalg.norm(contact_forces[:, knee_indices, :], axis=2) > 1.0, axis=1)
    time_out = episode_lengths >= max_episode_length - 1
    reset = reset | time_out
    return total, reset


def compute_anymal_observations(root_states, commands, dof_pos, default_dof_pos, dof_vel, gravity_vec, actions,
                                lin_vel_scale, ang_vel_scale, dof_pos_scale, dof_vel_scale):
    """tasks/anymal.py:354-386. NB gravity is projected with the *forward* rotation (:372)."""
    dt = root_states.dtype
    base_quat = root_states[:, 3:7]
    base_lin_vel = quat_rotate_inverse(base_quat, root_states[:, 7:10]) * dt.type(lin_vel_scale)
    base_ang_vel = quat_rotate_inverse(base_quat, root_states[:, 10:13]) * dt.type(ang_vel_scale)
    projected_gravity = quat_rotate(base_quat, gravity_vec)
    dof_pos_scaled = (dof_pos - default_dof_pos) * dt.type(dof_pos_scale)
    commands_scaled = commands * np.array([lin_vel_scale, lin_vel_scale, ang_vel_scale], dtype=dt)
    return np.concatenate([base_lin_vel, base_ang_vel, projected_gravity, commands_scaled, dof_pos_scaled,
                           dof_vel * dt.type(dof_vel_scale), actions], axis=-1).astype(dt)


# ----------------------------------------------------------------------------------------------
# tasks/cartpole.py
# ----------------------------------------------------------------------------------------------
def compute_cartpole_reward(pole_angle, pole_vel, cart_vel, cart_pos, reset_dist, reset_buf, progress_buf, max_episode_length):
    """tasks/cartpole.py:180-196."""
    dt = pole_angle.dtype
    reward = dt.type(1.0) - pole_angle * pole_angle - dt.type(0.01) * np.abs(cart_vel) - dt.type(0.005) * np.abs(pole_vel)
    reward = np.where(np.abs(cart_pos) > reset_dist, dt.type(-2.0), reward)
    reward = np.where(np.abs(pole_angle) > np.pi / 2, dt.type(-2.0), reward)
    reset = np.where(np.abs(cart_pos) > reset_dist, np.ones_like(reset_buf), reset_buf)
    reset = np.where(np.abs(pole_angle) > np.pi / 2, np.ones_like(reset_buf), reset)
    reset = np.where(progress_buf >= max_episode_length - 1, np.ones_like(reset_buf), reset)
    return reward.astype(dt), reset


# ----------------------------------------------------------------------------------------------
# counter-based RNG shared with the CUDA kernels (Philox4x32-10, Salmon et al. SC'11)
# ----------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(counter, key):
    """counter: (...,4) uint32, key: (...,2) uint32 -> (...,4) uint32, 10 rounds."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(key, dtype=np.uint32, copy=True)
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    for _ in range(10):
        p0 = c[..., 0].astype(np.uint64) * _PHILOX_M0
        p1 = c[..., 2].astype(np.uint64) * _PHILOX_M1
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        n0 = hi1 ^ c[..., 1] ^ k0
        n1 = lo1
        n2 = hi0 ^ c[..., 3] ^ k1
        n3 = lo0
        c = np.stack([n0, n1, n2, n3], axis=-1)
        with np.errstate(over="ignore"):
            k0 = (k0 + _PHILOX_W0).astype(np.uint32)
            k1 = (k1 + _PHILOX_W1).astype(np.uint32)
    return c


def philox_uniform(seed, env_ids, reset_count, n_draws):
    """The kernels' reset draws: block j of env e at its r-th reset = philox(counter=(e, r, j, 0),
    key=(seed_lo, seed_hi)); uniform = (x >> 8) * 2^-24 in [0,1).  Returns (len(env_ids), n_draws) f32."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    reset_count = np.asarray(reset_count, dtype=np.uint32)
    nblk = (n_draws + 3) // 4
    ctr = np.zeros((len(env_ids), nblk, 4), dtype=np.uint32)
    ctr[:, :, 0] = env_ids[:, None]
    ctr[:, :, 1] = reset_count[:, None]
    ctr[:, :, 2] = np.arange(nblk, dtype=np.uint32)[None, :]
    key = np.zeros((len(env_ids), nblk, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    out = philox4x32(ctr, key).reshape(len(env_ids), nblk * 4)[:, :n_draws]
    return ((out >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# whole post_physics_step of the flat task, in the reference's order
# ----------------------------------------------------------------------------------------------
def anymal_post_physics(state, cfg, actions, draws):
    """tasks/anymal.py:231-239 (+ reset_idx :278-304, VecTask.step tail vec_task.py:391-402).

    ``state``: dict with root (N,13), dof_pos, dof_vel (N,12), torques (N,12), contact (N,nb,3), commands (N,3),
    progress (N) int64, reset (N) int64 -- updated in place.  ``draws``: (N,27) uniforms used by envs that reset
    (12 position offsets, 12 velocities, cmd x, y, yaw -- the order of the torch_rand_float calls).
    Returns obs (N,48), obs_clamped, rew (N), timeout (N) int64.
    """
    dt = state["root"].dtype
    state["progress"] += 1
    ids = np.nonzero(state["reset"])[0]
    if len(ids) > 0:
        u = draws[ids].astype(dt)
        d0 = cfg["default_dof_pos"].astype(dt)
        state["dof_pos"][ids] = d0[None, :] * torch_rand_float(dt.type(0.5), dt.type(1.5), u[:, 0:12])
        state["dof_vel"][ids] = torch_rand_float(dt.type(-0.1), dt.type(0.1), u[:, 12:24])
        state["root"][ids] = cfg["init_root"].astype(dt)[None, :]
        for col, rng in ((0, cfg["cmd_x"]), (1, cfg["cmd_y"]), (2, cfg["cmd_yaw"])):
            state["commands"][ids, col] = torch_rand_float(dt.type(rng[0]), dt.type(rng[1]), u[:, 24 + col])
        state["progress"][ids] = 0
        state["reset"][ids] = 1
    n = state["root"].shape[0]
    grav = np.tile(np.array([[0.0, 0.0, -1.0]], dtype=dt), (n, 1))
    obs = compute_anymal_observations(state["root"], state["commands"], state["dof_pos"], np.tile(cfg["default_dof_pos"].astype(dt), (n, 1)),
                                      state["dof_vel"], grav, actions.astype(dt), cfg["lin_vel_scale"], cfg["ang_vel_scale"],
                                      cfg["dof_pos_scale"], cfg["dof_vel_scale"])
    rew, reset = compute_anymal_reward(state["root"], state["commands"], state["torques"], state["contact"], cfg["knee_bodies"],
                                       state["progress"], cfg["rew_scales"], cfg["base_body"], cfg["max_episode_length"])
    state["reset"][:] = reset.astype(np.int64)
    timeout = ((state["progress"] >= cfg["max_episode_length"] - 1) & (state["reset"] != 0)).astype(np.int64)
    obs_clamped = np.clip(obs, -dt.type(cfg["clip_obs"]), dt.type(cfg["clip_obs"]))
    return obs, obs_clamped, rew, timeout
