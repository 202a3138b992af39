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


# ----------------------------------------------------------------------------------------------
# tasks/anymal_terrain.py / tasks/Hound_terrain.py: post_physics_step of the rough-terrain tasks
# ----------------------------------------------------------------------------------------------
EPISODE_KEYS = ("lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torques", "joint_acc", "base_height", "air_time", "collision",
                "stumble", "action_rate", "hip")
REW_ORDER = ("termination", "lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torque", "joint_acc", "base_height", "air_time",
             "collision", "stumble", "action_rate", "hip")
HEIGHT_X = 0.1 * np.array([-8, -7, -6, -5, -4, -3, -2, 2, 3, 4, 5, 6, 7, 8], dtype=np.float32)
HEIGHT_Y = 0.1 * np.array([-5, -4, -3, -2, -1, 1, 2, 3, 4, 5], dtype=np.float32)


def height_points(n):
    """tasks/anymal_terrain.py:503-513: 14 x 10 grid, x-major (index = ix * 10 + iy)."""
    gx, gy = np.meshgrid(HEIGHT_X, HEIGHT_Y, indexing="ij")
    pts = np.zeros((n, gx.size, 3), dtype=np.float32)
    pts[:, :, 0] = gx.reshape(-1)
    pts[:, :, 1] = gy.reshape(-1)
    return pts


def get_heights(root, height_samples, border_size, hscale, vscale):
    """tasks/anymal_terrain.py:515-538 (trimesh branch); ``height_samples`` None = plane -> zeros."""
    n = root.shape[0]
    pts = height_points(n)
    if height_samples is None:
        return np.zeros((n, pts.shape[1]), dtype=np.float32)
    quat = np.repeat(root[:, 3:7], pts.shape[1], axis=0)
    p = quat_apply_yaw(quat, pts.reshape(-1, 3)).reshape(n, -1, 3) + root[:, None, :3]
    p = p + np.float32(border_size)
    p = (p / np.float32(hscale)).astype(np.int64)       # .long(): truncation toward zero
    px = np.clip(p[:, :, 0].reshape(-1), 0, height_samples.shape[0] - 2)
    py = np.clip(p[:, :, 1].reshape(-1), 0, height_samples.shape[1] - 2)
    h = np.minimum(height_samples[px, py], height_samples[px + 1, py + 1])
    return (h.reshape(n, -1) * np.float32(vscale)).astype(np.float32)


def terrain_post_physics(st, cfg, draws):
    """post_physics_step of AnymalTerrain (tasks/anymal_terrain.py:453-485) and HoundTerrain (Hound_terrain.py, same
    structure; termination :304-311, base-height target :347), followed by the VecTask.step tail (vec_task.py:394).

    ``st`` (updated in place): root (N,13), dof_pos, dof_vel, contact (N,nb,3), torques, commands (N,4), actions, last_actions,
    last_dof_vel, feet_air_time (N,4), progress int64, timeout_prev bool, episode_sums (13,N), and for custom origins
    terrain_levels, terrain_types (int64), env_origins (N,3).
    ``cfg``: rew_scales (14, REW_ORDER, already x dt), scales, knee/feet/base indices, allow_knee, hound (bool), base_height_target,
    noise_scale_vec (188) or None, dt, max_len, push (bool: this step pushes), default_dof_pos, init_root, cmd ranges, custom_origins,
    curriculum, terrain (height_samples, border_size, hscale, vscale, env_length, env_rows, terrain_origins), max_episode_length_s.
    ``draws``: reset (N,29), noise (N,188), push (N,2) uniforms in [0,1).
    Returns obs (N,188), rew, reset int64, timeout int64, measured heights, extras (14: 13 episode means + terrain level) or None.
    """
    f = np.float32
    n = st["root"].shape[0]
    rs = dict(zip(REW_ORDER, cfg["rew_scales"].astype(f)))
    st["progress"] += 1
    if cfg["push"]:      # :437-439  root lin vel x/y <- U(-1, 1) for all envs
        st["root"][:, 7:9] = torch_rand_float(f(-1.0), f(1.0), draws["push"].astype(f))
    root = st["root"]
    quat = root[:, 3:7].copy()
    base_lin = quat_rotate_inverse(quat, root[:, 7:10])
    base_ang = quat_rotate_inverse(quat, root[:, 10:13])
    grav = np.tile(np.array([[0, 0, -1]], f), (n, 1))
    pg = quat_rotate_inverse(quat, grav)
    fwd = quat_apply(quat, np.tile(np.array([[1, 0, 0]], f), (n, 1)))
    heading = np.arctan2(fwd[:, 1], fwd[:, 0]).astype(f)
    st["commands"][:, 2] = np.clip(f(0.5) * wrap_to_pi(st["commands"][:, 3] - heading), -1.0, 1.0)      # :469-471
    cf = st["contact"]
    # check_termination
    if cfg.get("hound", False):       # Hound_terrain.py:304-311
        reset = np.linalg.norm(cf[:, cfg["base_body"], :], axis=1) > 1.0
        reset = reset | np.any(np.linalg.norm(cf[:, cfg["knee"], :], axis=2) > 1.0, axis=1)
        reset = reset | np.any(np.linalg.norm(cf[:, cfg["base_indices"], :], axis=2) > 1.0, axis=1)
        reset = reset | (st["progress"] >= cfg["max_len"] - 1)
    else:                             # anymal_terrain.py:294-300
        reset = np.linalg.norm(cf[:, cfg["base_body"], :], axis=1) > 1.0
        if not cfg["allow_knee"]:
            reset = reset | np.any(np.linalg.norm(cf[:, cfg["knee"], :], axis=2) > 1.0, axis=1)
        reset = np.where(st["progress"] >= cfg["max_len"] - 1, True, reset)
    # compute_reward :315-382
    cmd = st["commands"]
    lin_err = np.sum(np.square(cmd[:, :2] - base_lin[:, :2]), axis=1)
    ang_err = np.square(cmd[:, 2] - base_ang[:, 2])
    r_lin_xy = np.exp(-lin_err / f(0.25)) * rs["lin_vel_xy"]
    r_ang_z = np.exp(-ang_err / f(0.25)) * rs["ang_vel_z"]
    r_lin_z = np.square(base_lin[:, 2]) * rs["lin_vel_z"]
    r_ang_xy = np.sum(np.square(base_ang[:, :2]), axis=1) * rs["ang_vel_xy"]
    r_orient = np.sum(np.square(pg[:, :2]), axis=1) * rs["orient"]
    r_height = np.square(root[:, 2] - f(cfg["base_height_target"])) * rs["base_height"]
    r_torque = np.sum(np.square(st["torques"]), axis=1) * rs["torque"]
    r_jacc = np.sum(np.square(st["last_dof_vel"] - st["dof_vel"]), axis=1) * rs["joint_acc"]
    knee_contact = np.linalg.norm(cf[:, cfg["knee"], :], axis=2) > 1.0
    r_coll = np.sum(knee_contact, axis=1).astype(f) * rs["collision"]
    if cfg.get("arm"):      # useful_hound.py:524-525: contacts on the `baseName` bodies are penalised as well
        base_contact = np.linalg.norm(cf[:, cfg["base_indices"], :], axis=2) > 1.0
        r_coll = r_coll + np.sum(base_contact, axis=1).astype(f) * rs["collision"]
    stumble = (np.linalg.norm(cf[:, cfg["feet"], :2], axis=2) > 5.0) & (np.abs(cf[:, cfg["feet"], 2]) < 1.0)
    r_stumble = np.sum(stumble, axis=1).astype(f) * rs["stumble"]
    r_arate = np.sum(np.square(st["last_actions"] - st["actions"]), axis=1) * rs["action_rate"]
    contact = cf[:, cfg["feet"], 2] > 1.0
    first = (st["feet_air_time"] > 0.0) & contact
    st["feet_air_time"] += f(cfg["dt"])
    r_air = np.sum((st["feet_air_time"] - f(0.5)) * first, axis=1).astype(f) * rs["air_time"]
    r_air = r_air * (np.linalg.norm(cmd[:, :2], axis=1) > 0.1)
    st["feet_air_time"] *= ~contact
    hips = [0, 3, 6, 9]
    r_hip = np.sum(np.abs(st["dof_pos"][:, hips] - cfg["default_dof_pos"][hips][None]), axis=1) * rs["hip"]
    rew = r_lin_xy + r_ang_z + r_lin_z + r_ang_xy + r_orient + r_height + r_torque + r_jacc + r_coll + r_arate + r_air + r_hip + r_stumble
    rew = np.clip(rew, 0.0, None).astype(f)
    rew = rew + rs["termination"] * reset * ~st["timeout_prev"]
    terms = dict(lin_vel_xy=r_lin_xy, ang_vel_z=r_ang_z, lin_vel_z=r_lin_z, ang_vel_xy=r_ang_xy, orient=r_orient, torques=r_torque,
                 joint_acc=r_jacc, collision=r_coll, stumble=r_stumble, action_rate=r_arate, air_time=r_air, base_height=r_height, hip=r_hip)
    for i, k in enumerate(EPISODE_KEYS):
        st["episode_sums"][i] += terms[k].astype(f)
    # reset_idx :384-425
    ids = np.nonzero(reset)[0]
    extras = None
    if len(ids) > 0:
        u = draws["reset"][ids].astype(f)
        d0 = cfg["default_dof_pos"].astype(f)
        st["dof_pos"][ids] = d0[None] * torch_rand_float(f(0.5), f(1.5), u[:, 0:12])
        st["dof_vel"][ids] = torch_rand_float(f(-0.1), f(0.1), u[:, 12:24])
        col = 24
        if cfg["custom_origins"]:
            t = cfg["terrain"]
            if cfg["curriculum"]:            # update_terrain_level :427-435 (norm WITHOUT dim: one scalar for all resetting envs, quirk Q10)
                dist = np.linalg.norm(st["root"][ids, :2] - st["env_origins"][ids, :2], axis=1)
                thr = f(np.sqrt(np.sum(np.square(st["commands"][ids, :2]).astype(f)))) * f(cfg["max_episode_length_s"]) * f(0.25)
                st["terrain_levels"][ids] -= 1 * (dist < thr)
                st["terrain_levels"][ids] += 1 * (dist > t["env_length"] / 2)
                st["terrain_levels"][ids] = np.clip(st["terrain_levels"][ids], 0, None) % t["env_rows"]
                st["env_origins"][ids] = t["terrain_origins"][st["terrain_levels"][ids], st["terrain_types"][ids]]
            st["root"][ids] = cfg["init_root"].astype(f)[None]
            st["root"][ids, :3] += st["env_origins"][ids]
            st["root"][ids, :2] += torch_rand_float(f(-0.5), f(0.5), u[:, 24:26])
            col = 26
        else:
            st["root"][ids] = cfg["init_root"].astype(f)[None]
        if cfg.get("arm"):      # useful_hound.py:594-602: arm joints <- clamp(default + noise * 2 (u - 0.5)), zero velocity
            a = cfg["arm"]
            ua = u[:, col:col + 6]
            st["arm_q"][ids] = np.clip(f(0.0) + f(a["dof_noise"]) * f(2.0) * (ua - f(0.5)), a["lower"].astype(f)[None], a["upper"].astype(f)[None])
            st["arm_qd"][ids] = 0.0
            col += 6
        st["commands"][ids, 0] = torch_rand_float(f(cfg["cmd_x"][0]), f(cfg["cmd_x"][1]), u[:, col])
        st["commands"][ids, 1] = torch_rand_float(f(cfg["cmd_y"][0]), f(cfg["cmd_y"][1]), u[:, col + 1])
        st["commands"][ids, 3] = torch_rand_float(f(cfg["cmd_yaw"][0]), f(cfg["cmd_yaw"][1]), u[:, col + 2])
        st["commands"][ids] *= (np.linalg.norm(st["commands"][ids, :2], axis=1) > 0.25)[:, None]
        st["last_actions"][ids] = 0.0
        st["last_dof_vel"][ids] = 0.0
        st["feet_air_time"][ids] = 0.0
        st["progress"][ids] = 0
        extras = np.zeros(14, f)
        for i in range(13):
            extras[i] = np.mean(st["episode_sums"][i][ids]) / f(cfg["max_episode_length_s"])
            st["episode_sums"][i][ids] = 0.0
        extras[13] = np.mean(st["terrain_levels"].astype(f)) if "terrain_levels" in st else 0.0
    # compute_observations :302-313 (base velocities / projected gravity are the pre-reset values; heights, DOFs, commands post-reset)
    t = cfg.get("terrain")
    measured = get_heights(st["root"], t["height_samples"] if t else None, t["border_size"] if t else 0, t["hscale"] if t else 1, t["vscale"] if t else 1)
    heights = np.clip(st["root"][:, 2:3] - f(0.5) - measured, -1, 1.0) * f(cfg["height_meas_scale"])
    obs = np.concatenate([base_lin * f(cfg["lin_vel_scale"]), base_ang * f(cfg["ang_vel_scale"]), pg,
                          st["commands"][:, :3] * np.array([cfg["lin_vel_scale"], cfg["lin_vel_scale"], cfg["ang_vel_scale"]], f),
                          st["dof_pos"] * f(cfg["dof_pos_scale"]), st["dof_vel"] * f(cfg["dof_vel_scale"]), heights, st["actions"]], axis=-1).astype(f)
    if cfg.get("arm"):      # useful_hound.py:482-497: + end-effector position, orientation (never-refreshed tensor, quirk Q12), arm command
        obs = np.concatenate([obs, st["eef_state"][:, :3], st["eef_state"][:, 3:7], st["arm_commands"]], axis=-1).astype(f)
    if cfg.get("noise_scale_vec") is not None:
        obs = obs + (f(2) * draws["noise"].astype(f) - f(1)) * cfg["noise_scale_vec"].astype(f)[None]
    st["last_actions"][:] = st["actions"]
    st["last_dof_vel"][:] = st["dof_vel"]
    timeout = ((st["progress"] >= cfg["max_len"] - 1) & reset).astype(np.int64)
    return obs.astype(f), rew.astype(f), reset.astype(np.int64), timeout, measured, extras


def osc_torques(mm, j_eef, dpose, eef_vel, q, qd, kp=150.0, kp_null=10.0, effort=1000.0, exact=False, default_q=0.0):
    """tasks/useful_hound.py:660-691 / tasks/hound_arm.py:462-493 / tasks/manipulator.py:534-560 (_compute_osc_torques): operational-space
    control of the k-DOF arm (k = 6, or 7 for the Franka).  ``mm`` (N,k,k) arm block of the mass matrix, ``j_eef`` (N,6,k) the Jacobian
    slice the task takes, ``dpose`` (N,6), ``eef_vel`` (N,6), ``q``/``qd`` (N,k), ``default_q`` the null-space posture (zeros for the
    hound's arm, ``franka_default_dof_pos`` for the Manipulator).  ``exact=False`` mimics the reference's float32 evaluation (float32 products around accurately inverted matrices);
    ``exact=True`` evaluates the whole law in float64 -- what the kernels do, because J M^-1 J^T is too ill-conditioned for float32 on
    these arms (see tests/kernel_checks.check_houndarm_step)."""
    f = np.float64 if exact else np.float32
    mm, j_eef, dpose, eef_vel, q, qd = (np.asarray(x).astype(f) for x in (mm, j_eef, dpose, eef_vel, q, qd))
    kp_v = np.full(6, kp, f)
    kd_v = (f(2) * np.sqrt(kp_v)).astype(f)
    k = mm.shape[-1]
    kpn = np.full(k, kp_null, f)
    kdn = (f(2) * np.sqrt(kpn)).astype(f)
    mm_inv = np.linalg.inv(mm.astype(np.float64)).astype(f)
    jt = np.transpose(j_eef, (0, 2, 1))
    m_eef_inv = j_eef @ mm_inv @ jt
    m_eef = np.linalg.inv(m_eef_inv.astype(np.float64)).astype(f)
    u = jt @ m_eef @ (kp_v * dpose - kd_v * eef_vel)[..., None]
    j_eef_inv = m_eef @ j_eef @ mm_inv
    u_null = kdn * -qd + kpn * (np.mod(np.asarray(default_q).astype(f) - q + f(np.pi), f(2 * np.pi)) - f(np.pi))      # python-style remainder (eager torch %)
    u_null = mm @ u_null[..., None]
    u = u + (np.eye(k, dtype=f)[None] - jt @ j_eef_inv) @ u_null
    return np.clip(u[..., 0], -effort, effort).astype(np.float32)
