#!/usr/bin/env python
"""Generate golden vectors from the reference's own @torch.jit.script task functions.

Runs ONLY in the build container (needs /root/reference, imported through ref_loader's stubs).
Inputs are RNG-free (sin-based), so the fixtures do not depend on a torch RNG version.  Outputs are
committed as tests/golden/*.npz; tests compare the numpy oracle and the CUDA kernels against them.

    python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def sinfill(shape, a, b, scale=1.0):
    n = int(np.prod(shape))
    x = np.sin(a * np.arange(n, dtype=np.float64) + b) * scale
    return torch.tensor(x.reshape(shape), dtype=torch.float32)


def make_root(n, a=0.37, b=0.1, z0=0.55):
    root = sinfill((n, 13), a, b)
    root[:, 3:7] = root[:, 3:7] / root[:, 3:7].norm(dim=1, keepdim=True)
    root[:, 7:13] *= 1.5
    root[:, 2] = z0 + 0.1 * root[:, 2]
    return root


def gen_anymal(task_mod, prefix, n_bodies, knee, base, out):
    n = 64
    root = make_root(n)
    commands = sinfill((n, 3), 0.91, 0.2) * torch.tensor([2.0, 1.0, 1.0])
    torques = sinfill((n, 12), 0.43, 0.6, 40.0)
    contact = sinfill((n, n_bodies, 3), 0.67, 0.7)
    # per-env magnitude so that some envs sit below, near and above the 1 N threshold
    mag = torch.tensor([0.3, 0.55, 0.57, 0.58, 1.2, 0.3, 0.9, 0.05])[torch.arange(n) % 8]
    contact = contact * mag[:, None, None]
    progress = torch.tensor([0, 10, 2497, 2498, 2499, 2500, 1, 300], dtype=torch.int64)[torch.arange(n) % 8]
    progress = torch.roll(progress, 3)  # decorrelate from the contact magnitudes
    knee_idx = torch.tensor(knee, dtype=torch.int64)
    scales = {"lin_vel_xy": 1.0 * 0.02, "ang_vel_z": 0.5 * 0.02, "torque": -0.000025 * 0.02}
    rew_fn = getattr(task_mod, f"compute_{prefix}_reward")
    obs_fn = getattr(task_mod, f"compute_{prefix}_observations")
    rew, reset = rew_fn(root, commands, torques, contact, knee_idx, progress, scales, base, 2500)
    dof_pos = sinfill((n, 12), 0.29, 0.3, 0.8)
    default = sinfill((1, 12), 1.3, 0.5, 0.6).repeat(n, 1)
    dof_vel = sinfill((n, 12), 0.53, 0.9, 8.0)
    grav = torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1)
    actions = sinfill((n, 12), 0.77, 1.1)
    obs = obs_fn(root, commands, dof_pos, default, dof_vel, grav, actions, 2.0, 0.25, 1.0, 0.05)
    np.savez(out, root=root.numpy(), commands=commands.numpy(), torques=torques.numpy(), contact=contact.numpy(),
             progress=progress.numpy(), knee=np.array(knee), base=np.array(base), max_len=np.array(2500),
             scale_lin=np.float32(scales["lin_vel_xy"]), scale_ang=np.float32(scales["ang_vel_z"]), scale_torque=np.float32(scales["torque"]),
             rew=rew.numpy(), reset=reset.numpy(), dof_pos=dof_pos.numpy(), default=default.numpy(), dof_vel=dof_vel.numpy(),
             actions=actions.numpy(), obs=obs.numpy(), obs_scales=np.array([2.0, 0.25, 1.0, 0.05], dtype=np.float32))
    print(out, "reset count", int(reset.sum()), "rew range", float(rew.min()), float(rew.max()))


def gen_cartpole(out):
    mod = ref_loader.load("tasks.cartpole")
    n = 64
    ang = sinfill((n,), 0.7, 0.1, 2.0)
    pv = sinfill((n,), 0.31, 0.4, 6.0)
    cv = sinfill((n,), 0.57, 0.2, 3.0)
    cp = sinfill((n,), 0.23, 0.8, 3.5)
    reset_buf = torch.zeros(n, dtype=torch.int64)
    reset_buf[::7] = 1
    progress = (torch.arange(n, dtype=torch.int64) * 9) % 503
    rew, reset = mod.compute_cartpole_reward(ang, pv, cv, cp, 3.0, reset_buf, progress, 500.0)
    np.savez(out, pole_angle=ang.numpy(), pole_vel=pv.numpy(), cart_vel=cv.numpy(), cart_pos=cp.numpy(), reset_buf=reset_buf.numpy(),
             progress=progress.numpy(), rew=rew.numpy(), reset=reset.numpy())
    print(out, "reset count", int(reset.sum()))


def gen_utils(out):
    tj = ref_loader.load("utils.torch_jit_utils")
    at = ref_loader.load("tasks.anymal_terrain")
    n = 32
    q = sinfill((n, 4), 0.61, 0.3)
    q = q / q.norm(dim=1, keepdim=True)
    q2 = sinfill((n, 4), 0.47, 1.3)
    q2 = q2 / q2.norm(dim=1, keepdim=True)
    v = sinfill((n, 3), 0.83, 0.5, 2.0)
    ang = torch.tensor([0, 3.5, -3.5, 7, -7, -0.1, 3.1415927, -3.1415927, 6.2831855, 100.0, -100.0, 1e-6])
    np.savez(out, q=q.numpy(), q2=q2.numpy(), v=v.numpy(),
             quat_rotate=tj.quat_rotate(q, v).numpy(), quat_rotate_inverse=tj.quat_rotate_inverse(q, v).numpy(),
             quat_apply=tj.quat_apply(q, v).numpy(), quat_mul=tj.quat_mul(q, q2).numpy(), normalize=tj.normalize(v).numpy(),
             quat_apply_yaw=at.quat_apply_yaw(q.clone(), v).numpy(), angles=ang.numpy(), wrap_to_pi=at.wrap_to_pi(ang.clone()).numpy(),
             rand_u=sinfill((n, 3), 0.2, 0.1).abs().numpy(),
             torch_rand_float=((1.5 - 0.5) * sinfill((n, 3), 0.2, 0.1).abs() + 0.5).numpy())
    print(out)


if __name__ == "__main__" and "--dr-only" not in sys.argv and "--houndarm-only" not in sys.argv and "--manipulator-only" not in sys.argv:
    torch.set_num_threads(1)
    gen_anymal(ref_loader.load("tasks.anymal"), "anymal", 13, [2, 5, 8, 11], 0, os.path.join(HERE, "anymal_flat.npz"))
    gen_anymal(ref_loader.load("tasks.hound"), "hound", 17, [2, 6, 10, 14], 0, os.path.join(HERE, "hound_flat.npz"))
    gen_cartpole(os.path.join(HERE, "cartpole.npz"))
    gen_utils(os.path.join(HERE, "jit_utils.npz"))
    if "--no-terrain" not in sys.argv:
        globals()["_run_terrain_after"] = True


# ------------------------------------------------------------------------------------------------------------------
# terrain tasks: the reference keeps this math in eager methods of the task class (tasks/anymal_terrain.py:294-485,
# Hound_terrain.py same lines).  They are executed here unmodified, bound to an attribute bag that stands in for the
# task object; torch's RNG entry points are replaced by tables so that the fixtures are RNG-free and the kernels can
# be fed the very same draws.
# ------------------------------------------------------------------------------------------------------------------
import types  # noqa: E402


class _NullGym:
    def __getattr__(self, name):
        return lambda *a, **k: True


def _terrain_bag(mod, cls_name, n, nb, terrain_kind, seed_phase, knee, feet, base_indices=None):
    ref_cls = getattr(mod, cls_name)

    class Bag:
        def __getattr__(self, name):
            f = getattr(ref_cls, name)
            return types.MethodType(f, self)

        def reset_idx(self, env_ids):
            self._cur_ids = env_ids
            self._draw_col = 0
            ref_cls.reset_idx(self, env_ids)

        def push_robots(self):
            self._cur_ids = torch.arange(self.num_envs)
            self._draw_col = None
            ref_cls.push_robots(self)

    b = Bag()
    dt = 4 * 0.005
    b.num_envs, b.num_dof, b.num_actions, b.num_bodies, b.device = n, 12, 12, nb, "cpu"
    b.gym, b.sim, b.viewer, b.enable_viewer_sync, b.debug_viz = _NullGym(), None, None, False, False
    b.dt, b.max_episode_length_s = dt, 20
    b.max_episode_length = int(20 / dt + 0.5)
    b.push_interval = int(15 / dt + 0.5)
    b.allow_knee_contacts = False if cls_name == "AnymalTerrain" else True
    b.lin_vel_scale, b.ang_vel_scale, b.dof_pos_scale, b.dof_vel_scale, b.height_meas_scale, b.action_scale = 2.0, 0.25, 1.0, 0.05, 5.0, 0.5
    raw = dict(termination=-1.0, lin_vel_xy=1.0, lin_vel_z=-4.0, ang_vel_z=0.5, ang_vel_xy=-0.05, orient=-0.2, torque=-0.00002, joint_acc=-0.0005,
               base_height=-0.5, air_time=1.0, collision=-0.25, stumble=-0.1, action_rate=-0.01, hip=-0.05)
    b.rew_scales = {k: v * dt for k, v in raw.items()}
    b.command_x_range, b.command_y_range, b.command_yaw_range = [-1.0, 1.0], [-1.0, 1.0], [-3.14, 3.14]
    b.base_init_state = torch.tensor([0, 0, 0.62, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=torch.float32)
    b.cfg = {"env": {"terrain": {"terrainType": terrain_kind}}}
    b.custom_origins = terrain_kind == "trimesh"
    b.curriculum, b.init_done = True, True
    b.add_noise = True
    b.knee_indices, b.feet_indices, b.base_index = torch.tensor(knee), torch.tensor(feet), 0
    if base_indices is not None:
        b.base_indices = torch.tensor(base_indices)
    b.extras = {}
    b.terrain_levels = torch.zeros(n, dtype=torch.int64)     # created for every terrain type (anymal_terrain.py:258)
    b.terrain_types = torch.zeros(n, dtype=torch.int64)
    b.env_origins = torch.zeros(n, 3)
    ph = seed_phase
    b.root_states = make_root(n, 0.37 + ph, 0.1, z0=0.55)
    b.root_states[:, 7:13] *= 0.3
    b.dof_state = torch.zeros(n * 12, 2)
    b.dof_pos = sinfill((n, 12), 0.29, 0.3 + ph, 0.8)
    b.dof_vel = sinfill((n, 12), 0.53, 0.9 + ph, 2.0)
    contact = sinfill((n, nb, 3), 0.67, 0.7 + ph)
    # most envs: small forces on base/knees (below the 1 N thresholds); every 8th env exceeds them
    mag = torch.tensor([0.1, 0.2, 0.3, 0.55, 0.4, 3.0, 0.45, 0.05])[torch.arange(n) % 8]
    b.contact_forces = contact * mag[:, None, None]
    # feet: realistic loads so that the air-time / stumble logic sees contacts, lift-offs and side hits
    ft = torch.tensor(feet)
    b.contact_forces[:, ft, 2] = sinfill((n, len(feet)), 0.77, 0.2 + ph, 60.0).clamp(min=0.0)
    b.contact_forces[:, ft, :2] = sinfill((n, len(feet), 2), 0.41, 0.5 + ph, 7.0)
    b.torques = sinfill((n, 12), 0.43, 0.6 + ph, 12.0)
    b.commands = sinfill((n, 4), 0.91, 0.2 + ph) * torch.tensor([1.0, 1.0, 1.0, 3.0])
    b.commands[::5, :2] *= 0.05
    b.commands_scale = torch.tensor([2.0, 2.0, 0.25])
    b.actions = sinfill((n, 12), 0.77, 1.1 + ph)
    b.last_actions = b.actions + sinfill((n, 12), 0.71, 0.4 + ph, 0.3)
    b.last_dof_vel = b.dof_vel + sinfill((n, 12), 0.59, 0.8 + ph, 0.8)
    b.feet_air_time = (sinfill((n, 4), 0.83, 0.3 + ph) * 0.4).clamp(min=0.0)
    b.default_dof_pos = sinfill((1, 12), 1.3, 0.5, 0.6).repeat(n, 1)
    b.gravity_vec = torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1)
    b.forward_vec = torch.tensor([[1.0, 0.0, 0.0]]).repeat(n, 1)
    b.progress_buf = torch.roll(torch.tensor([0, 10, 997, 998, 999, 1000, 1, 300], dtype=torch.int64)[torch.arange(n) % 8], 3)
    b.randomize_buf = torch.zeros(n, dtype=torch.int64)
    b.timeout_buf = (torch.arange(n) % 3 == 0)
    b.reset_buf = torch.zeros(n, dtype=torch.int64)
    b.obs_buf = torch.zeros(n, 188)
    b.rew_buf = torch.zeros(n)
    keys = ["lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torques", "joint_acc", "base_height", "air_time", "collision",
            "stumble", "action_rate", "hip"]
    b.episode_sums = {k: sinfill((n,), 0.31 + 0.01 * i, 0.2 + ph) for i, k in enumerate(keys)}
    b.height_points = ref_cls.init_height_points(b)
    b.measured_heights = None
    b.noise_scale_vec = torch.zeros(188)
    nl = 1.0
    b.noise_scale_vec[:3] = 0.1 * nl * 2.0
    b.noise_scale_vec[3:6] = 0.2 * nl * 0.25
    b.noise_scale_vec[6:9] = 0.05 * nl
    b.noise_scale_vec[12:24] = 0.01 * nl * 1.0
    b.noise_scale_vec[24:36] = 1.5 * nl * 0.05
    b.noise_scale_vec[36:176] = 0.06 * nl * 5.0
    return b, keys


def gen_terrain(cls_name, modname, out, nb, knee, feet, terrain_kind, push, base_indices=None):
    mod = ref_loader.load(modname)
    n = 64
    b, keys = _terrain_bag(mod, cls_name, n, nb, terrain_kind, 0.0 if terrain_kind == "plane" else 0.37, knee, feet, base_indices)
    if terrain_kind == "trimesh":
        sys.path.insert(0, os.path.join(HERE, "..", ".."))
        from isaacgymenv_b200.terrain import Terrain
        tcfg = dict(terrainType="trimesh", curriculum=True, mapLength=8.0, mapWidth=8.0, numLevels=3, numTerrains=5,
                    terrainProportions=[0.1, 0.1, 0.35, 0.25, 0.2], slopeTreshold=0.5)
        t = Terrain(tcfg, n, seed=7)
        b.terrain = types.SimpleNamespace(border_size=t.border_size, horizontal_scale=t.horizontal_scale, vertical_scale=t.vertical_scale,
                                          env_length=t.env_length, env_rows=t.env_rows)
        b.height_samples = torch.tensor(t.heightsamples).view(t.tot_rows, t.tot_cols)
        b.terrain_origins = torch.from_numpy(t.env_origins).to(torch.float)
        b.terrain_levels = torch.arange(n) % 3
        b.terrain_types = (torch.arange(n) * 7) % 5
        b.env_origins = b.terrain_origins[b.terrain_levels, b.terrain_types].clone()
        b.root_states[:, :2] = b.env_origins[:, :2] + sinfill((n, 2), 0.47, 0.6, 5.0)
        b.root_states[:, 2] += b.env_origins[:, 2]
    b.common_step_counter = (b.push_interval - 1) if push else 5
    csc_before = b.common_step_counter
    # RNG tables
    reset_draws = sinfill((n, 29), 0.173, 0.05).abs() * 0.999
    noise_draws = sinfill((n, 188), 0.0917, 0.33).abs() * 0.999
    push_draws = sinfill((n, 2), 0.61, 0.21).abs() * 0.999

    def fake_rand_float(lower, upper, shape, device):
        ids = b._cur_ids
        if b._draw_col is None:
            u = push_draws[ids]
        else:
            c0 = b._draw_col
            u = reset_draws[ids][:, c0:c0 + shape[1]]
            b._draw_col += shape[1]
        assert tuple(u.shape) == tuple(shape), (u.shape, shape)
        return (upper - lower) * u + lower

    mod.torch_rand_float = fake_rand_float
    real_rand_like = torch.rand_like
    torch.rand_like = lambda t: noise_draws.clone()
    inputs = dict(root=b.root_states.clone(), dof_pos=b.dof_pos.clone(), dof_vel=b.dof_vel.clone(), contact=b.contact_forces.clone(),
                  torques=b.torques.clone(), commands=b.commands.clone(), actions=b.actions.clone(), last_actions=b.last_actions.clone(),
                  last_dof_vel=b.last_dof_vel.clone(), feet_air_time=b.feet_air_time.clone(), progress=b.progress_buf.clone(),
                  timeout_prev=b.timeout_buf.clone(), default=b.default_dof_pos.clone(),
                  episode_sums=torch.stack([b.episode_sums[k] for k in keys]).clone())
    if terrain_kind == "trimesh":
        inputs.update(height_samples=b.height_samples.clone(), terrain_origins=b.terrain_origins.clone(), terrain_levels=b.terrain_levels.clone(),
                      terrain_types=b.terrain_types.clone(), env_origins=b.env_origins.clone())
    try:
        getattr(mod, cls_name).post_physics_step(b)
    finally:
        torch.rand_like = real_rand_like
    # VecTask.step tail (vec_task.py:394)
    timeout = (b.progress_buf >= b.max_episode_length - 1) & (b.reset_buf != 0)
    outs = dict(o_root=b.root_states, o_dof_pos=b.dof_pos, o_dof_vel=b.dof_vel, o_commands=b.commands, o_obs=b.obs_buf, o_rew=b.rew_buf,
                o_reset=b.reset_buf.to(torch.int64), o_progress=b.progress_buf, o_timeout=timeout.to(torch.int64), o_last_actions=b.last_actions,
                o_last_dof_vel=b.last_dof_vel, o_feet_air_time=b.feet_air_time, o_episode_sums=torch.stack([b.episode_sums[k] for k in keys]),
                o_measured_heights=b.measured_heights,
                o_extras=torch.tensor([float(b.extras["episode"]["rew_" + k]) for k in keys] + [float(b.extras["episode"]["terrain_level"])])
                if "episode" in b.extras else torch.zeros(14))
    if terrain_kind == "trimesh":
        outs.update(o_terrain_levels=b.terrain_levels, o_env_origins=b.env_origins)
    meta = dict(knee=np.array(knee), feet=np.array(feet), base_indices=np.array(base_indices if base_indices is not None else []),
                rew_scales=np.array([b.rew_scales[k] for k in ["termination", "lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torque",
                                                                 "joint_acc", "base_height", "air_time", "collision", "stumble", "action_rate", "hip"]], dtype=np.float32),
                noise_scale_vec=b.noise_scale_vec.numpy(), reset_draws=reset_draws.numpy(), noise_draws=noise_draws.numpy(), push_draws=push_draws.numpy(),
                common_step_counter=np.array(csc_before), push_interval=np.array(b.push_interval), max_len=np.array(b.max_episode_length),
                allow_knee=np.array(b.allow_knee_contacts), dt=np.float32(b.dt), custom_origins=np.array(b.custom_origins),
                border_size=np.float32(20.0), hscale=np.float32(0.1), vscale=np.float32(0.005), env_length=np.float32(8.0), env_rows=np.array(3))
    np.savez_compressed(out, **{k: v.numpy() for k, v in inputs.items()}, **{k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in outs.items()}, **meta)
    print(out, "resets", int(b.reset_buf.sum()), "rew range", float(b.rew_buf.min()), float(b.rew_buf.max()))


def gen_all_terrain():
    gen_terrain("AnymalTerrain", "tasks.anymal_terrain", os.path.join(HERE, "anymal_terrain_plane.npz"), 13, [2, 5, 8, 11], [3, 6, 9, 12], "plane", push=True)
    gen_terrain("AnymalTerrain", "tasks.anymal_terrain", os.path.join(HERE, "anymal_terrain_trimesh.npz"), 13, [2, 5, 8, 11], [3, 6, 9, 12], "trimesh", push=False)
    gen_terrain("HoundTerrain", "tasks.Hound_terrain", os.path.join(HERE, "hound_terrain_plane.npz"), 17, [2, 6, 10, 14], [4, 8, 12, 16], "plane", push=False,
                base_indices=[1, 5, 9, 13])


if __name__ == "__main__" and ("--terrain" in sys.argv or globals().get("_run_terrain_after")):
    gen_all_terrain()


# ------------------------------------------------------------------------------------------------------------------
# UsefulHound (hound + arm): post_physics_step and the operational-space torque law, executed from the reference
# (tasks/useful_hound.py:660-691, :727-760) on an attribute bag.
# ------------------------------------------------------------------------------------------------------------------
def gen_useful_hound(out):
    mod = ref_loader.load("tasks.useful_hound")
    n, nb = 64, 24
    knee, feet, base_indices = [2, 6, 10, 14], [4, 8, 12, 16], [1, 5, 9, 13]
    b, keys = _terrain_bag(mod, "UsefulHound", n, nb, "plane", 0.11, knee, feet, base_indices)
    b.num_actions = 18
    b.total_num_dof, b.hound_num_dof, b.arm_num_dof = 18, 12, 6
    b.dof_state = torch.zeros(n, 18, 2)
    b.dof_state[:, :12, 0] = b.dof_pos
    b.dof_state[:, :12, 1] = b.dof_vel
    b.dof_state[:, 12:, 0] = sinfill((n, 6), 0.33, 0.7, 1.2)
    b.dof_state[:, 12:, 1] = sinfill((n, 6), 0.47, 0.2, 1.5)
    b.hound_dof_pos, b.hound_dof_vel = b.dof_state[:, 0:12, 0], b.dof_state[:, 0:12, 1]
    b._q, b._qd = b.dof_state[:, 12:, 0], b.dof_state[:, 12:, 1]
    b.last_hound_dof_vel = b.last_dof_vel.clone()
    b.hound_default_dof_pos = b.default_dof_pos.clone()
    b.houndarm_default_dof_pos = torch.zeros(6)
    b.houndarm_dof_noise = 0.25
    b.houndarm_dof_lower_limits = torch.full((6,), -1.57)
    b.houndarm_dof_upper_limits = torch.full((6,), 1.57)
    b._houndarm_effort_limits = torch.full((6,), 1000.0)
    b._pos_control = torch.zeros(n, 6)
    b._effort_control = torch.zeros(n, 6)
    b.arm_commands = torch.zeros(n, 3)
    b.arm_kp = torch.full((6,), 150.0)
    b.arm_kd = 2 * torch.sqrt(b.arm_kp)
    b.arm_kp_null = torch.full((6,), 10.0)
    b.arm_kd_null = 2 * torch.sqrt(b.arm_kp_null)
    b.actions = sinfill((n, 18), 0.77, 1.1)
    b.last_actions = b.actions + sinfill((n, 18), 0.71, 0.4, 0.3)
    b.torques = sinfill((n, 18), 0.43, 0.6, 12.0)
    b.obs_buf = torch.zeros(n, 204)
    nv = torch.zeros(204)
    nv[:188] = b.noise_scale_vec
    b.noise_scale_vec = nv
    b._eef_state = sinfill((n, 13), 0.21, 0.9)          # never-refreshed rigid-body rows (quirk Q12): arbitrary but fixed here
    b.episode_sums = {k: sinfill((n,), 0.31 + 0.01 * i, 0.2) for i, k in enumerate(keys)}
    # --- OSC torque law ---
    mm = sinfill((n, 6, 6), 0.37, 0.3, 0.2)
    b._mm = mm @ mm.transpose(1, 2) + 0.5 * torch.eye(6)
    r = sinfill((n, 3), 0.59, 0.1, 0.4)
    jj = torch.zeros(n, 6, 6)
    jj[:, :3, :3] = torch.eye(3)
    jj[:, 3:, 3:] = torch.eye(3)
    jj[:, 0, 4], jj[:, 0, 5], jj[:, 1, 3], jj[:, 1, 5], jj[:, 2, 3], jj[:, 2, 4] = r[:, 2], -r[:, 1], -r[:, 2], r[:, 0], r[:, 1], -r[:, 0]
    b._j_eef = jj
    dpose = sinfill((n, 6), 0.83, 0.4) * torch.tensor([[0.1, 0.1, 0.1, 0.5, 0.5, 0.5]])
    u = mod.UsefulHound._compute_osc_torques(b, dpose)
    osc = dict(osc_mm=b._mm.numpy().copy(), osc_j=jj.numpy().copy(), osc_dpose=dpose.numpy().copy(), osc_eef_vel=b._eef_state[:, 7:].numpy().copy(),
               osc_q=b._q.numpy().copy(), osc_qd=b._qd.numpy().copy(), osc_u=u.numpy().copy())
    # --- post_physics_step ---
    b.common_step_counter = 5
    csc_before = 5
    reset_draws = sinfill((n, 35), 0.173, 0.05).abs() * 0.999
    noise_draws = sinfill((n, 204), 0.0917, 0.33).abs() * 0.999
    push_draws = sinfill((n, 2), 0.61, 0.21).abs() * 0.999
    state = {"col": 0}

    def fake_rand_float(lower, upper, shape, device):
        ids = b._cur_ids
        if b._draw_col is None:
            uu = push_draws[ids]
        else:
            c0 = b._draw_col
            uu = reset_draws[ids][:, c0:c0 + shape[1]]
            b._draw_col += shape[1]
        return (upper - lower) * uu + lower

    def fake_rand(shape, device=None):
        c0 = b._draw_col
        uu = reset_draws[b._cur_ids][:, c0:c0 + shape[1]]
        b._draw_col += shape[1]
        return uu.clone()

    mod.torch_rand_float = fake_rand_float
    real_rand_like, real_rand = torch.rand_like, torch.rand
    torch.rand_like = lambda t: noise_draws.clone()
    torch.rand = fake_rand
    inputs = dict(root=b.root_states.clone(), dof_state=b.dof_state.clone(), contact=b.contact_forces.clone(), torques=b.torques.clone(),
                  commands=b.commands.clone(), actions=b.actions.clone(), last_actions=b.last_actions.clone(), last_dof_vel=b.last_hound_dof_vel.clone(),
                  feet_air_time=b.feet_air_time.clone(), progress=b.progress_buf.clone(), timeout_prev=b.timeout_buf.clone(), default=b.hound_default_dof_pos.clone(),
                  episode_sums=torch.stack([b.episode_sums[k] for k in keys]).clone(), eef_state=b._eef_state.clone())
    try:
        mod.UsefulHound.post_physics_step(b)
    finally:
        torch.rand_like, torch.rand = real_rand_like, real_rand
    timeout = (b.progress_buf >= b.max_episode_length - 1) & (b.reset_buf != 0)
    outs = dict(o_root=b.root_states, o_dof_state=b.dof_state, o_commands=b.commands, o_obs=b.obs_buf, o_rew=b.rew_buf, o_reset=b.reset_buf.to(torch.int64),
                o_progress=b.progress_buf, o_timeout=timeout.to(torch.int64), o_last_actions=b.last_actions, o_last_dof_vel=b.last_hound_dof_vel,
                o_feet_air_time=b.feet_air_time, o_episode_sums=torch.stack([b.episode_sums[k] for k in keys]), o_measured_heights=b.measured_heights,
                o_extras=torch.tensor([float(b.extras["episode"]["rew_" + k]) for k in keys] + [float(b.extras["episode"]["terrain_level"])]))
    meta = dict(knee=np.array(knee), feet=np.array(feet), base_indices=np.array(base_indices),
                rew_scales=np.array([b.rew_scales[k] for k in ["termination", "lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torque", "joint_acc",
                                                                 "base_height", "air_time", "collision", "stumble", "action_rate", "hip"]], dtype=np.float32),
                noise_scale_vec=b.noise_scale_vec.numpy(), reset_draws=reset_draws.numpy(), noise_draws=noise_draws.numpy(), push_draws=push_draws.numpy(),
                common_step_counter=np.array(csc_before), push_interval=np.array(b.push_interval), max_len=np.array(b.max_episode_length), dt=np.float32(b.dt))
    np.savez_compressed(out, **{k: v.numpy() for k, v in inputs.items()}, **{k: v.numpy() for k, v in outs.items()}, **meta, **osc)
    print(out, "resets", int(b.reset_buf.sum()), "rew range", float(b.rew_buf.min()), float(b.rew_buf.max()), "osc |u| max", float(u.abs().max()))


if __name__ == "__main__" and "--dr-only" not in sys.argv and "--houndarm-only" not in sys.argv and "--manipulator-only" not in sys.argv and ("--useful" in sys.argv or globals().get("_run_terrain_after")):
    gen_useful_hound(os.path.join(HERE, "useful_hound_plane.npz"))


# ------------------------------------------------------------------------------------------------
# domain-randomisation samplers (utils/dr_utils.py): bucketing is deterministic -> golden values; the samplers are random ->
# their first two moments and range over 200k numpy draws, per parameter block and gym step count
# ------------------------------------------------------------------------------------------------
def gen_dr(out):
    dr = ref_loader.load("utils.dr_utils")
    blocks = {
        "mass": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform", "schedule": "linear", "schedule_steps": 3000},
        "friction": {"range": [0.7, 1.3], "operation": "scaling", "distribution": "uniform", "schedule": "linear", "schedule_steps": 3000, "num_buckets": 500},
        "gravity": {"range": [0, 0.4], "operation": "additive", "distribution": "gaussian", "schedule": "linear", "schedule_steps": 3000},
        "gauss_scaling": {"range": [1.2, 0.3], "operation": "scaling", "distribution": "gaussian", "schedule": "constant", "schedule_steps": 1000},
        "loguniform": {"range": [0.3, 3.0], "operation": "scaling", "distribution": "loguniform"},
        "gauss_buckets": {"range": [1.0, 0.04], "operation": "scaling", "distribution": "gaussian", "num_buckets": 250},
    }
    res = {}
    np.random.seed(12345)
    for name, p in blocks.items():
        for step in (0, 500, 1500, 3000, 10000):
            s = dr.generate_random_samples(dict(p), 200000, step)
            res[f"stat_{name}_{step}"] = np.array([s.mean(), s.std(), s.min(), s.max()])
    v = np.linspace(0.55, 1.45, 181)
    for name in ("friction", "gauss_buckets"):
        res[f"bucket_in_{name}"] = v
        res[f"bucket_out_{name}"] = np.array([dr.get_bucketed_val(x, blocks[name]) for x in v])
    import json

    res["blocks_json"] = np.array(json.dumps(blocks))
    np.savez_compressed(out, **res)
    print(out, len(res), "arrays")


if __name__ == "__main__" and ("--dr" in sys.argv or "--dr-only" in sys.argv):
    gen_dr(os.path.join(HERE, "dr_utils.npz"))


# ------------------------------------------------------------------------------------------------
# Houndarm (tasks/hound_arm.py): the jit reward function and the eager OSC torque law run on an attribute bag
# ------------------------------------------------------------------------------------------------
def gen_houndarm(out):
    mod = ref_loader.load("tasks.hound_arm")
    n = 96
    eef_pos = sinfill((n, 3), 0.41, 0.3, 0.3)
    commands = sinfill((n, 3), 0.57, 0.9, 0.3)
    commands[:24] = eef_pos[:24] + sinfill((24, 3), 0.93, 0.2, 0.008)        # a quarter of the envs inside the 2 cm reach ball
    eef_vel = sinfill((n, 6), 0.29, 0.5, 0.4)
    progress = (torch.arange(n) * 7 % 160).long()
    reset = (torch.arange(n) % 11 == 0).long()
    states = {"eef_pos": eef_pos, "eef_vel": eef_vel, "commands": commands}
    settings = {"r_dist_scale": 0.1, "r_lift_scale": 1.5, "r_align_scale": 2.0, "r_stack_scale": 16.0, "r_vel_scale": 0.1}
    rew, rst = mod.compute_houndarm_reward(reset, progress, torch.zeros(n, 6), states, settings, 150.0)
    b = types.SimpleNamespace()
    mm = sinfill((n, 6, 6), 0.37, 0.3, 0.2)
    b._mm = mm @ mm.transpose(1, 2) + 0.5 * torch.eye(6)
    b._j_eef = sinfill((n, 6, 6), 0.61, 0.8, 0.7) + torch.eye(6)
    b._q = sinfill((n, 6), 0.23, 0.4, 1.2)
    b._qd = sinfill((n, 6), 0.19, 0.7, 2.0)
    b.states = {"eef_vel": eef_vel}
    b.kp = torch.full((6,), 150.0)
    b.kd = 2 * torch.sqrt(b.kp)
    b.kp_null = torch.full((6,), 10.0)
    b.kd_null = 2 * torch.sqrt(b.kp_null)
    b.houndarm_default_dof_pos = torch.zeros(6)
    b._houndarm_effort_limits = torch.full((6,), 1000.0)
    b.device = "cpu"
    dpose = sinfill((n, 6), 0.83, 0.4) * torch.tensor([[0.1, 0.1, 0.1, 0.5, 0.5, 0.5]])
    u = mod.Houndarm._compute_osc_torques(b, dpose)
    np.savez_compressed(out, eef_pos=eef_pos.numpy(), commands=commands.numpy(), eef_vel=eef_vel.numpy(), progress=progress.numpy(), reset=reset.numpy(),
                        rew=rew.numpy(), reset_out=rst.numpy(), mm=b._mm.numpy(), j_eef=b._j_eef.numpy(), q=b._q.numpy(), qd=b._qd.numpy(),
                        dpose=dpose.numpy(), u=u.numpy())
    print(out, "rew range", float(rew.min()), float(rew.max()), "in reach", int((torch.norm(eef_pos - commands, dim=-1) < 0.02).sum()),
          "resets", int(rst.sum()), "|u| max", float(u.abs().max()))


if __name__ == "__main__" and "--houndarm-only" in sys.argv:
    gen_houndarm(os.path.join(HERE, "houndarm.npz"))


# ------------------------------------------------------------------------------------------------
# Manipulator (tasks/manipulator.py): reward function, the 7-joint OSC torque law and the reset draw logic (incl. the two trailing joints
# set back without noise, :417) run on an attribute bag
# ------------------------------------------------------------------------------------------------
def gen_manipulator(out):
    mod = ref_loader.load("tasks.manipulator")
    n = 96
    eef_pos = sinfill((n, 3), 0.43, 0.3, 0.3)
    commands = sinfill((n, 3), 0.59, 0.9, 0.3)
    commands[:24] = eef_pos[:24] + sinfill((24, 3), 0.91, 0.2, 0.008)
    eef_vel = sinfill((n, 6), 0.31, 0.5, 0.4)
    progress = (torch.arange(n) * 47 % 1100).long()
    reset = (torch.arange(n) % 11 == 0).long()
    states = {"eef_pos": eef_pos, "eef_vel": eef_vel, "commands": commands}
    settings = {"r_dist_scale": 0.1, "r_lift_scale": 1.5, "r_align_scale": 2.0, "r_stack_scale": 16.0, "r_vel_scale": 0.1}
    rew, rst = mod.compute_franka_reward(reset, progress, torch.zeros(n, 6), states, settings, 1000.0)
    b = types.SimpleNamespace()
    mm = sinfill((n, 7, 7), 0.37, 0.3, 0.2)
    b._mm = mm @ mm.transpose(1, 2) + 0.5 * torch.eye(7)
    b._j_eef = sinfill((n, 6, 7), 0.61, 0.8, 0.7) + torch.eye(6, 7)
    b._q = sinfill((n, 7), 0.23, 0.4, 1.2)
    b._qd = sinfill((n, 7), 0.19, 0.7, 2.0)
    b.states = {"eef_vel": eef_vel}
    b.kp = torch.full((6,), 150.0)
    b.kd = 2 * torch.sqrt(b.kp)
    b.kp_null = torch.full((7,), 10.0)
    b.kd_null = 2 * torch.sqrt(b.kp_null)
    b.franka_default_dof_pos = torch.tensor([0, 0.1963, 0, -2.6180, 0, 2.9416, 0.7854])
    b._franka_effort_limits = torch.tensor([87.0, 87.0, 87.0, 87.0, 12.0, 12.0, 12.0])
    b.device = "cpu"
    dpose = sinfill((n, 6), 0.83, 0.4) * torch.tensor([[0.1, 0.1, 0.1, 0.5, 0.5, 0.5]])
    u = mod.Manipulator._compute_osc_torques(b, dpose)
    # reset_idx (:385-447) on a bag whose gym calls are no-ops; torch's global generator supplies the draws, recorded for the test
    r = types.SimpleNamespace()
    r.device = "cpu"
    r.command_x_range, r.command_y_range, r.command_z_range = [-0.5, 0.5], [-0.5, 0.5], [0.2, 0.6]
    r.commands = torch.zeros(n, 3)
    r.commands_x, r.commands_y, r.commands_z = (r.commands.view(n, 3)[..., i] for i in range(3))
    r.franka_default_dof_pos = b.franka_default_dof_pos
    r.franka_dof_noise = 0.25
    r.franka_dof_lower_limits = torch.tensor([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
    r.franka_dof_upper_limits = torch.tensor([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
    r._q, r._qd = sinfill((n, 7), 0.3, 0.2, 1.0), sinfill((n, 7), 0.7, 0.2, 1.0)
    r._pos_control, r._effort_control = torch.zeros(n, 7), torch.ones(n, 7)
    r._global_indices = torch.arange(n, dtype=torch.int32).view(n, -1)
    r._dof_state = torch.zeros(n, 7, 2)
    noop = lambda *a, **k: None
    r.gym = types.SimpleNamespace(set_dof_position_target_tensor_indexed=noop, set_dof_actuation_force_tensor_indexed=noop, set_dof_state_tensor_indexed=noop)
    r.sim = None
    r.progress_buf, r.reset_buf = torch.full((n,), 5, dtype=torch.long), torch.ones(n, dtype=torch.long)
    env_ids = torch.arange(0, n, 3)
    torch.manual_seed(123)
    st = torch.get_rng_state()
    k = len(env_ids)
    draws = torch.cat([torch.rand(k, 1), torch.rand(k, 1), torch.rand(k, 1), torch.rand(k, 7)], dim=1)      # the order reset_idx draws in
    torch.set_rng_state(st)
    q_before = r._q.clone()
    mod.Manipulator.reset_idx(r, env_ids)
    np.savez_compressed(out, eef_pos=eef_pos.numpy(), commands=commands.numpy(), eef_vel=eef_vel.numpy(), progress=progress.numpy(), reset=reset.numpy(),
                        rew=rew.numpy(), reset_out=rst.numpy(), mm=b._mm.numpy(), j_eef=b._j_eef.numpy(), q=b._q.numpy(), qd=b._qd.numpy(),
                        dpose=dpose.numpy(), u=u.numpy(), effort=b._franka_effort_limits.numpy(), default_q=b.franka_default_dof_pos.numpy(),
                        reset_env_ids=env_ids.numpy(), reset_draws=draws.numpy(), reset_q_before=q_before.numpy(), reset_q=r._q.numpy(), reset_qd=r._qd.numpy(),
                        reset_commands=r.commands.numpy(), reset_progress=r.progress_buf.numpy(), reset_reset=r.reset_buf.numpy(),
                        lower=r.franka_dof_lower_limits.numpy(), upper=r.franka_dof_upper_limits.numpy())
    print(out, "rew range", float(rew.min()), float(rew.max()), "resets", int(rst.sum()), "|u| max", float(u.abs().max()),
          "clamped torques", int((u.abs() >= b._franka_effort_limits - 1e-6).sum()))


if __name__ == "__main__" and "--manipulator-only" in sys.argv:
    gen_manipulator(os.path.join(HERE, "manipulator.npz"))
