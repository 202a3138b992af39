"""Rough-terrain ANYmal task, drop-in for the reference's ``tasks/anymal_terrain.py`` (``AnymalTerrain`` :43-538).

Same configuration keys, attributes and step semantics: explicit PD torques recomputed ``decimation`` times per policy step
(:441-451) plus the generic loop's extra sim step (``vec_task.py:379-382``), 188 observations with the 14 x 10 height scan,
13-term reward, terrain curriculum, pushes, observation noise, ``extras["episode"]`` reward means, boolean ``reset_buf``.

``step()`` is two launches of hand-written sm_100a kernels (``k_terrain_phys``: 5 sim steps + termination + reward;
``k_terrain_post``: reset + curriculum + observations + noise + the ``extras`` means, reduced by the last block to arrive); every task
buffer (``commands``, ``torques``, ``last_actions``, ``feet_air_time``, ``episode_sums``, ``terrain_levels`` ...) aliases
sim-owned device memory.  The terrain itself comes from ``isaacgymenv_b200.terrain`` (Isaac Gym's ``terrain_utils`` is not
part of the reference tree) and is collided as a height grid.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _abi, _lib, gymapi, gymtorch
from ..terrain import Terrain
from ..utils.torch_math import get_axis_params, to_torch, torch_rand_float
from .anymal import default_asset_root
from .base.vec_task import VecTask

REW_ORDER = ("termination", "lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torque", "joint_acc", "base_height", "air_time",
             "collision", "stumble", "action_rate", "hip")
EPISODE_KEYS = ("lin_vel_xy", "lin_vel_z", "ang_vel_z", "ang_vel_xy", "orient", "torques", "joint_acc", "base_height", "air_time", "collision",
                "stumble", "action_rate", "hip")


class AnymalTerrain(VecTask):
    ACTOR_NAME = "anymal"
    BASE_NAME = "base"
    HOUND_TERMINATION = False
    BASE_HEIGHT_TARGET = 0.52

    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture=False, force_render=False):
        self.cfg = cfg
        self.height_samples = None
        self.custom_origins = False
        self.debug_viz = cfg["env"]["enableDebugVis"]
        self.init_done = False
        learn = cfg["env"]["learn"]
        self.lin_vel_scale, self.ang_vel_scale = learn["linearVelocityScale"], learn["angularVelocityScale"]
        self.dof_pos_scale, self.dof_vel_scale = learn["dofPositionScale"], learn["dofVelocityScale"]
        self.height_meas_scale = learn["heightMeasurementScale"]
        self.action_scale = cfg["env"]["control"]["actionScale"]
        self.rew_scales = {
            "termination": learn["terminalReward"], "lin_vel_xy": learn["linearVelocityXYRewardScale"], "lin_vel_z": learn["linearVelocityZRewardScale"],
            "ang_vel_z": learn["angularVelocityZRewardScale"], "ang_vel_xy": learn["angularVelocityXYRewardScale"], "orient": learn["orientationRewardScale"],
            "torque": learn["torqueRewardScale"], "joint_acc": learn["jointAccRewardScale"], "base_height": learn["baseHeightRewardScale"],
            "air_time": learn["feetAirTimeRewardScale"], "collision": learn["kneeCollisionRewardScale"], "stumble": learn["feetStumbleRewardScale"],
            "action_rate": learn["actionRateRewardScale"], "hip": learn["hipRewardScale"]}
        rng = cfg["env"]["randomCommandVelocityRanges"]
        self.command_x_range, self.command_y_range, self.command_yaw_range = rng["linear_x"], rng["linear_y"], rng["yaw"]
        init = cfg["env"]["baseInitState"]
        self.base_init_state = init["pos"] + init["rot"] + init["vLinear"] + init["vAngular"]
        self.named_default_joint_angles = cfg["env"]["defaultJointAngles"]
        self.decimation = cfg["env"]["control"]["decimation"]
        self.dt = self.decimation * cfg["sim"]["dt"]
        self.max_episode_length_s = learn["episodeLength_s"]
        self.max_episode_length = int(self.max_episode_length_s / self.dt + 0.5)
        self.push_interval = int(learn["pushInterval_s"] / self.dt + 0.5)
        self.allow_knee_contacts = learn["allowKneeContacts"]
        self.Kp, self.Kd = cfg["env"]["control"]["stiffness"], cfg["env"]["control"]["damping"]
        self.curriculum = cfg["env"]["terrain"]["curriculum"]
        for k in self.rew_scales:
            self.rew_scales[k] *= self.dt
        self.seed = int(cfg.get("seed", 42))
        # cfg["task"]["randomize"] is never read by the reference's terrain tasks (no apply_randomizations call in
        # tasks/anymal_terrain.py); their only randomisation is the friction-bucket draw of _create_envs (:235-247), done below

        super().__init__(config=self.cfg, rl_device=rl_device, sim_device=sim_device, graphics_device_id=graphics_device_id, headless=headless,
                         virtual_screen_capture=virtual_screen_capture, force_render=force_render)
        self.dt = self.decimation * cfg["sim"]["dt"]      # VecTask.__init__ sets dt to the sim dt; the task uses the policy dt

        self.root_states = gymtorch.wrap_tensor(self.gym.acquire_actor_root_state_tensor(self.sim))
        self.dof_state = gymtorch.wrap_tensor(self.gym.acquire_dof_state_tensor(self.sim))
        self.dof_pos = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 0]
        self.dof_vel = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 1]
        self.contact_forces = gymtorch.wrap_tensor(self.gym.acquire_net_contact_force_tensor(self.sim)).view(self.num_envs, -1, 3)
        self.common_step_counter = 0
        self._device_step = False
        self.extras = {}
        self.noise_scale_vec = self._get_noise_scale_vec(cfg)
        self.commands_scale = torch.tensor([self.lin_vel_scale, self.lin_vel_scale, self.ang_vel_scale], device=self.device, requires_grad=False)
        self.gravity_vec = to_torch(get_axis_params(-1.0, self.up_axis_idx), device=self.device).repeat((self.num_envs, 1))
        self.forward_vec = to_torch([1.0, 0.0, 0.0], device=self.device).repeat((self.num_envs, 1))
        self.default_dof_pos = torch.zeros_like(self.dof_pos, dtype=torch.float, device=self.device, requires_grad=False)
        for i in range(self.num_dof):
            self.default_dof_pos[:, i] = self.named_default_joint_angles.get(self.dof_names[i], 0.0)
        self.height_points = self.init_height_points()
        self._create_fused_task()
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        self.init_done = True
        _lib.check(self._lib.b2g_task_terrain_set_init_done(self.sim.handle, 1))

    # ------------------------------------------------------------------ sim construction
    def create_sim(self):
        self.up_axis_idx = 2
        self.sim = super().create_sim(self.device_id, self.graphics_device_id, self.physics_engine, self.sim_params)
        terrain_type = self.cfg["env"]["terrain"]["terrainType"]
        if terrain_type == "plane":
            self._create_ground_plane()
        elif terrain_type == "trimesh":
            self._create_trimesh()
            self.custom_origins = True
        self._create_envs(self.num_envs, self.cfg["env"]["envSpacing"], int(np.sqrt(self.num_envs)))

    def _get_noise_scale_vec(self, cfg):
        learn = cfg["env"]["learn"]
        v = torch.zeros(self.num_obs, device=self.device)
        self.add_noise = learn["addNoise"]
        lvl = learn["noiseLevel"]
        v[:3] = learn["linearVelocityNoise"] * lvl * self.lin_vel_scale
        v[3:6] = learn["angularVelocityNoise"] * lvl * self.ang_vel_scale
        v[6:9] = learn["gravityNoise"] * lvl
        v[12:24] = learn["dofPositionNoise"] * lvl * self.dof_pos_scale
        v[24:36] = learn["dofVelocityNoise"] * lvl * self.dof_vel_scale
        v[36:176] = learn["heightMeasurementNoise"] * lvl * self.height_meas_scale
        return v

    def _create_ground_plane(self):
        t = self.cfg["env"]["terrain"]
        p = gymapi.PlaneParams()
        p.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        p.static_friction, p.dynamic_friction, p.restitution = t["staticFriction"], t["dynamicFriction"], t["restitution"]
        self.gym.add_ground(self.sim, p)

    def _create_trimesh(self):
        t = self.cfg["env"]["terrain"]
        self.terrain = Terrain(t, num_robots=self.num_envs, seed=self.seed)
        hp = gymapi.HeightFieldParams()
        hp.nbRows, hp.nbColumns = self.terrain.tot_rows, self.terrain.tot_cols
        hp.row_scale = hp.column_scale = self.terrain.horizontal_scale
        hp.vertical_scale = self.terrain.vertical_scale
        hp.transform.p.x = hp.transform.p.y = -self.terrain.border_size
        hp.static_friction, hp.dynamic_friction, hp.restitution = t["staticFriction"], t["dynamicFriction"], t["restitution"]
        # the reference hands PhysX a triangle mesh of this very grid (anymal_terrain.py:196-209); this engine collides the grid itself
        self.gym.add_heightfield(self.sim, self.terrain.heightsamples, hp)
        self.height_samples = torch.tensor(self.terrain.heightsamples).view(self.terrain.tot_rows, self.terrain.tot_cols).to(self.device)

    def _asset_options(self):
        o = gymapi.AssetOptions()
        o.default_dof_drive_mode = gymapi.DOF_MODE_EFFORT
        o.collapse_fixed_joints = True
        o.replace_cylinder_with_capsule = True
        o.flip_visual_attachments = True
        o.fix_base_link = self.cfg["env"]["urdfAsset"]["fixBaseLink"]
        o.density, o.angular_damping, o.linear_damping, o.armature, o.thickness, o.disable_gravity = 0.001, 0.0, 0.0, 0.0, 0.01, False
        return o

    def _extra_termination_names(self, body_names):
        return []

    def _create_envs(self, num_envs, spacing, num_per_row):
        asset_root = self.cfg["env"].get("assetRoot", default_asset_root())
        asset_file = self.cfg["env"]["urdfAsset"]["file"]
        asset = self.gym.load_asset(self.sim, asset_root, asset_file, self._asset_options())
        self.num_dof = self.gym.get_asset_dof_count(asset)
        self.num_bodies = self.gym.get_asset_rigid_body_count(asset)
        rigid_shape_prop = self.gym.get_asset_rigid_shape_properties(asset)
        friction_range = self.cfg["env"]["learn"]["frictionRange"]
        num_buckets = 100
        friction_buckets = torch_rand_float(friction_range[0], friction_range[1], (num_buckets, 1), device=self.device)
        self.base_init_state = to_torch(self.base_init_state, device=self.device, requires_grad=False)
        start_pose = gymapi.Transform()
        start_pose.p = gymapi.Vec3(*self.base_init_state[:3])
        body_names = self.gym.get_asset_rigid_body_names(asset)
        self.dof_names = self.gym.get_asset_dof_names(asset)
        foot_name, knee_name = self.cfg["env"]["urdfAsset"]["footName"], self.cfg["env"]["urdfAsset"]["kneeName"]
        feet_names = [s for s in body_names if foot_name in s]
        knee_names = [s for s in body_names if knee_name in s]
        extra_names = self._extra_termination_names(body_names)
        dof_props = self.gym.get_asset_dof_properties(asset)
        self.env_origins = torch.zeros(self.num_envs, 3, device=self.device, requires_grad=False)
        tcfg = self.cfg["env"]["terrain"]
        if not self.curriculum:
            tcfg["maxInitMapLevel"] = tcfg["numLevels"] - 1
        self.terrain_levels = torch.randint(0, tcfg["maxInitMapLevel"] + 1, (self.num_envs,), device=self.device)
        self.terrain_types = torch.randint(0, tcfg["numTerrains"], (self.num_envs,), device=self.device)
        if self.custom_origins:
            self.terrain_origins = torch.from_numpy(self.terrain.env_origins).to(self.device).to(torch.float)
            spacing = 0.0
            self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]
            start_xy = (self.env_origins[:, :2] + torch_rand_float(-1.0, 1.0, (self.num_envs, 2), device=self.device)).cpu().numpy()
            origins_cpu = self.env_origins.cpu().numpy()
        fr = friction_buckets.flatten().cpu().numpy()
        lower = gymapi.Vec3(-spacing, -spacing, 0.0)
        upper = gymapi.Vec3(spacing, spacing, spacing)
        self.envs, self.anymal_handles = [], []
        for i in range(self.num_envs):
            env = self.gym.create_env(self.sim, lower, upper, num_per_row)
            if self.custom_origins:
                start_pose = gymapi.Transform()
                start_pose.p = gymapi.Vec3(float(start_xy[i, 0]), float(start_xy[i, 1]), float(origins_cpu[i, 2]))
            for s in range(len(rigid_shape_prop)):
                rigid_shape_prop[s].friction = float(fr[i % num_buckets])
            self.gym.set_asset_rigid_shape_properties(asset, rigid_shape_prop)
            handle = self.gym.create_actor(env, asset, start_pose, self.ACTOR_NAME, i, 0, 0)
            self.envs.append(env)
            self.anymal_handles.append(handle)
        self.gym.set_actor_dof_properties(self.envs[0], self.anymal_handles[0], dof_props)
        find = lambda n: self.gym.find_actor_rigid_body_handle(self.envs[0], self.anymal_handles[0], n)
        self.feet_indices = torch.tensor([find(n) for n in feet_names], dtype=torch.long, device=self.device)
        self.knee_indices = torch.tensor([find(n) for n in knee_names], dtype=torch.long, device=self.device)
        self.base_indices = torch.tensor([find(n) for n in extra_names], dtype=torch.long, device=self.device)
        self.base_index = find(self.BASE_NAME)

    # ------------------------------------------------------------------ fused task
    def _task_tensor(self, kind):
        d = _abi.TensorDesc()
        _lib.check(self._lib.b2g_task_tensor(self.sim.handle, kind, C.byref(d)), "task tensor")
        return _lib.desc_to_torch(d)

    def _fused_cfg(self) -> _abi.TerrainCfg:
        c = _abi.TerrainCfg()
        c.lin_vel_scale, c.ang_vel_scale, c.dof_pos_scale, c.dof_vel_scale = self.lin_vel_scale, self.ang_vel_scale, self.dof_pos_scale, self.dof_vel_scale
        c.height_meas_scale, c.action_scale = self.height_meas_scale, self.action_scale
        c.kp, c.kd, c.torque_limit = self.Kp, self.Kd, 80.0
        c.decimation, c.extra_sim_steps, c.dt = int(self.decimation), int(self.control_freq_inv), float(self.dt)
        for i, k in enumerate(REW_ORDER):
            c.rew[i] = float(self.rew_scales[k])
        c.base_height_target = self.BASE_HEIGHT_TARGET
        big = 3.0e38
        c.clip_obs, c.clip_actions = float(min(self.clip_obs, big)), float(min(self.clip_actions, big))
        for i in range(2):
            c.cmd_x[i], c.cmd_y[i], c.cmd_yaw[i] = self.command_x_range[i], self.command_y_range[i], self.command_yaw_range[i]
        for i, v in enumerate(self.default_dof_pos[0].tolist()):
            c.default_dof_pos[i] = v
        for i, v in enumerate(self.base_init_state.tolist()):
            c.init_root[i] = v
        nv = self.noise_scale_vec.tolist()
        c.add_noise = 1 if self.add_noise else 0
        c.noise_lin_vel, c.noise_ang_vel, c.noise_gravity, c.noise_dof_pos, c.noise_dof_vel, c.noise_height = nv[0], nv[3], nv[6], nv[12], nv[24], nv[36]
        c.base_body = int(self.base_index)
        for name, cnt, arr, src in (("knee", "n_knee", c.knee_bodies, self.knee_indices), ("feet", "n_feet", c.feet_bodies, self.feet_indices),
                                    ("extra", "n_term_extra", c.term_extra_bodies, self.base_indices)):
            vals = src.tolist()
            setattr(c, cnt, len(vals))
            for i, v in enumerate(vals):
                arr[i] = int(v)
        c.hound_termination = 1 if self.HOUND_TERMINATION else 0
        c.allow_knee_contacts = 1 if self.allow_knee_contacts else 0
        for i, d in enumerate([0, 3, 6, 9]):
            c.hip_dofs[i] = d
        c.max_episode_length, c.push_interval, c.max_episode_length_s = int(self.max_episode_length), int(self.push_interval), float(self.max_episode_length_s)
        c.custom_origins, c.curriculum = int(self.custom_origins), int(bool(self.curriculum))
        x = 0.1 * np.array([-8, -7, -6, -5, -4, -3, -2, 2, 3, 4, 5, 6, 7, 8], dtype=np.float32)
        y = 0.1 * np.array([-5, -4, -3, -2, -1, 1, 2, 3, 4, 5], dtype=np.float32)
        c.n_hx, c.n_hy = len(x), len(y)
        for i, v in enumerate(x):
            c.hx[i] = float(v)
        for i, v in enumerate(y):
            c.hy[i] = float(v)
        if self.custom_origins:
            t = self.terrain
            c.hs_rows, c.hs_cols = t.tot_rows, t.tot_cols
            c.border_size, c.hscale, c.vscale, c.env_length = float(t.border_size), t.horizontal_scale, t.vertical_scale, float(t.env_length)
            c.env_rows, c.env_cols = t.env_rows, t.env_cols
        c.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        c.arm_chain = -1          # no manipulator (UsefulHound overrides)
        return c

    def _create_fused_task(self):
        self._lib = _lib.load()
        cfg = self._fused_cfg()
        hs = og = None
        if self.custom_origins:
            self._hs_host = np.ascontiguousarray(self.terrain.heightsamples, np.int16)
            self._og_host = np.ascontiguousarray(self.terrain.env_origins, np.float32)
            hs, og = self._hs_host.ctypes.data_as(C.c_void_p), self._og_host.ctypes.data_as(C.c_void_p)
        _lib.check(self._lib.b2g_task_terrain_create(self.sim.handle, C.byref(cfg), hs, og), "task create")
        t = self._task_tensor
        self.obs_buf, self.obs_clamped, self.rew_buf = t(_abi.TT_OBS), t(_abi.TT_OBS_CLAMPED), t(_abi.TT_REW)
        self._reset_i64, self.progress_buf, self._timeout_i64 = t(_abi.TT_RESET), t(_abi.TT_PROGRESS), t(_abi.TT_TIMEOUT)
        self.reset_buf, self.timeout_buf = self._reset_i64, self._timeout_i64
        self.commands, self.actions, self.torques = t(_abi.TT_COMMANDS), t(_abi.TT_ACTIONS), t(_abi.TT_TORQUES)
        self.last_actions, self.last_dof_vel, self.feet_air_time = t(_abi.TT_LAST_ACTIONS), t(_abi.TT_LAST_DOF_VEL), t(_abi.TT_FEET_AIR_TIME)
        self._episode_sums = t(_abi.TT_EPISODE_SUMS)
        self.episode_sums = {k: self._episode_sums[i] for i, k in enumerate(EPISODE_KEYS)}
        self.measured_heights = t(_abi.TT_MEASURED_HEIGHTS)
        self._extras = t(_abi.TT_EXTRAS)
        env_origins, levels, types = t(_abi.TT_ENV_ORIGINS), t(_abi.TT_TERRAIN_LEVELS), t(_abi.TT_TERRAIN_TYPES)
        env_origins.copy_(self.env_origins)
        levels.copy_(self.terrain_levels)
        types.copy_(self.terrain_types)
        self.env_origins, self.terrain_levels, self.terrain_types = env_origins, levels, types

    def init_height_points(self):
        y = 0.1 * torch.tensor([-5, -4, -3, -2, -1, 1, 2, 3, 4, 5], device=self.device, requires_grad=False)
        x = 0.1 * torch.tensor([-8, -7, -6, -5, -4, -3, -2, 2, 3, 4, 5, 6, 7, 8], device=self.device, requires_grad=False)
        gx, gy = torch.meshgrid(x, y, indexing="ij")
        self.num_height_points = gx.numel()
        pts = torch.zeros(self.num_envs, self.num_height_points, 3, device=self.device, requires_grad=False)
        pts[:, :, 0], pts[:, :, 1] = gx.flatten(), gy.flatten()
        return pts

    def reset_idx(self, env_ids):
        """Reference :384-425.  Used for the constructor's reset of all envs (torch RNG, like the reference); during
        stepping resets happen inside the kernels."""
        n = len(env_ids)
        nleg = 12
        self.dof_pos[env_ids, :nleg] = self.default_dof_pos[env_ids, :nleg] * torch_rand_float(0.5, 1.5, (n, nleg), device=self.device)
        self.dof_vel[env_ids, :nleg] = torch_rand_float(-0.1, 0.1, (n, nleg), device=self.device)
        self.root_states[env_ids] = self.base_init_state
        if self.custom_origins:
            self.root_states[env_ids, :3] += self.env_origins[env_ids]
            self.root_states[env_ids, :2] += torch_rand_float(-0.5, 0.5, (n, 2), device=self.device)
        self.commands[env_ids, 0] = torch_rand_float(self.command_x_range[0], self.command_x_range[1], (n, 1), device=self.device).squeeze(1)
        self.commands[env_ids, 1] = torch_rand_float(self.command_y_range[0], self.command_y_range[1], (n, 1), device=self.device).squeeze(1)
        self.commands[env_ids, 3] = torch_rand_float(self.command_yaw_range[0], self.command_yaw_range[1], (n, 1), device=self.device).squeeze(1)
        self.commands[env_ids] *= (torch.norm(self.commands[env_ids, :2], dim=1) > 0.25).unsqueeze(1)
        self.last_actions[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.progress_buf[env_ids] = 0
        self._reset_i64[env_ids] = 1
        self._episode_sums[:, env_ids] = 0.0

    def enable_device_step_counter(self, enable: bool = True):
        """Keep ``common_step_counter`` (push schedule, noise stream) on the device so that ``step()`` has no per-step host
        state and a rollout can be captured in a CUDA graph; the Python attribute then counts eager calls only."""
        _lib.check(self._lib.b2g_task_terrain_device_step(self.sim.handle, 1 if enable else 0), "device_step")
        self._device_step = bool(enable)

    def pre_physics_step(self, actions):
        raise NotImplementedError("the terrain tasks run fused (step() launches the kernels directly)")

    def post_physics_step(self):
        raise NotImplementedError("the terrain tasks run fused (step() launches the kernels directly)")

    def step(self, actions):
        a = actions.to(self.device, torch.float32)
        if not a.is_contiguous():
            a = a.contiguous()
        self._last_actions_in = a
        self.common_step_counter += 1
        if not self._device_step:
            _lib.check(self._lib.b2g_task_terrain_set_step(self.sim.handle, int(self.common_step_counter)))
        _lib.check(self._lib.b2g_task_step(self.sim.handle, C.c_void_p(a.data_ptr()), self.sim.stream()), "step")
        self.control_steps += 1
        # reference: reset_buf / timeout_buf are bool in the terrain tasks (anymal_terrain.py:295, vec_task.py:394)
        self.reset_buf = self._reset_i64.bool()
        self.timeout_buf = self._timeout_i64.bool()
        ex = self._extras
        self.extras["episode"] = {"rew_" + k: ex[i] for i, k in enumerate(EPISODE_KEYS)}
        self.extras["episode"]["terrain_level"] = ex[13]
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def reset(self):
        self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
        return self.obs_dict
