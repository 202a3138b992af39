"""Flat-terrain ANYmal-C velocity-tracking task, drop-in for the reference's ``tasks/anymal.py``.

Same constructor, attributes and step semantics as the reference class (``Anymal`` :42-304):
12 position-drive actions, 48 observations (:354-386), reward = velocity tracking + torque penalty
(:311-351), reset on base/knee contact or time-out, resets applied at the start of the *next*
``post_physics_step`` (:231-239).

Two execution paths, selected by ``cfg["env"]["fusedStep"]`` (default True):
* fused: ``step()`` is ONE launch of the hand-written sm_100a kernel ``k_anymal_step`` (clamp + targets, all
  physics sub-steps, reset, observations, reward, time-outs); ``obs_buf`` / ``rew_buf`` / ``reset_buf`` /
  ``progress_buf`` / ``commands`` / ``actions`` alias sim-owned device memory through DLPack;
* generic: the reference's hook structure (``pre_physics_step`` -> ``gym.simulate`` -> ``post_physics_step``)
  on the gym tensor API with torch ops; used to cross-check the fused kernel and as the template for
  user-written tasks.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _abi, _lib, gymapi, gymtorch
from ..utils.torch_math import get_axis_params, quat_rotate, quat_rotate_inverse, to_torch, torch_rand_float
from .base.vec_task import VecTask

_ASSET_ROOT_CANDIDATES = (os.environ.get("B2G_ASSET_ROOT", ""), "/root/reference/assets")


def default_asset_root():
    for c in _ASSET_ROOT_CANDIDATES:
        if c and os.path.isdir(c):
            return c
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets")


class Anymal(VecTask):
    ASSET_FILE = "urdf/anymal_c/urdf/anymal.urdf"
    ACTOR_NAME = "anymal"
    BASE_NAME = "base"
    KNEE_KEY = "THIGH"

    def _extremity_key(self, collapse):
        return "SHANK" if collapse else "FOOT"

    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture=False, force_render=False):
        self.cfg = cfg
        learn = cfg["env"]["learn"]
        self.lin_vel_scale = learn["linearVelocityScale"]
        self.ang_vel_scale = learn["angularVelocityScale"]
        self.dof_pos_scale = learn["dofPositionScale"]
        self.dof_vel_scale = learn["dofVelocityScale"]
        self.action_scale = cfg["env"]["control"]["actionScale"]
        self.rew_scales = {"lin_vel_xy": learn["linearVelocityXYRewardScale"], "ang_vel_z": learn["angularVelocityZRewardScale"],
                           "torque": learn["torqueRewardScale"]}
        self.randomization_params = cfg["task"]["randomization_params"]
        self.randomize = cfg["task"]["randomize"]
        rng = cfg["env"]["randomCommandVelocityRanges"]
        self.command_x_range, self.command_y_range, self.command_yaw_range = rng["linear_x"], rng["linear_y"], rng["yaw"]
        plane = cfg["env"]["plane"]
        self.plane_static_friction, self.plane_dynamic_friction, self.plane_restitution = plane["staticFriction"], plane["dynamicFriction"], plane["restitution"]
        init = cfg["env"]["baseInitState"]
        self.base_init_state = init["pos"] + init["rot"] + init["vLinear"] + init["vAngular"]
        self.named_default_joint_angles = cfg["env"]["defaultJointAngles"]
        self.cfg["env"]["numObservations"] = 48
        self.cfg["env"]["numActions"] = 12
        self.fused = bool(cfg["env"].get("fusedStep", True))
        # step() of the generic hook path synchronises with the host (reset_buf.nonzero()), and domain randomisation keeps host-side
        # schedules and generators: neither can be replayed from a CUDA graph (learning/ppo.py and train.py look at this flag)
        self.needs_host_sync = (not self.fused) or bool(self.randomize)
        self.seed = int(cfg.get("seed", 42))

        super().__init__(config=self.cfg, rl_device=rl_device, sim_device=sim_device, graphics_device_id=graphics_device_id,
                         headless=headless, virtual_screen_capture=virtual_screen_capture, force_render=force_render)

        self.dt = self.sim_params.dt
        self.max_episode_length_s = learn["episodeLength_s"]
        self.max_episode_length = int(self.max_episode_length_s / self.dt + 0.5)
        self.Kp = cfg["env"]["control"]["stiffness"]
        self.Kd = cfg["env"]["control"]["damping"]
        for k in self.rew_scales:
            self.rew_scales[k] *= self.dt

        # sim state tensors (views of live sim memory, reference :110-126)
        self.gym.refresh_dof_state_tensor(self.sim)
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_net_contact_force_tensor(self.sim)
        self.gym.refresh_dof_force_tensor(self.sim)
        self.root_states = gymtorch.wrap_tensor(self.gym.acquire_actor_root_state_tensor(self.sim))
        self.dof_state = gymtorch.wrap_tensor(self.gym.acquire_dof_state_tensor(self.sim))
        self.dof_pos = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 0]
        self.dof_vel = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 1]
        self.contact_forces = gymtorch.wrap_tensor(self.gym.acquire_net_contact_force_tensor(self.sim)).view(self.num_envs, -1, 3)
        self.torques = gymtorch.wrap_tensor(self.gym.acquire_dof_force_tensor(self.sim)).view(self.num_envs, self.num_dof)

        self.default_dof_pos = torch.zeros_like(self.dof_pos, dtype=torch.float, device=self.device, requires_grad=False)
        for i in range(self.cfg["env"]["numActions"]):
            self.default_dof_pos[:, i] = self.named_default_joint_angles[self.dof_names[i]]
        self.initial_root_states = self.root_states.clone()
        self.initial_root_states[:] = to_torch(self.base_init_state, device=self.device, requires_grad=False)
        self.gravity_vec = to_torch(get_axis_params(-1.0, self.up_axis_idx), device=self.device).repeat((self.num_envs, 1))

        if self.fused:
            self._create_fused_task()
        else:
            self.commands = torch.zeros(self.num_envs, 3, dtype=torch.float, device=self.device, requires_grad=False)
            self.actions = torch.zeros(self.num_envs, self.num_actions, dtype=torch.float, device=self.device, requires_grad=False)
        self.commands_x = self.commands.view(self.num_envs, 3)[..., 0]
        self.commands_y = self.commands.view(self.num_envs, 3)[..., 1]
        self.commands_yaw = self.commands.view(self.num_envs, 3)[..., 2]
        self.extras = {}
        # reference :154-156: randomise once before the first sim step
        if self.randomize:
            self.apply_randomizations(self.randomization_params)
        self.reset_idx(torch.arange(self.num_envs, device=self.device))

    # ------------------------------------------------------------------ sim construction
    def create_sim(self):
        self.up_axis_idx = 2
        self.sim = super().create_sim(self.device_id, self.graphics_device_id, self.physics_engine, self.sim_params)
        self._create_ground_plane()
        self._create_envs(self.num_envs, self.cfg["env"]["envSpacing"], int(np.sqrt(self.num_envs)))

    def _create_ground_plane(self):
        p = gymapi.PlaneParams()
        p.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        p.static_friction = self.plane_static_friction
        p.dynamic_friction = self.plane_dynamic_friction
        p.restitution = self.plane_restitution
        self.gym.add_ground(self.sim, p)

    def _asset_options(self):
        o = gymapi.AssetOptions()
        o.default_dof_drive_mode = gymapi.DOF_MODE_NONE
        o.collapse_fixed_joints = True
        o.replace_cylinder_with_capsule = True
        o.flip_visual_attachments = True
        o.fix_base_link = self.cfg["env"]["urdfAsset"]["fixBaseLink"]
        o.density = 0.001
        o.angular_damping = 0.0
        o.linear_damping = 0.0
        o.armature = 0.0
        o.thickness = 0.01
        o.disable_gravity = False
        return o

    def _create_envs(self, num_envs, spacing, num_per_row):
        asset_root = self.cfg["env"].get("assetRoot", default_asset_root())
        options = self._asset_options()
        asset = self.gym.load_asset(self.sim, asset_root, self.ASSET_FILE, options)
        self.num_dof = self.gym.get_asset_dof_count(asset)
        self.num_bodies = self.gym.get_asset_rigid_body_count(asset)
        start_pose = gymapi.Transform()
        start_pose.p = gymapi.Vec3(*self.base_init_state[:3])
        body_names = self.gym.get_asset_rigid_body_names(asset)
        self.dof_names = self.gym.get_asset_dof_names(asset)
        feet_names = [s for s in body_names if self._extremity_key(options.collapse_fixed_joints) in s]
        knee_names = [s for s in body_names if self.KNEE_KEY in s]
        self.feet_indices = torch.zeros(len(feet_names), dtype=torch.long, device=self.device, requires_grad=False)
        self.knee_indices = torch.zeros(len(knee_names), dtype=torch.long, device=self.device, requires_grad=False)
        dof_props = self.gym.get_asset_dof_properties(asset)
        for i in range(self.num_dof):
            dof_props["driveMode"][i] = gymapi.DOF_MODE_POS
            dof_props["stiffness"][i] = self.cfg["env"]["control"]["stiffness"]
            dof_props["damping"][i] = self.cfg["env"]["control"]["damping"]
        lower = gymapi.Vec3(-spacing, -spacing, 0.0)
        upper = gymapi.Vec3(spacing, spacing, spacing)
        self.envs, self.actor_handles = [], []
        for i in range(num_envs):
            env = self.gym.create_env(self.sim, lower, upper, num_per_row)
            handle = self.gym.create_actor(env, asset, start_pose, self.ACTOR_NAME, i, 1, 0)
            self.envs.append(env)
            self.actor_handles.append(handle)
        # drive properties are per sim in this engine: one call instead of num_envs identical ones
        self.gym.set_actor_dof_properties(self.envs[0], self.actor_handles[0], dof_props)
        self.gym.enable_actor_dof_force_sensors(self.envs[0], self.actor_handles[0])
        for i, n in enumerate(feet_names):
            self.feet_indices[i] = self.gym.find_actor_rigid_body_handle(self.envs[0], self.actor_handles[0], n)
        for i, n in enumerate(knee_names):
            self.knee_indices[i] = self.gym.find_actor_rigid_body_handle(self.envs[0], self.actor_handles[0], n)
        self.base_index = self.gym.find_actor_rigid_body_handle(self.envs[0], self.actor_handles[0], self.BASE_NAME)
        self.anymal_handles = self.actor_handles

    # ------------------------------------------------------------------ fused path
    def _task_tensor(self, kind):
        d = _abi.TensorDesc()
        _lib.check(self._lib.b2g_task_tensor(self.sim.handle, kind, C.byref(d)), "task tensor")
        return _lib.desc_to_torch(d)

    def _fused_cfg(self) -> _abi.AnymalCfg:
        c = _abi.AnymalCfg()
        c.lin_vel_scale, c.ang_vel_scale = self.lin_vel_scale, self.ang_vel_scale
        c.dof_pos_scale, c.dof_vel_scale, c.action_scale = self.dof_pos_scale, self.dof_vel_scale, self.action_scale
        c.rew_lin_vel_xy, c.rew_ang_vel_z, c.rew_torque = self.rew_scales["lin_vel_xy"], self.rew_scales["ang_vel_z"], self.rew_scales["torque"]
        big = 3.0e38
        c.clip_obs = float(min(self.clip_obs, big))
        c.clip_actions = float(min(self.clip_actions, big))
        for i in range(2):
            c.cmd_x[i], c.cmd_y[i], c.cmd_yaw[i] = self.command_x_range[i], self.command_y_range[i], self.command_yaw_range[i]
        d0 = self.default_dof_pos[0].tolist()
        for i, v in enumerate(d0):
            c.default_dof_pos[i] = v
        for i, v in enumerate(self.base_init_state):
            c.init_root[i] = v
        c.base_body = int(self.base_index)
        knees = self.knee_indices.tolist()
        c.n_knee = len(knees)
        for i, k in enumerate(knees):
            c.knee_bodies[i] = int(k)
        c.max_episode_length = int(self.max_episode_length)
        c.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        return c

    def _create_fused_task(self):
        self._lib = _lib.load()
        cfg = self._fused_cfg()
        _lib.check(self._lib.b2g_task_anymal_create(self.sim.handle, C.byref(cfg)), "task create")
        self.obs_buf = self._task_tensor(_abi.TT_OBS)
        self.obs_clamped = self._task_tensor(_abi.TT_OBS_CLAMPED)
        self.rew_buf = self._task_tensor(_abi.TT_REW)
        self.reset_buf = self._task_tensor(_abi.TT_RESET)
        self.progress_buf = self._task_tensor(_abi.TT_PROGRESS)
        self.timeout_buf = self._task_tensor(_abi.TT_TIMEOUT)
        self.commands = self._task_tensor(_abi.TT_COMMANDS)
        self.actions = self._task_tensor(_abi.TT_ACTIONS)
        self.rand_override = self._task_tensor(_abi.TT_RAND_OVERRIDE)

    def set_reset_draws(self, draws):
        """Test hook: (N, 27) uniforms in [0,1) used by the next resets instead of the Philox stream; None = Philox."""
        if draws is None:
            _lib.check(self._lib.b2g_task_set_rand_override(self.sim.handle, 0))
        else:
            self.rand_override.copy_(draws.to(self.device, torch.float32))
            _lib.check(self._lib.b2g_task_set_rand_override(self.sim.handle, 1))

    def step(self, actions: torch.Tensor):
        if not self.fused:
            return super().step(actions)
        a = actions.to(self.device, torch.float32)
        if self.randomize:      # vec_task.py:370-372 action noise; the envs flagged now are the ones the kernel resets this step
            if self.dr_randomizations.get("actions", None):
                a = self.dr_randomizations["actions"]["noise_lambda"](a)
            resetting = self.reset_buf != 0
        if not a.is_contiguous():
            a = a.contiguous()
        # keep a reference until the stream has consumed it
        self._last_actions_in = a
        _lib.check(self._lib.b2g_task_anymal_step(self.sim.handle, C.c_void_p(a.data_ptr()), self.sim.stream()), "step")
        self.control_steps += 1
        self.sim.frame_count += 1
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        if self.randomize:
            # the reference re-randomises inside reset_idx (:280-281), i.e. for the environments reset in this step
            if self._dr.count_steps:
                self.randomize_buf += 1
            self.apply_randomizations(self.randomization_params, reset_mask=resetting)
            obs = self.obs_buf
            if self.dr_randomizations.get("observations", None):      # vec_task.py:396-398, then the clamp of :402
                obs = self.dr_randomizations["observations"]["noise_lambda"](obs)
            self.obs_dict["obs"] = torch.clamp(obs, -self.clip_obs, self.clip_obs).to(self.rl_device)
        else:
            self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def reset(self):
        if self.fused:
            self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
            return self.obs_dict
        return super().reset()

    # ------------------------------------------------------------------ generic path (reference hook structure)
    def pre_physics_step(self, actions):
        self.actions = actions.clone().to(self.device)
        targets = self.action_scale * self.actions + self.default_dof_pos
        self._targets = targets.contiguous()
        self.gym.set_dof_position_target_tensor(self.sim, gymtorch.unwrap_tensor(self._targets))

    def post_physics_step(self):
        self.progress_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids) > 0:
            self.reset_idx(env_ids)
        self.compute_observations()
        self.compute_reward(self.actions)

    def compute_reward(self, actions):
        base_quat = self.root_states[:, 3:7]
        lin = quat_rotate_inverse(base_quat, self.root_states[:, 7:10])
        ang = quat_rotate_inverse(base_quat, self.root_states[:, 10:13])
        lin_err = torch.sum(torch.square(self.commands[:, :2] - lin[:, :2]), dim=1)
        ang_err = torch.square(self.commands[:, 2] - ang[:, 2])
        rew = torch.exp(-lin_err / 0.25) * self.rew_scales["lin_vel_xy"] + torch.exp(-ang_err / 0.25) * self.rew_scales["ang_vel_z"] \
            + torch.sum(torch.square(self.torques), dim=1) * self.rew_scales["torque"]
        reset = torch.norm(self.contact_forces[:, self.base_index, :], dim=1) > 1.0
        reset = reset | torch.any(torch.norm(self.contact_forces[:, self.knee_indices, :], dim=2) > 1.0, dim=1)
        reset = reset | (self.progress_buf >= self.max_episode_length - 1)
        self.rew_buf[:] = torch.clip(rew, 0.0, None)
        self.reset_buf[:] = reset

    def compute_observations(self):
        self.gym.refresh_dof_state_tensor(self.sim)
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_net_contact_force_tensor(self.sim)
        self.gym.refresh_dof_force_tensor(self.sim)
        base_quat = self.root_states[:, 3:7]
        scale = torch.tensor([self.lin_vel_scale, self.lin_vel_scale, self.ang_vel_scale], device=self.device)
        self.obs_buf[:] = torch.cat((quat_rotate_inverse(base_quat, self.root_states[:, 7:10]) * self.lin_vel_scale,
                                     quat_rotate_inverse(base_quat, self.root_states[:, 10:13]) * self.ang_vel_scale,
                                     quat_rotate(base_quat, self.gravity_vec),
                                     self.commands * scale,
                                     (self.dof_pos - self.default_dof_pos) * self.dof_pos_scale,
                                     self.dof_vel * self.dof_vel_scale,
                                     self.actions), dim=-1)

    def reset_idx(self, env_ids):
        """Reference :278-304.  On the fused path the constructor's reset of all envs runs as a kernel; later
        resets happen inside the step kernel."""
        if self.fused:
            if len(env_ids) != self.num_envs:
                raise NotImplementedError("fused task: partial resets happen inside step(); use fusedStep=false for manual resets")
            _lib.check(self._lib.b2g_task_anymal_reset_all(self.sim.handle, self.sim.stream()), "reset_all")
            return
        if self.randomize:      # reference :280-281
            self.apply_randomizations(self.randomization_params)
        positions_offset = torch_rand_float(0.5, 1.5, (len(env_ids), self.num_dof), device=self.device)
        velocities = torch_rand_float(-0.1, 0.1, (len(env_ids), self.num_dof), device=self.device)
        self.dof_pos[env_ids] = self.default_dof_pos[env_ids] * positions_offset
        self.dof_vel[env_ids] = velocities
        env_ids_int32 = env_ids.to(dtype=torch.int32)
        self.gym.set_actor_root_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.initial_root_states),
                                                     gymtorch.unwrap_tensor(env_ids_int32), len(env_ids_int32))
        self.gym.set_dof_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.dof_state),
                                              gymtorch.unwrap_tensor(env_ids_int32), len(env_ids_int32))
        self.commands_x[env_ids] = torch_rand_float(self.command_x_range[0], self.command_x_range[1], (len(env_ids), 1), device=self.device).squeeze()
        self.commands_y[env_ids] = torch_rand_float(self.command_y_range[0], self.command_y_range[1], (len(env_ids), 1), device=self.device).squeeze()
        self.commands_yaw[env_ids] = torch_rand_float(self.command_yaw_range[0], self.command_yaw_range[1], (len(env_ids), 1), device=self.device).squeeze()
        self.progress_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
