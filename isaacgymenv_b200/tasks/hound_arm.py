"""Houndarm / Manipulator: fixed-base 6-DOF (7-DOF Franka: class ``Manipulator`` at the end) arm reach task, drop-in for the reference's ``tasks/hound_arm.py`` (``Houndarm`` :74-545):
6 actions = end-effector pose deltas scaled by ``cmd_limit`` and turned into joint torques by the operational-space law
(:462-493) -- or 6 raw joint torques with ``controlType: joint_tor`` --, 10 observations [eef position, eef quaternion,
commanded position] (:383-392), reward ``compute_houndarm_reward`` (:550-567), reset draws (:394-459), 150-step episodes.

This is SURVEY 8(f) row 4.  Two execution paths, selected by ``cfg["env"]["fusedStep"]`` (default True):
* fused: ``step()`` is ONE launch of ``k_houndarm_step`` (one thread per environment: OSC torques from the current mass matrix /
  Jacobian row / end-effector velocity, the sub-steps, resets, observations, reward, time-outs) -- graph-capturable;
* generic: the reference's hook structure on the gym-tensor API of the shim (``k_simulate`` + ``k_body_state`` + ``k_jacobian`` +
  ``k_mass_matrix``; OSC and reward as torch ops).
Replicated as they are: the Jacobian row is taken at the JOINT index (``get_actor_joint_dict()['joint6']`` = 5, which in the
fixed-base Jacobian -- root body left out -- is body 6, ``end_link``; :314-318), ``reset_buf[env_ids] = 0`` on reset (quirk Q5),
the no-op ``u_null[:, 6:] *= 0``.  ``asset_options.disable_gravity`` is honoured by the sim.  On the fused path ``self.states`` is
refreshed on demand (``_refresh()``), not every step."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _abi, _lib, gymapi, gymtorch
from ..utils.torch_math import tensor_clamp, to_torch, torch_rand_float
from .anymal import default_asset_root
from .base.vec_task import VecTask


def compute_houndarm_reward(reset_buf, progress_buf, eef_pos, eef_vel, commands, dist_scale: float, vel_scale: float, max_episode_length: float):
    """Reference :550-567.  Distance term 1 - tanh(10 d); a velocity term that only counts within 2 cm of the target."""
    d = torch.norm(eef_pos - commands, dim=-1)
    in_reach = d < 0.02
    rewards = (1.0 - torch.tanh(10.0 * d)) * dist_scale + (1.0 - torch.tanh(10.0 * torch.norm(eef_vel, dim=-1))) * in_reach * vel_scale
    rewards = torch.clip(rewards, 0.0, None)
    reset = torch.where(progress_buf >= max_episode_length - 1, torch.ones_like(reset_buf), reset_buf)
    return rewards, reset


def osc_torques(mm, j_eef, dpose, eef_vel, q, qd, kp, kd, kp_null, kd_null, default_q, effort_limits):
    """Reference :462-493: Khatib's operational-space law with a null-space posture term, clamped to the effort limits."""
    mm_inv = torch.inverse(mm)
    j_t = j_eef.transpose(1, 2)
    m_eef = torch.inverse(j_eef @ mm_inv @ j_t)
    u = j_t @ m_eef @ (kp * dpose - kd * eef_vel).unsqueeze(-1)
    j_eef_inv = m_eef @ j_eef @ mm_inv
    u_null = kd_null * -qd + kp_null * ((default_q - q + np.pi) % (2 * np.pi) - np.pi)
    u_null = mm @ u_null.unsqueeze(-1)
    u = u + (torch.eye(mm.shape[-1], device=mm.device).unsqueeze(0) - j_t @ j_eef_inv) @ u_null
    return tensor_clamp(u.squeeze(-1), -effort_limits.unsqueeze(0), effort_limits.unsqueeze(0))


class Houndarm(VecTask):
    # what distinguishes the reference's two arm-reach task files (hound_arm.py is manipulator.py with these changed); the attribute names
    # the reference spells with the arm's name (``houndarm_dof_noise`` / ``franka_dof_noise`` ...) exist under both spellings' own prefix
    ARM = "houndarm"
    N_ARM_DOFS = 6
    DEFAULT_DOF_POS = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)                     # hound_arm.py:160-162
    ASSET_FILE = "urdf/open_manipulator_p_gazebo/urdf/open_manipulator_p.urdf"
    ASSET_KEY = "assetFileNamehoundarm"
    EEF_LINK, EEF_JOINT = "end_link", "joint6"
    RESET_TAIL = 0                                                         # joints at the end of the chain reset without noise
    FLIP_VISUAL = False

    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture=False, force_render=False):
        self.cfg = cfg
        env = cfg["env"]
        self.max_episode_length = env["episodeLength"]
        self.action_scale = env["actionScale"]
        self.arm_dof_noise = env[f"{self.ARM}DofNoise"]
        setattr(self, f"{self.ARM}_dof_noise", self.arm_dof_noise)
        rng = env["randomCommandPositionRanges"]
        self.command_x_range, self.command_y_range, self.command_z_range = rng["x"], rng["y"], rng["z"]
        self.reward_settings = {"r_dist_scale": env["distRewardScale"], "r_lift_scale": env["liftRewardScale"], "r_align_scale": env["alignRewardScale"],
                                "r_stack_scale": env["stackRewardScale"], "r_vel_scale": env["velRewardScale"]}
        self.control_type = env["controlType"]
        if self.control_type not in ("osc", "joint_tor"):
            raise ValueError("Invalid control type specified. Must be one of: {osc, joint_tor}")
        cfg["env"]["numObservations"] = 10 if self.control_type == "osc" else 26
        cfg["env"]["numActions"] = 6 if self.control_type == "osc" else 8
        if self.control_type != "osc":
            # the reference's joint_tor variant declares 26 observations / 8 actions but builds 10 / uses 6 (:383-392, :499-501)
            raise NotImplementedError("controlType joint_tor is inconsistent in the reference (26 obs declared, 10 built); only osc is served")
        if cfg.get("task", {}).get("randomize", False):
            raise NotImplementedError(f"{type(self).__name__}: task.randomize is not wired (the reference task never calls apply_randomizations)")
        self.states, self.handles = {}, {}
        self.fused = bool(env.get("fusedStep", True))
        # the generic path finds the environments to reset with nonzero() (as the reference does): not CUDA-graph capturable
        self.needs_host_sync = not self.fused
        self.up_axis, self.up_axis_idx = "z", 2
        self.seed = int(cfg.get("seed", 42))
        super().__init__(config=cfg, rl_device=rl_device, sim_device=sim_device, graphics_device_id=graphics_device_id, headless=headless,
                         virtual_screen_capture=virtual_screen_capture, force_render=force_render)
        self.arm_default_dof_pos = to_torch(list(self.DEFAULT_DOF_POS), device=self.device)
        setattr(self, f"{self.ARM}_default_dof_pos", self.arm_default_dof_pos)
        self.kp = to_torch([150.0] * 6, device=self.device)
        self.kd = 2 * torch.sqrt(self.kp)
        self.kp_null = to_torch([10.0] * self.N_ARM_DOFS, device=self.device)
        self.kd_null = 2 * torch.sqrt(self.kp_null)
        self.cmd_limit = to_torch([0.1, 0.1, 0.1, 0.5, 0.5, 0.5], device=self.device).unsqueeze(0)
        self.commands = torch.zeros(self.num_envs, 3, dtype=torch.float, device=self.device)
        self.actions = torch.zeros(self.num_envs, self.num_actions, device=self.device)
        if self.fused:
            self._create_fused_task()
        self.commands_x, self.commands_y, self.commands_z = (self.commands.view(self.num_envs, 3)[..., i] for i in range(3))
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        self._refresh()

    # ---- scene ----
    def create_sim(self):
        self.sim_params.up_axis = gymapi.UP_AXIS_Z
        self.sim_params.gravity.x, self.sim_params.gravity.y, self.sim_params.gravity.z = 0.0, 0.0, -9.81
        self.sim = super().create_sim(self.device_id, self.graphics_device_id, self.physics_engine, self.sim_params)
        plane = gymapi.PlaneParams()
        plane.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        self.gym.add_ground(self.sim, plane)
        self._create_envs(self.num_envs, self.cfg["env"]["envSpacing"], int(np.sqrt(self.num_envs)))

    def _create_envs(self, num_envs, spacing, num_per_row):
        lower, upper = gymapi.Vec3(-spacing, -spacing, 0.0), gymapi.Vec3(spacing, spacing, spacing)
        asset_root = self.cfg["env"].get("assetRoot", default_asset_root())
        asset_file = self.ASSET_FILE
        if "asset" in self.cfg["env"]:
            asset_file = self.cfg["env"]["asset"].get(self.ASSET_KEY, asset_file)
        opt = gymapi.AssetOptions()
        opt.replace_cylinder_with_capsule = False
        opt.flip_visual_attachments = self.FLIP_VISUAL
        opt.fix_base_link = True
        opt.collapse_fixed_joints = False
        opt.disable_gravity = True
        opt.thickness = 0.001
        opt.default_dof_drive_mode = gymapi.DOF_MODE_EFFORT
        opt.use_mesh_materials = True
        asset = self.gym.load_asset(self.sim, asset_root, asset_file, opt)
        self.num_arm_bodies = self.gym.get_asset_rigid_body_count(asset)
        self.num_arm_dofs = self.gym.get_asset_dof_count(asset)
        setattr(self, f"num_{self.ARM}_bodies", self.num_arm_bodies)
        setattr(self, f"num_{self.ARM}_dofs", self.num_arm_dofs)
        if self.num_arm_dofs != self.N_ARM_DOFS:
            raise ValueError(f"{type(self).__name__} expects a {self.N_ARM_DOFS}-DOF arm, the asset has {self.num_arm_dofs}")
        props = self.gym.get_asset_dof_properties(asset)
        for i in range(self.num_arm_dofs):
            props["driveMode"][i] = gymapi.DOF_MODE_POS if i > 6 else gymapi.DOF_MODE_EFFORT
            props["stiffness"][i] = 0.0
            props["damping"][i] = 0.0
        self.arm_dof_lower_limits = to_torch(props["lower"].astype("float32").copy(), device=self.device)
        self.arm_dof_upper_limits = to_torch(props["upper"].astype("float32").copy(), device=self.device)
        self._arm_effort_limits = to_torch(props["effort"].astype("float32").copy(), device=self.device)
        setattr(self, f"{self.ARM}_dof_lower_limits", self.arm_dof_lower_limits)
        setattr(self, f"{self.ARM}_dof_upper_limits", self.arm_dof_upper_limits)
        setattr(self, f"_{self.ARM}_effort_limits", self._arm_effort_limits)
        pose = gymapi.Transform()
        pose.p = gymapi.Vec3(-0.45, 0.0, 0.0)
        pose.r = gymapi.Quat(0.0, 0.0, 0.0, 1.0)
        self.envs, self.arms = [], []
        setattr(self, f"{self.ARM}s", self.arms)
        for i in range(num_envs):
            env = self.gym.create_env(self.sim, lower, upper, num_per_row)
            self.arms.append(self.gym.create_actor(env, asset, pose, self.ARM, i, 0, 0))
            self.gym.set_actor_dof_properties(env, self.arms[-1], props)
            self.envs.append(env)
        self.gym.prepare_sim(self.sim)
        self.init_data()

    def init_data(self):
        env, actor = self.envs[0], 0
        self.handles = {"endpoint_tip": self.gym.find_actor_rigid_body_handle(env, actor, self.EEF_LINK)}
        self.num_dofs = self.gym.get_sim_dof_count(self.sim) // self.num_envs
        n = self.num_envs
        self._root_state = gymtorch.wrap_tensor(self.gym.acquire_actor_root_state_tensor(self.sim)).view(n, -1, 13)
        self._dof_state = gymtorch.wrap_tensor(self.gym.acquire_dof_state_tensor(self.sim)).view(n, -1, 2)
        self._rigid_body_state = gymtorch.wrap_tensor(self.gym.acquire_rigid_body_state_tensor(self.sim)).view(n, -1, 13)
        self._q, self._qd = self._dof_state[..., 0], self._dof_state[..., 1]
        self._eef_state = self._rigid_body_state[:, self.handles["endpoint_tip"], :]
        k = self.N_ARM_DOFS
        jacobian = gymtorch.wrap_tensor(self.gym.acquire_jacobian_tensor(self.sim, self.ARM))
        hand_joint_index = self.gym.get_actor_joint_dict(env, actor)[self.EEF_JOINT]
        self._j_eef = jacobian[:, hand_joint_index, :, :k]
        self._mm = gymtorch.wrap_tensor(self.gym.acquire_mass_matrix_tensor(self.sim, self.ARM))[:, :k, :k]
        self._pos_control = torch.zeros((n, self.num_dofs), dtype=torch.float, device=self.device)
        self._effort_control = torch.zeros_like(self._pos_control)
        self._arm_control = self._effort_control[:, :k]
        self._global_indices = torch.arange(n, dtype=torch.int32, device=self.device).view(n, -1)

    # ---- fused path ----
    def _create_fused_task(self):
        self._lib = _lib.load()
        e = self.cfg["env"]
        c = _abi.HoundarmCfg(clip_obs=float(self.clip_obs), clip_actions=float(self.clip_actions), action_scale=float(self.action_scale),
                             dof_noise=float(self.arm_dof_noise), kp=150.0, kp_null=10.0, dist_scale=float(e["distRewardScale"]),
                             vel_scale=float(e["velRewardScale"]), eef_body=int(self.handles["endpoint_tip"]),
                             jac_body=int(self.gym.get_actor_joint_dict(self.envs[0], 0)[self.EEF_JOINT]) + 1,
                             max_episode_length=int(self.max_episode_length), seed=int(self.seed) & 0xFFFFFFFFFFFFFFFF,
                             n_reset_tail=int(self.RESET_TAIL))
        for i, v in enumerate(self.DEFAULT_DOF_POS):
            c.default_dof_pos[i] = float(v)
        for i, v in enumerate(self.cmd_limit.flatten().tolist()):
            c.cmd_limit[i] = v
        for i, v in enumerate(list(self.command_x_range) + list(self.command_y_range) + list(self.command_z_range)):
            c.cmd_range[i] = float(v)
        _lib.check(self._lib.b2g_task_houndarm_create(self.sim.handle, C.byref(c)), "task create")

        def tt(kind):
            d = _abi.TensorDesc()
            _lib.check(self._lib.b2g_task_tensor(self.sim.handle, kind, C.byref(d)), "task tensor")
            return _lib.desc_to_torch(d)

        self.obs_buf, self.obs_clamped, self.rew_buf = tt(_abi.TT_OBS), tt(_abi.TT_OBS_CLAMPED), tt(_abi.TT_REW)
        self.reset_buf, self.progress_buf, self.timeout_buf = tt(_abi.TT_RESET), tt(_abi.TT_PROGRESS), tt(_abi.TT_TIMEOUT)
        self.commands, self.actions = tt(_abi.TT_COMMANDS), tt(_abi.TT_ACTIONS)
        self.reset_buf.zero_()

    def step(self, actions: torch.Tensor):
        if not self.fused:
            return super().step(actions)
        a = actions.to(self.device, torch.float32)
        if not a.is_contiguous():
            a = a.contiguous()
        self._last_actions_in = a          # keep a reference until the stream has consumed it
        _lib.check(self._lib.b2g_task_step(self.sim.handle, C.c_void_p(a.data_ptr()), self.sim.stream()), "step")
        self.control_steps += 1
        self.sim.frame_count += 1
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def reset(self):
        if self.fused:
            self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
            return self.obs_dict
        return super().reset()

    # ---- state ----
    def _update_states(self):
        self.states.update({"q": self._q[:, :], "q_gripper": self._q[:, -2:], "eef_pos": self._eef_state[:, :3], "eef_quat": self._eef_state[:, 3:7],
                            "eef_vel": self._eef_state[:, 7:], "commands": self.commands[:, :]})

    def _refresh(self):
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_dof_state_tensor(self.sim)
        self.gym.refresh_rigid_body_state_tensor(self.sim)
        self.gym.refresh_jacobian_tensors(self.sim)
        self.gym.refresh_mass_matrix_tensors(self.sim)
        self._update_states()

    def compute_reward(self, actions):
        rew, reset = compute_houndarm_reward(self.reset_buf, self.progress_buf, self.states["eef_pos"], self.states["eef_vel"], self.states["commands"],
                                             self.reward_settings["r_dist_scale"], self.reward_settings["r_vel_scale"], self.max_episode_length)
        self.rew_buf[:], self.reset_buf[:] = rew, reset

    def compute_observations(self):
        self._refresh()
        self.obs_buf = torch.cat([self.states[k] for k in ("eef_pos", "eef_quat", "commands")], dim=-1)
        return self.obs_buf

    def reset_idx(self, env_ids):
        k = len(env_ids)
        self.commands_x[env_ids] = torch_rand_float(self.command_x_range[0], self.command_x_range[1], (k, 1), device=self.device).squeeze()
        self.commands_y[env_ids] = torch_rand_float(self.command_y_range[0], self.command_y_range[1], (k, 1), device=self.device).squeeze()
        self.commands_z[env_ids] = torch_rand_float(self.command_z_range[0], self.command_z_range[1], (k, 1), device=self.device).squeeze()
        noise = torch.rand((k, self.N_ARM_DOFS), device=self.device)
        pos = tensor_clamp(self.arm_default_dof_pos.unsqueeze(0) + self.arm_dof_noise * 2.0 * (noise - 0.5),
                           self.arm_dof_lower_limits.unsqueeze(0), self.arm_dof_upper_limits)
        if self.RESET_TAIL:      # manipulator.py:417 ("gripper" positions; on this asset they are the arm's last joints)
            pos[:, -self.RESET_TAIL:] = self.arm_default_dof_pos[-self.RESET_TAIL:]
        self._q[env_ids, :] = pos
        self._qd[env_ids, :] = torch.zeros_like(self._qd[env_ids])
        self._pos_control[env_ids, :] = pos
        self._effort_control[env_ids, :] = torch.zeros_like(pos)
        ids = self._global_indices[env_ids].flatten()
        self.gym.set_dof_position_target_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self._pos_control), gymtorch.unwrap_tensor(ids), len(ids))
        self.gym.set_dof_actuation_force_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self._effort_control), gymtorch.unwrap_tensor(ids), len(ids))
        self.gym.set_dof_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self._dof_state), gymtorch.unwrap_tensor(ids), len(ids))
        self.progress_buf[env_ids] = 0
        self.reset_buf[env_ids] = 0

    # ---- control ----
    def _compute_osc_torques(self, dpose):
        k = self.N_ARM_DOFS
        return osc_torques(self._mm, self._j_eef, dpose, self.states["eef_vel"], self._q[:, :k], self._qd[:, :k], self.kp, self.kd, self.kp_null,
                           self.kd_null, self.arm_default_dof_pos[:k], self._arm_effort_limits[:k])

    def pre_physics_step(self, actions):
        self.actions = actions.clone().to(self.device)
        u_arm = self.actions * self.cmd_limit / self.action_scale
        if self.control_type == "osc":
            u_arm = self._compute_osc_torques(dpose=u_arm)
        self._arm_control[:, :] = u_arm
        self.gym.set_dof_actuation_force_tensor(self.sim, gymtorch.unwrap_tensor(self._effort_control))

    def post_physics_step(self):
        self.progress_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids) > 0:
            self.reset_idx(env_ids)
        self.compute_observations()
        self.compute_reward(self.actions)


class Manipulator(Houndarm):
    """``tasks/manipulator.py`` (``Manipulator`` :75-617): the same reach task on the 7-DOF Franka arm (the asset's hand and fingers sit
    behind the end of its ``<robot>`` element, ``model/urdf.py::_parse_xml_lenient``): six pose-change actions -> seven joint torques
    through the 7 x 7 operational-space law with the null-space posture ``franka_default_dof_pos`` (:153-155), 1000-step episodes, reset
    around that posture with the last two joints set back without noise (:407-417).  Mass properties: the URDF has no <inertial>, Isaac
    Gym derives them from the collision meshes' convex hulls at the default density -- so does the model compiler (18.98 kg)."""
    ARM = "franka"
    N_ARM_DOFS = 7
    DEFAULT_DOF_POS = (0.0, 0.1963, 0.0, -2.6180, 0.0, 2.9416, 0.7854)
    ASSET_FILE = "urdf/franka_description/robots/franka_panda_manipulator.urdf"
    ASSET_KEY = "assetFileNameFranka"
    EEF_LINK, EEF_JOINT = "panda_link7", "panda_joint7"
    RESET_TAIL = 2
    FLIP_VISUAL = True
