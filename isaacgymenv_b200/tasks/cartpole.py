"""Cartpole task, drop-in for the reference's ``tasks/cartpole.py`` (``Cartpole`` :36-175): 1 action (cart force =
action * maxEffort, :159-163), 4 observations [x, xdot, theta, thetadot] (:131-142), reward / reset (:180-196),
reset draws (:144-158), episode length 500.

``fusedStep`` (default) runs ``step()`` as one launch of ``k_cartpole_step``; otherwise the reference's hook structure
runs on the gym tensor API.  The reference's ``sim_device=cpu`` configuration of this task cannot be served: this
engine has no CPU path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _abi, _lib, gymapi, gymtorch
from .anymal import default_asset_root
from .base.vec_task import VecTask


class Cartpole(VecTask):
    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture=False, force_render=False):
        self.cfg = cfg
        self.reset_dist = cfg["env"]["resetDist"]
        self.max_push_effort = cfg["env"]["maxEffort"]
        self.max_episode_length = 500
        self.cfg["env"]["numObservations"] = 4
        self.cfg["env"]["numActions"] = 1
        self.fused = bool(cfg["env"].get("fusedStep", True))
        self.needs_host_sync = (not self.fused) or bool(cfg.get("task", {}).get("randomize", False))      # see tasks/anymal.py
        self.seed = int(cfg.get("seed", 42))
        super().__init__(config=self.cfg, rl_device=rl_device, sim_device=sim_device, graphics_device_id=graphics_device_id,
                         headless=headless, virtual_screen_capture=virtual_screen_capture, force_render=force_render)
        self.dof_state = gymtorch.wrap_tensor(self.gym.acquire_dof_state_tensor(self.sim))
        self.dof_pos = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 0]
        self.dof_vel = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 1]
        if self.fused:
            self._create_fused_task()

    def create_sim(self):
        self.up_axis = self.cfg["sim"]["up_axis"]
        self.sim = super().create_sim(self.device_id, self.graphics_device_id, self.physics_engine, self.sim_params)
        plane = gymapi.PlaneParams()
        plane.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        self.gym.add_ground(self.sim, plane)
        self._create_envs(self.num_envs, self.cfg["env"]["envSpacing"], int(np.sqrt(self.num_envs)))

    def _create_envs(self, num_envs, spacing, num_per_row):
        lower = gymapi.Vec3(0.5 * -spacing, -spacing, 0.0)
        upper = gymapi.Vec3(0.5 * spacing, spacing, spacing)
        asset_root = self.cfg["env"].get("assetRoot", default_asset_root())
        asset_file = "urdf/cartpole.urdf"
        if "asset" in self.cfg["env"]:
            asset_file = self.cfg["env"]["asset"].get("assetFileName", asset_file)
        options = gymapi.AssetOptions()
        options.fix_base_link = True
        asset = self.gym.load_asset(self.sim, asset_root, asset_file, options)
        self.num_dof = self.gym.get_asset_dof_count(asset)
        pose = gymapi.Transform()
        pose.p.z = 2.0
        pose.r = gymapi.Quat(0.0, 0.0, 0.0, 1.0)
        self.envs, self.cartpole_handles = [], []
        for i in range(num_envs):
            env = self.gym.create_env(self.sim, lower, upper, num_per_row)
            handle = self.gym.create_actor(env, asset, pose, "cartpole", i, 1, 0)
            self.envs.append(env)
            self.cartpole_handles.append(handle)
        props = self.gym.get_actor_dof_properties(self.envs[0], self.cartpole_handles[0])
        props["driveMode"][0] = gymapi.DOF_MODE_EFFORT
        props["driveMode"][1] = gymapi.DOF_MODE_NONE
        props["stiffness"][:] = 0.0
        props["damping"][:] = 0.0
        self.gym.set_actor_dof_properties(self.envs[0], self.cartpole_handles[0], props)

    # ---- fused path ----
    def _task_tensor(self, kind):
        d = _abi.TensorDesc()
        _lib.check(self._lib.b2g_task_tensor(self.sim.handle, kind, C.byref(d)), "task tensor")
        return _lib.desc_to_torch(d)

    def _create_fused_task(self):
        self._lib = _lib.load()
        big = 3.0e38
        cfg = _abi.CartpoleCfg(reset_dist=self.reset_dist, max_push_effort=self.max_push_effort, clip_obs=float(min(self.clip_obs, big)),
                               clip_actions=float(min(self.clip_actions, big)), max_episode_length=int(self.max_episode_length),
                               seed=int(self.seed) & 0xFFFFFFFFFFFFFFFF)
        _lib.check(self._lib.b2g_task_cartpole_create(self.sim.handle, C.byref(cfg)), "task create")
        self.obs_buf = self._task_tensor(_abi.TT_OBS)
        self.obs_clamped = self._task_tensor(_abi.TT_OBS_CLAMPED)
        self.rew_buf = self._task_tensor(_abi.TT_REW)
        self.reset_buf = self._task_tensor(_abi.TT_RESET)
        self.progress_buf = self._task_tensor(_abi.TT_PROGRESS)
        self.timeout_buf = self._task_tensor(_abi.TT_TIMEOUT)
        self.rand_override = self._task_tensor(_abi.TT_RAND_OVERRIDE)

    def step(self, actions):
        if not self.fused:
            return super().step(actions)
        a = actions.to(self.device, torch.float32).reshape(self.num_envs, 1).contiguous()
        self._last_actions_in = a
        _lib.check(self._lib.b2g_task_step(self.sim.handle, C.c_void_p(a.data_ptr()), self.sim.stream()), "step")
        self.control_steps += 1
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def reset(self):
        if self.fused:
            self.obs_dict["obs"] = self.obs_clamped.to(self.rl_device)
            return self.obs_dict
        return super().reset()

    # ---- generic path (reference hook structure) ----
    def compute_reward(self):
        pole_angle, pole_vel = self.obs_buf[:, 2], self.obs_buf[:, 3]
        cart_vel, cart_pos = self.obs_buf[:, 1], self.obs_buf[:, 0]
        reward = 1.0 - pole_angle * pole_angle - 0.01 * torch.abs(cart_vel) - 0.005 * torch.abs(pole_vel)
        out_x = torch.abs(cart_pos) > self.reset_dist
        out_a = torch.abs(pole_angle) > np.pi / 2
        reward = torch.where(out_x | out_a, torch.full_like(reward, -2.0), reward)
        reset = torch.where(out_x | out_a | (self.progress_buf >= self.max_episode_length - 1), torch.ones_like(self.reset_buf), self.reset_buf)
        self.rew_buf[:], self.reset_buf[:] = reward, reset

    def compute_observations(self, env_ids=None):
        self.gym.refresh_dof_state_tensor(self.sim)
        self.obs_buf[:, 0] = self.dof_pos[:, 0]
        self.obs_buf[:, 1] = self.dof_vel[:, 0]
        self.obs_buf[:, 2] = self.dof_pos[:, 1]
        self.obs_buf[:, 3] = self.dof_vel[:, 1]
        return self.obs_buf

    def reset_idx(self, env_ids):
        positions = 0.2 * (torch.rand((len(env_ids), self.num_dof), device=self.device) - 0.5)
        velocities = 0.5 * (torch.rand((len(env_ids), self.num_dof), device=self.device) - 0.5)
        self.dof_pos[env_ids, :] = positions[:]
        self.dof_vel[env_ids, :] = velocities[:]
        env_ids_int32 = env_ids.to(dtype=torch.int32)
        self.gym.set_dof_state_tensor_indexed(self.sim, gymtorch.unwrap_tensor(self.dof_state), gymtorch.unwrap_tensor(env_ids_int32), len(env_ids_int32))
        self.reset_buf[env_ids] = 0
        self.progress_buf[env_ids] = 0

    def pre_physics_step(self, actions):
        forces = torch.zeros(self.num_envs * self.num_dof, device=self.device, dtype=torch.float)
        forces[::self.num_dof] = actions.to(self.device).squeeze() * self.max_push_effort
        self._forces = forces
        self.gym.set_dof_actuation_force_tensor(self.sim, gymtorch.unwrap_tensor(forces))

    def post_physics_step(self):
        self.progress_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids) > 0:
            self.reset_idx(env_ids)
        self.compute_observations()
        self.compute_reward()
