"""Hound quadruped + 6-DOF manipulator arm, drop-in for the reference's ``tasks/useful_hound.py`` (``UsefulHound`` :77-760): the
HoundTerrain task plus an arm driven by an operational-space torque law (``_compute_osc_torques`` :660-691) every decimation
step, 18 actions, 204 observations (:482-497), arm joints re-drawn on reset (:594-601), knee + shoulder contacts in both the
termination (:467-473) and the collision penalty (:524-525).

Reference quirks kept on purpose: the Jacobian slice ``jacobian[:, joint6, :, :6]`` of a floating-base actor is
the base's six columns; the end-effector state tensor is never refreshed (``env.refreshEefState: false`` keeps it at zero,
``true`` makes it live); the arm command is never sampled; the arm reward is unused.  The mass-matrix block and the Jacobian
slice the OSC law reads are those of the previous ``post_physics_step`` (refreshed before resets).
"""
from __future__ import annotations

import torch

from .. import _abi
from ..utils.torch_math import torch_rand_float
from .hound_terrain import HoundTerrain


class UsefulHound(HoundTerrain):
    ACTOR_NAME = "UsefulHound"
    BASE_HEIGHT_TARGET = 0.52

    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture=False, force_render=False):
        self.arm_action_scale = cfg["env"]["control"]["houndarmactionScale"]
        self.houndarm_dof_noise = cfg["env"]["houndarmDofNoise"]
        self.arm_control_type = cfg["env"]["houndarmcontrolType"]
        rng = cfg["env"]["randomArmCommandPositionRanges"]
        self.arm_command_x_range, self.arm_command_y_range, self.arm_command_z_range = rng["x"], rng["y"], rng["z"]
        super().__init__(cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture, force_render)

    def _get_noise_scale_vec(self, cfg):
        v = torch.zeros(self.num_obs, device=self.device)
        v[:188] = super()._get_noise_scale_vec(cfg)[:188]
        return v

    def _create_envs(self, num_envs, spacing, num_per_row):
        super()._create_envs(num_envs, spacing, num_per_row)
        self.total_num_dof = self.num_dof
        self.hound_num_dof = self.total_num_dof - 6
        self.arm_num_dof = 6
        env, actor = self.envs[0], self.anymal_handles[0]
        self.eef_index = self.gym.find_actor_rigid_body_handle(env, actor, self.cfg["env"]["urdfAsset"]["endpointName"])
        self.hand_joint_index = self.gym.get_actor_joint_dict(env, actor)["joint6"]
        props = self.gym.get_actor_dof_properties(env, actor)
        self.houndarm_dof_lower_limits = torch.tensor(props["lower"][12:].astype("float32").copy(), device=self.device)
        self.houndarm_dof_upper_limits = torch.tensor(props["upper"][12:].astype("float32").copy(), device=self.device)
        self._houndarm_effort_limits = torch.tensor(props["effort"][12:].astype("float32").copy(), device=self.device)

    def _fused_cfg(self) -> _abi.TerrainCfg:
        c = super()._fused_cfg()
        c.n_ctrl_dof, c.arm_chain = self.hound_num_dof, 4
        c.arm_kp, c.arm_kp_null = 150.0, 10.0
        c.arm_action_scale, c.arm_dof_noise = float(self.arm_action_scale), float(self.houndarm_dof_noise)
        for i, v in enumerate([0.1, 0.1, 0.1, 0.5, 0.5, 0.5]):
            c.arm_cmd_limit[i] = v
        c.eef_body, c.jac_body = int(self.eef_index), int(self.hand_joint_index)
        c.refresh_eef = 1 if self.cfg["env"].get("refreshEefState", False) else 0
        return c

    def _create_fused_task(self):
        super()._create_fused_task()
        t = self._task_tensor
        self._mm, self._j_eef, self._eef_state = t(_abi.TT_ARM_MM), t(_abi.TT_ARM_JAC), t(_abi.TT_EEF_STATE)
        self.arm_commands = t(_abi.TT_ARM_COMMANDS)
        self.hound_dof_pos, self.hound_dof_vel = self.dof_pos[:, :12], self.dof_vel[:, :12]
        self._q, self._qd = self.dof_pos[:, 12:], self.dof_vel[:, 12:]
        self.last_hound_dof_vel = self.last_dof_vel[:, :12]
        self.hound_default_dof_pos = self.default_dof_pos[:, :12]
        self.houndarm_default_dof_pos = torch.zeros(6, device=self.device)
        self.arm_kp = torch.full((6,), 150.0, device=self.device)
        self.arm_kd = 2 * torch.sqrt(self.arm_kp)
        self.arm_kp_null = torch.full((6,), 10.0, device=self.device)
        self.arm_kd_null = 2 * torch.sqrt(self.arm_kp_null)
        self.arm_cmd_limit = torch.tensor([[0.1, 0.1, 0.1, 0.5, 0.5, 0.5]], device=self.device)

    def _default_angle(self, name):
        return self.named_default_joint_angles.get(name, 0.0)

    def reset_idx(self, env_ids):
        """Reference :569-637 (constructor-time reset of all envs; in-kernel afterwards)."""
        super().reset_idx(env_ids)
        n = len(env_ids)
        noise = torch.rand((n, 6), device=self.device)
        pos = torch.max(torch.min(self.houndarm_default_dof_pos.unsqueeze(0) + self.houndarm_dof_noise * 2.0 * (noise - 0.5),
                                  self.houndarm_dof_upper_limits), self.houndarm_dof_lower_limits.unsqueeze(0))
        self.dof_pos[env_ids, 12:] = pos
        self.dof_vel[env_ids, 12:] = 0.0
