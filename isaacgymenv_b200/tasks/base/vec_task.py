"""Vectorised-task base classes with the public surface of the reference's ``tasks/base/vec_task.py``
(``Env`` :68-205, ``VecTask`` :208-455, sim-parameter parsing :514-562), re-implemented on top of the
B200 gym shim (``isaacgymenv_b200.gymapi`` / ``gymtorch``).

What a learner (rl_games' ``RLGPUEnv``, ``utils/rlgames_utils.py:242-295``) relies on is kept:
``step(actions) -> (obs_dict, rew, reset, extras)``, ``reset()``, ``reset_done()``, ``obs_buf / rew_buf /
reset_buf / progress_buf / timeout_buf``, ``extras["time_outs"]``, the ``*_space`` properties and the
constructor signature.  Rendering, the viewer and domain randomisation are outside the hot path and
are not provided (``headless=True`` only).  ``task.randomize`` is served by the tensorised ``utils/domain_rand.py``.
"""
from __future__ import annotations

import abc
from typing import Any, Dict, Tuple

import numpy as np
import torch

from ... import gymapi
from ... import spaces


class Env(abc.ABC):
    def __init__(self, config: Dict[str, Any], rl_device: str, sim_device: str, graphics_device_id: int, headless: bool):
        """Device selection and spaces (reference :68-132)."""
        parts = sim_device.split(":")
        self.device_type = parts[0]
        self.device_id = int(parts[1]) if len(parts) > 1 else 0
        self.device = "cpu"
        if config["sim"]["use_gpu_pipeline"]:
            if self.device_type.lower() in ("cuda", "gpu"):
                self.device = f"cuda:{self.device_id}"
            else:
                print("GPU Pipeline can only be used with GPU simulation. Forcing CPU Pipeline.")
                config["sim"]["use_gpu_pipeline"] = False
        self.rl_device = rl_device
        self.headless = headless
        self.graphics_device_id = graphics_device_id
        if not config.get("enableCameraSensors", False) and self.headless:
            self.graphics_device_id = -1

        env_cfg = config["env"]
        self.num_environments = env_cfg["numEnvs"]
        self.num_agents = env_cfg.get("numAgents", 1)
        self.num_observations = env_cfg.get("numObservations", 0)
        self.num_states = env_cfg.get("numStates", 0)
        self.num_actions = env_cfg["numActions"]
        self.control_freq_inv = env_cfg.get("controlFrequencyInv", 1)
        self.obs_space = spaces.Box(np.full(self.num_obs, -np.inf), np.full(self.num_obs, np.inf))
        self.state_space = spaces.Box(np.full(self.num_states, -np.inf), np.full(self.num_states, np.inf))
        self.act_space = spaces.Box(np.full(self.num_actions, -1.0), np.full(self.num_actions, 1.0))
        self.clip_obs = env_cfg.get("clipObservations", np.inf)
        self.clip_actions = env_cfg.get("clipActions", np.inf)
        self.total_train_env_frames: int = 0
        self.control_steps: int = 0
        self.render_fps: int = env_cfg.get("renderFPS", -1)
        self.last_frame_time: float = 0.0
        self.record_frames: bool = False

    @abc.abstractmethod
    def allocate_buffers(self):
        """Create obs_buf, rew_buf, reset_buf, ... on the sim device."""

    @abc.abstractmethod
    def step(self, actions: torch.Tensor) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor, Dict[str, Any]]:
        """Advance every environment by one policy step."""

    @abc.abstractmethod
    def reset(self) -> Dict[str, torch.Tensor]:
        """Return the current observations (the reference does not re-simulate here)."""

    @abc.abstractmethod
    def reset_idx(self, env_ids: torch.Tensor):
        """Reset the given environments."""

    @property
    def observation_space(self):
        return self.obs_space

    @property
    def action_space(self):
        return self.act_space

    @property
    def num_envs(self) -> int:
        return self.num_environments

    @property
    def num_acts(self) -> int:
        return self.num_actions

    @property
    def num_obs(self) -> int:
        return self.num_observations

    def set_train_info(self, env_frames, *args, **kwargs):
        """Learner -> env: total frames so far (reference :187-194)."""
        self.total_train_env_frames = env_frames

    def get_env_state(self):
        """Checkpoint hook; the hot-path tasks keep no learner-visible state (reference :196-200)."""
        return None

    def set_env_state(self, env_state):
        pass


class VecTask(Env):
    CONTACT_SLOTS = 0      # 0 = library default (4); robots whose links rest on more candidates at once ask for more (Hound: 6)

    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 24}

    def __init__(self, config, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture: bool = False,
                 force_render: bool = False):
        super().__init__(config, rl_device, sim_device, graphics_device_id, headless)
        if not hasattr(self, "cfg"):
            self.cfg = config
        self.virtual_screen_capture = virtual_screen_capture
        self.virtual_display = None
        self.force_render = force_render
        self.sim_params = self._parse_sim_params(self.cfg["physics_engine"], self.cfg["sim"])
        # contact slots per solver lane (b2g_sim_params.max_contacts_per_chain): a yaml value under sim.physx wins, else the task's own
        if self.CONTACT_SLOTS and not getattr(self.sim_params.physx, "max_contacts_per_chain", 0):
            self.sim_params.physx.max_contacts_per_chain = int(self.CONTACT_SLOTS)
        if self.cfg["physics_engine"] == "physx":
            self.physics_engine = gymapi.SIM_PHYSX
        elif self.cfg["physics_engine"] == "flex":
            self.physics_engine = gymapi.SIM_FLEX
        else:
            raise ValueError(f"Invalid physics engine backend: {self.cfg['physics_engine']}")
        self.dt: float = self.sim_params.dt
        self.gym = gymapi.acquire_gym()
        self.first_randomization = True
        self.original_props = {}
        self.dr_randomizations = {}
        self.actor_params_generator = None
        self.extern_actor_params = {i: None for i in range(self.num_envs)}
        self.last_step = -1
        self.last_rand_step = -1
        self.sim_initialized = False
        self.create_sim()
        self.gym.prepare_sim(self.sim)
        self.sim_initialized = True
        self.set_viewer()
        self.allocate_buffers()
        self.obs_dict = {}

    # ---- viewer: headless only ----
    def set_viewer(self):
        self.enable_viewer_sync = True
        self.viewer = None
        if not self.headless:
            raise NotImplementedError("rendering is not part of the B200 hot path: create the task with headless=True")

    def render(self, mode="rgb_array"):
        return None

    def allocate_buffers(self):
        """Reference :301-324: int64 counters, reset_buf starts at ones."""
        n, dev = self.num_envs, self.device
        self.obs_buf = torch.zeros((n, self.num_obs), device=dev, dtype=torch.float)
        self.states_buf = torch.zeros((n, self.num_states), device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(n, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(n, device=dev, dtype=torch.long)
        self.timeout_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.progress_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.randomize_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.extras = {}

    def create_sim(self, compute_device: int, graphics_device: int, physics_engine, sim_params: gymapi.SimParams):
        """Reference :326-342.  One sim per task instance (the reference keeps a process-global singleton because
        Isaac Gym cannot host two sims; this engine can)."""
        sim = self.gym.create_sim(compute_device, graphics_device, physics_engine, sim_params)
        if sim is None:
            raise RuntimeError("*** Failed to create sim")
        return sim

    def get_state(self):
        return torch.clamp(self.states_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)

    @abc.abstractmethod
    def pre_physics_step(self, actions: torch.Tensor):
        """Apply the actions (position targets / torques)."""

    @abc.abstractmethod
    def post_physics_step(self):
        """Resets, observations, rewards."""

    def step(self, actions: torch.Tensor):
        """Reference :360-408 (the generic, un-fused path: hooks + gym.simulate)."""
        if self.dr_randomizations.get("actions", None):
            actions = self.dr_randomizations["actions"]["noise_lambda"](actions)
        action_tensor = torch.clamp(actions, -self.clip_actions, self.clip_actions)
        self.pre_physics_step(action_tensor)
        for _ in range(self.control_freq_inv):
            self.gym.simulate(self.sim)
        if self.device == "cpu":
            self.gym.fetch_results(self.sim, True)
        self.post_physics_step()
        self.control_steps += 1
        if getattr(self, "_dr", None) is not None and self._dr.count_steps:
            self.randomize_buf += 1
        self.timeout_buf = (self.progress_buf >= self.max_episode_length - 1) & (self.reset_buf != 0)
        if self.dr_randomizations.get("observations", None):
            self.obs_buf = self.dr_randomizations["observations"]["noise_lambda"](self.obs_buf)
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def zero_actions(self) -> torch.Tensor:
        return torch.zeros([self.num_envs, self.num_actions], dtype=torch.float32, device=self.rl_device)

    def reset_idx(self, env_idx):
        pass

    def reset(self):
        """Reference :426-438: returns the clamped current obs_buf; no physics."""
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict

    def reset_done(self):
        """Reference :440-455."""
        done_env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        if len(done_env_ids) > 0:
            self.reset_idx(done_env_ids)
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, done_env_ids

    def _parse_sim_params(self, physics_engine: str, config_sim: Dict[str, Any]) -> gymapi.SimParams:
        """yaml ``sim:`` block -> SimParams (reference :514-562)."""
        sp = gymapi.SimParams()
        if config_sim["up_axis"] not in ("z", "y"):
            raise ValueError(f"Invalid physics up-axis: {config_sim['up_axis']}")
        sp.dt = config_sim["dt"]
        sp.num_client_threads = config_sim.get("num_client_threads", 0)
        sp.use_gpu_pipeline = config_sim["use_gpu_pipeline"]
        sp.substeps = config_sim.get("substeps", 2)
        sp.up_axis = gymapi.UP_AXIS_Z if config_sim["up_axis"] == "z" else gymapi.UP_AXIS_Y
        sp.gravity = gymapi.Vec3(*config_sim["gravity"])
        if physics_engine == "physx":
            for opt, val in config_sim.get("physx", {}).items():
                if opt == "contact_collection":
                    val = gymapi.ContactCollection(val)
                setattr(sp.physx, opt, val)
        else:
            for opt, val in config_sim.get("flex", {}).items():
                setattr(sp.flex, opt, val)
        return sp

    # kept under the reference's (name-mangled) private name too
    _VecTask__parse_sim_params = _parse_sim_params

    def apply_randomizations(self, dr_params, reset_mask=None):
        """Reference :610-840, tensorised (``utils/domain_rand.py``): per-environment mass / drive-gain / friction tensors
        read by the step kernels, gravity, observation and action noise.  No per-environment Python loop, no host sync."""
        if getattr(self, "_dr", None) is None:
            from ...utils.domain_rand import DomainRandomizer

            self._dr = DomainRandomizer(self, count_steps=bool(dr_params.get("count_steps", False)))
        self._dr.apply(dr_params, reset_mask)
        self.first_randomization = False
