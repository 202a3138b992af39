"""Two ways to run the kernel code under test with numpy in / numpy out:

* ``CudaBackend``: the product -- ``libb200gym.so`` through its C ABI on ``cuda:0`` (``-m gpu`` tests);
* ``EmuBackend``: the same per-thread kernel bodies compiled for the host by ``tests/emu`` (CPU tests).

Both expose the same small interface so one test body checks both against the oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from isaacgymenv_b200 import _abi

TASK_KEYS = ("obs", "obs_clamped", "rew", "reset", "progress", "timeout", "commands", "actions")
TERRAIN_KEYS = ("torques", "last_actions", "last_dof_vel", "feet_air_time", "episode_sums", "env_origins", "terrain_levels", "terrain_types", "measured",
                "arm_mm", "arm_jac", "eef_state", "arm_commands")


class EmuBackend:
    name = "emu"

    def __init__(self, art, params, props, n):
        from tests.emu import emu

        self.emu = emu
        self.art, self.params, self.props, self.n = art, params, props, n
        self.model = _abi.pack_model(art)
        nd, nb = art.num_dofs, art.num_bodies
        self.root = np.zeros((n, 13), np.float32)
        self.root[:, 6] = 1
        self.dof = np.zeros((n, nd, 2), np.float32)
        self.dof_force = np.zeros((n, nd), np.float32)
        self.contact = np.zeros((n, nb, 3), np.float32)
        self.task = None

    def set_state(self, root, dof):
        self.root[:] = root
        self.dof[:] = dof

    def get_state(self):
        return self.root.copy(), self.dof.copy()

    def set_env_scale(self, scale, friction=None):
        self._scale = np.ascontiguousarray(scale, np.float32)
        self._fric = None if friction is None else np.ascontiguousarray(friction, np.float32)

    def set_link_scale(self, scale):
        self._lscale = np.ascontiguousarray(scale, np.float32)

    def add_heightfield(self, hf_t, samples):
        self._hf = (hf_t, np.ascontiguousarray(samples, np.int16))

    def set_friction(self, friction):
        self._fric = np.ascontiguousarray(friction, np.float32)

    def contact_stats(self, reset=False):
        return self.emu.contact_stats(reset)

    def simulate(self, target, actuation):
        self.emu.set_env_scale(getattr(self, "_scale", None))
        self.emu.set_link_scale(getattr(self, "_lscale", None))
        hf = getattr(self, "_hf", None)
        try:
            f, c = self.emu.simulate(self.model, self.params, self.props, self.root, self.dof, target, actuation, friction=getattr(self, "_fric", None),
                                     heightfield=hf[0] if hf else None, hf_samples=hf[1] if hf else None)
        finally:
            self.emu.set_env_scale(None)
            self.emu.set_link_scale(None)
        self.dof_force[:], self.contact[:] = f, c
        return f, c

    def forward_dynamics(self, tau):
        self.emu.set_env_scale(getattr(self, "_scale", None))
        self.emu.set_link_scale(getattr(self, "_lscale", None))
        try:
            return self.emu.forward_dynamics(self.model, self.params, self.props, self.root.copy(), self.dof.copy(), tau)
        finally:
            self.emu.set_env_scale(None)
            self.emu.set_link_scale(None)

    def jacobian_mass_matrix(self):
        return self.emu.jacobian_mass_matrix(self.model, self.props, self.root, self.dof)

    # ---- fused flat task ----
    def cartpole_create(self, cfg):
        n = self.n
        self.cfg = cfg
        self.kind = "cartpole"
        self.task = dict(obs=np.zeros((n, 4), np.float32), obs_clamped=np.zeros((n, 4), np.float32), rew=np.zeros(n, np.float32),
                         reset=np.ones(n, np.int64), progress=np.zeros(n, np.int64), timeout=np.zeros(n, np.int64),
                         commands=np.zeros((n, 3), np.float32), actions=np.zeros((n, 1), np.float32), reset_count=np.zeros(n, np.int32))

    def houndarm_create(self, cfg):
        n, nd = self.n, self.art.num_dofs
        self.cfg, self.kind = cfg, "houndarm"
        self.task = dict(obs=np.zeros((n, 10), np.float32), obs_clamped=np.zeros((n, 10), np.float32), rew=np.zeros(n, np.float32),
                         reset=np.ones(n, np.int64), progress=np.zeros(n, np.int64), timeout=np.zeros(n, np.int64),
                         commands=np.zeros((n, 3), np.float32), actions=np.zeros((n, 6), np.float32), reset_count=np.zeros(n, np.int32))

    def task_step(self, actions, draws=None, post_only=False):
        if getattr(self, "kind", "anymal") == "terrain":
            self._terrain_step(actions, draws, post_only)
        elif getattr(self, "kind", "anymal") == "houndarm":
            self.emu.set_env_scale(getattr(self, "_scale", None))
            try:
                self.emu.houndarm(self.model, self.params, self.props, self.cfg, 2 if post_only else 1, self._bufs(), actions, draws)
            finally:
                self.emu.set_env_scale(None)
        elif getattr(self, "kind", "anymal") == "cartpole":
            self.emu.cartpole(self.model, self.params, self.props, self.cfg, 2 if post_only else 1, self._bufs(), actions, draws)
        elif post_only:
            self.anymal_post_only(actions, draws)
        else:
            self.anymal_step(actions, draws)

    def terrain_create(self, cfg, height_samples=None, terrain_origins=None, heightfield=None):
        n, nd = self.n, self.art.num_dofs
        nhp = cfg.n_hx * cfg.n_hy
        arm = cfg.arm_chain >= 0
        nctrl = cfg.n_ctrl_dof if cfg.n_ctrl_dof > 0 else nd
        no = 12 + 2 * nctrl + nhp + nd + (10 if arm else 0)
        self.cfg, self.kind = cfg, "terrain"
        self.hs = None if height_samples is None else np.ascontiguousarray(height_samples, np.int16)
        self.origins = None if terrain_origins is None else np.ascontiguousarray(terrain_origins, np.float32)
        self.heightfield = heightfield
        self.init_done, self.common_step = 1, 0
        z = lambda *sh, dt=np.float32: np.zeros(sh, dt)
        self.task = dict(obs=z(n, no), obs_clamped=z(n, no), rew=z(n), reset=np.ones(n, np.int64), progress=z(n, dt=np.int64), timeout=z(n, dt=np.int64),
                         commands=z(n, 4), actions=z(n, nd), reset_count=z(n, dt=np.int32), torques=z(n, nd), last_actions=z(n, nd), last_dof_vel=z(n, nd),
                         feet_air_time=z(n, 4), episode_sums=z(13, n), env_origins=z(n, 3), terrain_levels=z(n, dt=np.int64), terrain_types=z(n, dt=np.int64),
                         scratch=z(n, 9), resetw=z(n), report=z(13, n), measured=z(n, nhp), arm_mm=z(n, 6, 6), arm_jac=z(n, 6, 6),
                         eef_state=z(n, 13), arm_commands=z(n, 3))

    def _terrain_step(self, actions, draws, post_only):
        b = self._bufs()
        b["actions_in"] = np.ascontiguousarray(actions, np.float32)
        b["terrain_origins"], b["height_samples"] = self.origins, self.hs
        d = draws or {}
        b["reset_override"] = None if d.get("reset") is None else np.ascontiguousarray(d["reset"], np.float32)
        b["noise_override"] = None if d.get("noise") is None else np.ascontiguousarray(d["noise"], np.float32)
        b["push_override"] = None if d.get("push") is None else np.ascontiguousarray(d["push"], np.float32)
        hf = self.heightfield
        mode = 3 if post_only == 2 else (2 if post_only else 1)
        self.emu.terrain(self.model, self.params, self.props, self.cfg, mode, b, self.common_step, self.init_done,
                         hf[0] if hf else None, hf[1] if hf else None)

    def anymal_create(self, cfg):
        n, nd = self.n, self.art.num_dofs
        no = 12 + 3 * nd
        self.cfg = cfg
        self.task = dict(obs=np.zeros((n, no), np.float32), obs_clamped=np.zeros((n, no), np.float32), rew=np.zeros(n, np.float32),
                         reset=np.ones(n, np.int64), progress=np.zeros(n, np.int64), timeout=np.zeros(n, np.int64),
                         commands=np.zeros((n, 3), np.float32), actions=np.zeros((n, nd), np.float32), reset_count=np.zeros(n, np.int32))

    def _bufs(self):
        b = dict(self.task)
        b.update(root=self.root, dof=self.dof, dof_force=self.dof_force, contact=self.contact)
        return b

    def anymal_reset_all(self, draws=None):
        self.emu.anymal(self.model, self.params, self.props, self.cfg, 0, self._bufs(), None, draws)

    def anymal_step(self, actions, draws=None):
        self.emu.anymal(self.model, self.params, self.props, self.cfg, 1, self._bufs(), actions, draws)

    def anymal_post_only(self, actions, draws=None):
        self.emu.anymal(self.model, self.params, self.props, self.cfg, 2, self._bufs(), actions, draws)

    def get_task(self):
        out = {k: self.task[k].copy() for k in TASK_KEYS}
        if getattr(self, "kind", "") == "terrain":
            out.update({k: self.task[k].copy() for k in TERRAIN_KEYS})
            out["report"] = self.task["report"].copy()
        out["dof_force"], out["contact"] = self.dof_force.copy(), self.contact.copy()
        return out

    def set_step(self, common_step, init_done=1):
        self.common_step, self.init_done = int(common_step), int(init_done)

    def set_task(self, **kw):
        for k, v in kw.items():
            if k == "dof_force":
                self.dof_force[:] = v
            elif k == "contact":
                self.contact[:] = v
            else:
                self.task[k][:] = v

    def close(self):
        pass


class CudaBackend:
    name = "cuda"

    def __init__(self, art, params, props, n):
        import torch

        from isaacgymenv_b200 import _lib

        self.torch, self._lib = torch, _lib
        self.lib = _lib.load()
        self.art, self.n = art, n
        self.model = _abi.pack_model(art)
        self.sim = C.c_void_p()
        p = _abi.SimParams.from_buffer_copy(params)
        ground = (p.has_ground, p.plane_static_friction, p.plane_dynamic_friction, p.plane_restitution)
        _lib.check(self.lib.b2g_sim_create(0, C.byref(p), C.byref(self.sim)), "create")
        if ground[0]:
            _lib.check(self.lib.b2g_sim_add_ground(self.sim, ground[1], ground[2], ground[3]))
        pose = (C.c_float * 7)(0, 0, 0, 0, 0, 0, 1)
        _lib.check(self.lib.b2g_sim_add_articulation(self.sim, C.byref(self.model), C.byref(props), n, pose, 1.0, 1), "add")
        _lib.check(self.lib.b2g_sim_prepare(self.sim), "prepare")
        self.t = {k: self._tensor(k) for k in (_abi.T_ROOT_STATE, _abi.T_DOF_STATE, _abi.T_NET_CONTACT, _abi.T_DOF_FORCE,
                                                _abi.T_DOF_TARGET, _abi.T_DOF_ACTUATION)}
        self.task = None
        self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _tensor(self, kind, task=False):
        d = _abi.TensorDesc()
        fn = self.lib.b2g_task_tensor if task else self.lib.b2g_sim_tensor
        self._lib.check(fn(self.sim, kind, C.byref(d)), "tensor")
        return self._lib.desc_to_torch(d)

    def _put(self, t, a):
        t.copy_(self.torch.from_numpy(np.ascontiguousarray(a)).to(t.device).reshape(t.shape))

    def set_state(self, root, dof):
        self._put(self.t[_abi.T_ROOT_STATE], np.asarray(root, np.float32))
        self._put(self.t[_abi.T_DOF_STATE], np.asarray(dof, np.float32))

    def get_state(self):
        nd = self.art.num_dofs
        return self.t[_abi.T_ROOT_STATE].cpu().numpy().copy(), self.t[_abi.T_DOF_STATE].cpu().numpy().reshape(self.n, nd, 2).copy()

    def simulate(self, target, actuation):
        self._put(self.t[_abi.T_DOF_TARGET], np.asarray(target, np.float32))
        self._put(self.t[_abi.T_DOF_ACTUATION], np.asarray(actuation, np.float32))
        self._lib.check(self.lib.b2g_sim_simulate(self.sim, self.stream), "simulate")
        self.torch.cuda.synchronize()
        nd, nb = self.art.num_dofs, self.art.num_bodies
        return (self.t[_abi.T_DOF_FORCE].cpu().numpy().reshape(self.n, nd).copy(),
                self.t[_abi.T_NET_CONTACT].cpu().numpy().reshape(self.n, nb, 3).copy())

    def add_heightfield(self, hf_t, samples):
        smp = np.ascontiguousarray(samples, np.int16)
        self._lib.check(self.lib.b2g_sim_add_heightfield(self.sim, C.byref(hf_t), smp.ctypes.data_as(C.c_void_p)), "add_heightfield")

    def set_friction(self, friction):
        self._put(self._tensor(_abi.T_FRICTION), np.asarray(friction, np.float32))

    def contact_stats(self, reset=False):
        out = (C.c_int64 * 4)()
        self._lib.check(self.lib.b2g_sim_contact_stats(self.sim, out, 1 if reset else 0), "contact_stats")
        return [int(v) for v in out]

    def set_link_scale(self, scale):
        """(N, nd + 1, 6) per-link [mass, stiffness, damping scale, lower, upper limit offset, spare] rows (B2G_T_LINK_SCALE)."""
        self._put(self._tensor(_abi.T_LINK_SCALE), np.asarray(scale, np.float32))

    def set_env_scale(self, scale, friction=None):
        """(N,4) per-env [mass, stiffness, damping, spare] scales (+ optional (N) shape friction)."""
        self._put(self._tensor(_abi.T_ENV_SCALE), np.asarray(scale, np.float32))
        if friction is not None:
            self._put(self._tensor(_abi.T_FRICTION), np.asarray(friction, np.float32))

    def forward_dynamics(self, tau):
        torch = self.torch
        self._put(self.t[_abi.T_DOF_ACTUATION], np.asarray(tau, np.float32))
        qdd = torch.zeros(self.n, self.art.num_dofs, device="cuda:0")
        a0 = torch.zeros(self.n, 6, device="cuda:0")
        self._lib.check(self.lib.b2g_sim_forward_dynamics(self.sim, C.c_void_p(qdd.data_ptr()), C.c_void_p(a0.data_ptr()), self.stream), "fd")
        torch.cuda.synchronize()
        return qdd.cpu().numpy(), a0.cpu().numpy()

    def jacobian_mass_matrix(self):
        j, m = self._tensor(_abi.T_JACOBIAN), self._tensor(_abi.T_MASS_MATRIX)
        self._lib.check(self.lib.b2g_sim_refresh(self.sim, _abi.T_JACOBIAN, self.stream), "refresh jacobian")
        self._lib.check(self.lib.b2g_sim_refresh(self.sim, _abi.T_MASS_MATRIX, self.stream), "refresh mass matrix")
        self.torch.cuda.synchronize()
        return j.cpu().numpy().copy(), m.cpu().numpy().copy()

    def anymal_create(self, cfg):
        self._lib.check(self.lib.b2g_task_anymal_create(self.sim, C.byref(cfg)), "task create")
        kinds = dict(obs=_abi.TT_OBS, obs_clamped=_abi.TT_OBS_CLAMPED, rew=_abi.TT_REW, reset=_abi.TT_RESET, progress=_abi.TT_PROGRESS,
                     timeout=_abi.TT_TIMEOUT, commands=_abi.TT_COMMANDS, actions=_abi.TT_ACTIONS, rand=_abi.TT_RAND_OVERRIDE)
        self.task = {k: self._tensor(v, task=True) for k, v in kinds.items()}

    def cartpole_create(self, cfg):
        self._lib.check(self.lib.b2g_task_cartpole_create(self.sim, C.byref(cfg)), "task create")
        kinds = dict(obs=_abi.TT_OBS, obs_clamped=_abi.TT_OBS_CLAMPED, rew=_abi.TT_REW, reset=_abi.TT_RESET, progress=_abi.TT_PROGRESS,
                     timeout=_abi.TT_TIMEOUT, commands=_abi.TT_COMMANDS, actions=_abi.TT_ACTIONS, rand=_abi.TT_RAND_OVERRIDE)
        self.task = {k: self._tensor(v, task=True) for k, v in kinds.items()}

    def houndarm_create(self, cfg):
        self._lib.check(self.lib.b2g_task_houndarm_create(self.sim, C.byref(cfg)), "task create")
        kinds = dict(obs=_abi.TT_OBS, obs_clamped=_abi.TT_OBS_CLAMPED, rew=_abi.TT_REW, reset=_abi.TT_RESET, progress=_abi.TT_PROGRESS,
                     timeout=_abi.TT_TIMEOUT, commands=_abi.TT_COMMANDS, actions=_abi.TT_ACTIONS, rand=_abi.TT_RAND_OVERRIDE)
        self.task = {k: self._tensor(v, task=True) for k, v in kinds.items()}

    def terrain_create(self, cfg, height_samples=None, terrain_origins=None, heightfield=None):
        hs = None if height_samples is None else np.ascontiguousarray(height_samples, np.int16)
        og = None if terrain_origins is None else np.ascontiguousarray(terrain_origins, np.float32)
        if heightfield is not None:
            self._lib.check(self.lib.b2g_sim_add_heightfield(self.sim, C.byref(heightfield[0]),
                                                             np.ascontiguousarray(heightfield[1], np.int16).ctypes.data_as(C.c_void_p)), "add_heightfield")
        self._lib.check(self.lib.b2g_task_terrain_create(self.sim, C.byref(cfg), None if hs is None else hs.ctypes.data_as(C.c_void_p),
                                                         None if og is None else og.ctypes.data_as(C.c_void_p)), "terrain create")
        self.kind = "terrain"
        kinds = dict(obs=_abi.TT_OBS, obs_clamped=_abi.TT_OBS_CLAMPED, rew=_abi.TT_REW, reset=_abi.TT_RESET, progress=_abi.TT_PROGRESS,
                     timeout=_abi.TT_TIMEOUT, commands=_abi.TT_COMMANDS, actions=_abi.TT_ACTIONS, rand=_abi.TT_RAND_OVERRIDE, torques=_abi.TT_TORQUES,
                     last_actions=_abi.TT_LAST_ACTIONS, last_dof_vel=_abi.TT_LAST_DOF_VEL, feet_air_time=_abi.TT_FEET_AIR_TIME,
                     episode_sums=_abi.TT_EPISODE_SUMS, env_origins=_abi.TT_ENV_ORIGINS, terrain_levels=_abi.TT_TERRAIN_LEVELS,
                     terrain_types=_abi.TT_TERRAIN_TYPES, noise=_abi.TT_NOISE_OVERRIDE, push=_abi.TT_PUSH_OVERRIDE, extras=_abi.TT_EXTRAS,
                     measured=_abi.TT_MEASURED_HEIGHTS, arm_mm=_abi.TT_ARM_MM, arm_jac=_abi.TT_ARM_JAC, eef_state=_abi.TT_EEF_STATE,
                     arm_commands=_abi.TT_ARM_COMMANDS)
        self.task = {k: self._tensor(v, task=True) for k, v in kinds.items()}
        self.set_step(0, 1)

    def set_step(self, common_step, init_done=1):
        self._lib.check(self.lib.b2g_task_terrain_set_step(self.sim, int(common_step)))
        self._lib.check(self.lib.b2g_task_terrain_set_init_done(self.sim, int(init_done)))

    def task_step(self, actions, draws=None, post_only=False):
        if getattr(self, "kind", "") == "terrain" and draws is not None:
            self._put(self.task["rand"], np.asarray(draws["reset"], np.float32))
            self._put(self.task["noise"], np.asarray(draws["noise"], np.float32))
            self._put(self.task["push"], np.asarray(draws["push"], np.float32))
            self._lib.check(self.lib.b2g_task_set_rand_override(self.sim, 1))
        else:
            self._draws(draws)
        fn = self.lib.b2g_task_osc_probe if post_only == 2 else (self.lib.b2g_task_post_only if post_only else self.lib.b2g_task_step)
        self._lib.check(fn(self.sim, self._actions(actions), self.stream), "task step")
        self.torch.cuda.synchronize()

    def _draws(self, draws):
        if draws is None:
            self._lib.check(self.lib.b2g_task_set_rand_override(self.sim, 0))
        else:
            self._put(self.task["rand"], np.asarray(draws, np.float32))
            self._lib.check(self.lib.b2g_task_set_rand_override(self.sim, 1))

    def anymal_reset_all(self, draws=None):
        self._draws(draws)
        self._lib.check(self.lib.b2g_task_anymal_reset_all(self.sim, self.stream), "reset_all")
        self.torch.cuda.synchronize()

    def _actions(self, actions):
        self._a = self.torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32)).to("cuda:0")
        return C.c_void_p(self._a.data_ptr())

    def anymal_step(self, actions, draws=None):
        self._draws(draws)
        self._lib.check(self.lib.b2g_task_anymal_step(self.sim, self._actions(actions), self.stream), "step")
        self.torch.cuda.synchronize()

    def anymal_post_only(self, actions, draws=None):
        self._draws(draws)
        self._lib.check(self.lib.b2g_task_anymal_post_only(self.sim, self._actions(actions), self.stream), "post_only")
        self.torch.cuda.synchronize()

    def get_task(self):
        nd, nb = self.art.num_dofs, self.art.num_bodies
        out = {k: self.task[k].cpu().numpy().copy() for k in TASK_KEYS}
        if getattr(self, "kind", "") == "terrain":
            out.update({k: self.task[k].cpu().numpy().copy() for k in TERRAIN_KEYS})
            out["extras"] = self.task["extras"].cpu().numpy().copy()
        out["dof_force"] = self.t[_abi.T_DOF_FORCE].cpu().numpy().reshape(self.n, nd).copy()
        out["contact"] = self.t[_abi.T_NET_CONTACT].cpu().numpy().reshape(self.n, nb, 3).copy()
        return out

    def set_task(self, **kw):
        for k, v in kw.items():
            if k == "dof_force":
                self._put(self.t[_abi.T_DOF_FORCE], np.asarray(v, np.float32))
            elif k == "contact":
                self._put(self.t[_abi.T_NET_CONTACT], np.asarray(v, np.float32))
            else:
                self._put(self.task[k], np.asarray(v, self.task[k].cpu().numpy().dtype))

    def close(self):
        if self.sim:
            self.t, self.task = {}, None
            self.lib.b2g_sim_destroy(self.sim)
            self.sim = C.c_void_p()
