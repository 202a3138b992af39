"""TEST INFRASTRUCTURE: ctypes front-end of the host lane emulator (tests/emu/emu.cpp), which runs the
CUDA kernels' per-thread code on the CPU so the kernel logic can be checked against the oracle
without a GPU.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
_LIB = os.path.join(_HERE, "libb2g_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "emu.cpp")] + [os.path.join(_ROOT, "isaacgymenv_b200", "csrc", f) for f in
            ("b2g_threads.cuh", "b2g_dynamics.cuh", "b2g_math.cuh", "b2g_dev.h", "b2g_host_pack.h")] + [os.path.join(_ROOT, "include", "b200gym.h")]
    stale = not os.path.isfile(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(_ROOT, "include"),
                               "-I" + os.path.join(_ROOT, "isaacgymenv_b200", "csrc"), "-x", "c++", os.path.join(_HERE, "emu.cpp"),
                               "-o", _LIB, "-lpthread"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a, ct=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


_env_scale_keep = None


def set_env_scale(scale):
    """(N,4) per-env [mass, stiffness, damping, spare] scales for the following simulate / forward_dynamics calls; None = ones."""
    global _env_scale_keep
    _env_scale_keep = None if scale is None else np.ascontiguousarray(scale, dtype=np.float32)
    lib().emu_set_env_scale.restype = None
    lib().emu_set_env_scale(_p(_env_scale_keep))


def hfc_stats(reset=True):
    """(links tested against the coarse heightfield bound, links skipped) since the last reset."""
    out = (C.c_longlong * 2)()
    lib().emu_hfc_stats.restype = None
    lib().emu_hfc_stats(out, C.c_int(1 if reset else 0))
    return int(out[0]), int(out[1])


def simulate(model, params, props, root, dof, target, actuation, heightfield=None, hf_samples=None, friction=None):
    n = root.shape[0]
    assert root.dtype == np.float32 and dof.dtype == np.float32
    dof_force = np.zeros((n, model.n_dof), dtype=np.float32)
    contact = np.zeros((n, model.n_bodies, 3), dtype=np.float32)
    hs = None if hf_samples is None else np.ascontiguousarray(hf_samples, dtype=np.int16)
    fr = None if friction is None else np.ascontiguousarray(friction, dtype=np.float32)
    rc = lib().emu_simulate(C.byref(model), C.byref(params), C.byref(props), C.byref(heightfield) if heightfield is not None else None,
                            _p(hs, C.c_int16), _p(fr), C.c_int(n), _p(root), _p(dof),
                            _p(np.ascontiguousarray(target, dtype=np.float32)), _p(np.ascontiguousarray(actuation, dtype=np.float32)),
                            _p(dof_force), _p(contact))
    assert rc == 0, rc
    return dof_force, contact


def forward_dynamics(model, params, props, root, dof, tau):
    n = root.shape[0]
    qdd = np.zeros((n, model.n_dof), dtype=np.float32)
    a0 = np.zeros((n, 6), dtype=np.float32)
    rc = lib().emu_forward_dynamics(C.byref(model), C.byref(params), C.byref(props), C.c_int(n), _p(root), _p(dof),
                                    _p(np.ascontiguousarray(tau, dtype=np.float32)), _p(qdd), _p(a0))
    assert rc == 0, rc
    return qdd, a0


def anymal(model, params, props, cfg, mode, bufs, actions=None, rand_override=None):
    """mode 0 = reset_all, 1 = step.  bufs: dict of numpy arrays (root, dof, dof_force, contact, obs, obs_clamped,
    rew, reset, progress, timeout, commands, actions, reset_count) updated in place."""
    n = bufs["root"].shape[0]
    ai = None if actions is None else np.ascontiguousarray(actions, dtype=np.float32)
    ro = None if rand_override is None else np.ascontiguousarray(rand_override, dtype=np.float32)
    rc = lib().emu_anymal(C.byref(model), C.byref(params), C.byref(props), C.byref(cfg), C.c_int(mode), C.c_int(n),
                          _p(bufs["root"]), _p(bufs["dof"]), _p(bufs["dof_force"]), _p(bufs["contact"]), _p(ai), _p(bufs["obs"]),
                          _p(bufs["obs_clamped"]), _p(bufs["rew"]), _p(bufs["reset"], C.c_longlong), _p(bufs["progress"], C.c_longlong),
                          _p(bufs["timeout"], C.c_longlong), _p(bufs["commands"]), _p(bufs["actions"]), _p(bufs["reset_count"], C.c_int), _p(ro))
    assert rc == 0, rc


def cartpole(model, params, props, cfg, mode, bufs, actions, rand_override=None):
    """mode 1 = step, 2 = post_physics_step only."""
    n = bufs["root"].shape[0]
    ai = np.ascontiguousarray(actions, dtype=np.float32)
    ro = None if rand_override is None else np.ascontiguousarray(rand_override, dtype=np.float32)
    rc = lib().emu_cartpole(C.byref(model), C.byref(params), C.byref(props), C.byref(cfg), C.c_int(mode), C.c_int(n),
                            _p(bufs["root"]), _p(bufs["dof"]), _p(bufs["dof_force"]), _p(bufs["contact"]), _p(ai), _p(bufs["obs"]),
                            _p(bufs["obs_clamped"]), _p(bufs["rew"]), _p(bufs["reset"], C.c_longlong), _p(bufs["progress"], C.c_longlong),
                            _p(bufs["timeout"], C.c_longlong), _p(bufs["actions"]), _p(bufs["reset_count"], C.c_int), _p(ro))
    assert rc == 0, rc


def houndarm(model, params, props, cfg, mode, bufs, actions, rand_override=None):
    """mode 1 = step, 2 = post_physics_step only."""
    n = bufs["root"].shape[0]
    ai = np.ascontiguousarray(actions, dtype=np.float32)
    ro = None if rand_override is None else np.ascontiguousarray(rand_override, dtype=np.float32)
    rc = lib().emu_houndarm(C.byref(model), C.byref(params), C.byref(props), C.byref(cfg), C.c_int(mode), C.c_int(n),
                            _p(bufs["root"]), _p(bufs["dof"]), _p(bufs["dof_force"]), _p(bufs["contact"]), _p(ai), _p(bufs["obs"]),
                            _p(bufs["obs_clamped"]), _p(bufs["rew"]), _p(bufs["reset"], C.c_longlong), _p(bufs["progress"], C.c_longlong),
                            _p(bufs["timeout"], C.c_longlong), _p(bufs["commands"]), _p(bufs["actions"]), _p(bufs["reset_count"], C.c_int), _p(ro))
    assert rc == 0, rc


class _TerrainBufs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("root", "dof", "dof_force", "contact", "actions_in", "obs", "obs_clamped", "rew", "reset", "progress",
                                          "timeout", "commands", "actions", "torques", "last_actions", "last_dof_vel", "feet_air_time",
                                          "episode_sums", "env_origins", "terrain_levels", "terrain_types", "terrain_origins", "height_samples",
                                          "scratch", "resetw", "report", "measured", "arm_mm", "arm_jac", "eef_state", "arm_commands", "reset_count", "reset_override", "noise_override",
                                          "push_override", "friction")]


def terrain(model, params, props, cfg, mode, bufs, common_step, init_done, heightfield=None, hf_samples=None):
    """mode 1 = step, 2 = post_physics_step only, 3 = OSC probe.  bufs: dict name -> numpy array (or None) for every _TerrainBufs field."""
    tb = _TerrainBufs()
    keep = []
    for k, _ in _TerrainBufs._fields_:
        a = bufs.get(k)
        if a is None:
            setattr(tb, k, None)
        else:
            assert a.flags["C_CONTIGUOUS"], k
            keep.append(a)
            setattr(tb, k, a.ctypes.data)
    n = bufs["root"].shape[0]
    hs = None if hf_samples is None else np.ascontiguousarray(hf_samples, dtype=np.int16)
    rc = lib().emu_terrain(C.byref(model), C.byref(params), C.byref(props), C.byref(heightfield) if heightfield is not None else None,
                           _p(hs, C.c_int16), C.byref(cfg), C.c_int(mode), C.c_int(n), C.c_longlong(common_step), C.c_int(init_done), C.byref(tb))
    assert rc == 0, rc


def jacobian_mass_matrix(model, props, root, dof):
    n, nd, nb = root.shape[0], model.n_dof, model.n_bodies
    rows = nb - (1 if model.fixed_base else 0)
    ncol = nd + (0 if model.fixed_base else 6)
    jac = np.zeros((n, rows, 6, ncol), np.float32)
    mm = np.zeros((n, nd, nd), np.float32)
    rc = lib().emu_jac_mm(C.byref(model), C.byref(props), C.c_int(n), _p(np.ascontiguousarray(root, np.float32)), _p(np.ascontiguousarray(dof, np.float32)),
                          _p(jac), C.c_int(rows * 6 * ncol), _p(mm))
    assert rc == 0, rc
    return jac, mm


_link_scale_keep = None


def set_link_scale(scale):
    """(N, nd + 1, 6) per-link [mass, stiffness, damping scale, lower, upper limit offset, spare] rows used by the next emulated calls; None = ones."""
    global _link_scale_keep
    _link_scale_keep = None if scale is None else np.ascontiguousarray(scale, dtype=np.float32)
    lib().emu_set_link_scale.restype = None
    lib().emu_set_link_scale(_p(_link_scale_keep))


def contact_stats(reset=False):
    """[active contact points, dropped candidates, env sub-steps with a drop, env sub-steps] of the emulated sub-steps."""
    out = (C.c_longlong * 4)()
    lib().emu_contact_stats(out, C.c_int(1 if reset else 0))
    return [int(v) for v in out]
