"""ORACLE -- TEST INFRASTRUCTURE ONLY.  The CPU analogue of ``VecTask.step`` for the flat Anymal task, built from the
oracle pieces (C dynamics restatement + numpy task math in the reference's order).  Used by ``bench.py`` as the timed
CPU baseline (``cpu_baseline`` and ``--impl reference``: kind "port" -- Isaac Gym / PhysX, the reference's real CPU
pipeline, is a closed binary that is not installed, so the restated algorithm is what can be timed) and by
``__graft_entry__.smoke()`` / tests as the checker.  Never imported by the product path.

Threads: environments are independent, so the batch is cut into contiguous slices, one per host thread; the C
oracle runs without the GIL (ctypes), numpy task math runs on the whole batch.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import ctypes as C
import hashlib
import subprocess

from . import dyn_oracle as O
from . import task_math as tm

_HERE = os.path.dirname(os.path.abspath(__file__))
_NATIVE = os.path.join(_HERE, "_build", "liboracle_step_native.so")


def _host_signature(srcs):
    """Identifies (source state, CPU model/flags): -march=native code must never run on another CPU type."""
    h = hashlib.sha1()
    for s in srcs:
        with open(s, "rb") as fh:
            h.update(fh.read())
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith(("flags", "model name")):
                    h.update(line.encode())
                    if line.startswith("flags"):
                        break
    except OSError:
        pass
    return h.hexdigest()


def build_native(force: bool = False):
    """gcc -O3 -march=native -fopenmp of the whole-step CPU baseline (oracle/dyn/oracle_step_omp.c), built on the machine
    that runs it; falls back to a portable -O3 build when -march=native is refused.  Returns (path, flags)."""
    srcs = [os.path.join(_HERE, "dyn", f) for f in ("oracle_step_omp.c", "oracle_dyn_impl.h")] + [os.path.join(_HERE, "..", "include", "b200gym.h")]
    sig, sig_path = _host_signature(srcs), _NATIVE + ".sig"
    if not force and os.path.isfile(_NATIVE) and os.path.isfile(sig_path):
        with open(sig_path) as fh:
            got = fh.read().split("\n")
        if got and got[0] == sig:
            return _NATIVE, got[1] if len(got) > 1 else "?"
    os.makedirs(os.path.dirname(_NATIVE), exist_ok=True)
    base = ["gcc", "-fPIC", "-shared", "-fopenmp", "-Wno-misleading-indentation", "-Wno-unused-function", "-I" + os.path.join(_HERE, "..", "include"),
            "-o", _NATIVE, srcs[0], "-lm"]
    for flags in (["-O3", "-march=native"], ["-O3"]):
        try:
            subprocess.check_call(base[:1] + flags + base[1:], stderr=subprocess.DEVNULL)
            with open(sig_path, "w") as fh:
                fh.write(sig + "\n" + " ".join(flags + ["-fopenmp"]))
            return _NATIVE, " ".join(flags + ["-fopenmp"])
        except (subprocess.CalledProcessError, OSError):
            continue
    raise RuntimeError("could not build the native CPU baseline (gcc missing?)")


class CpuAnymalStepNative:
    """Whole flat-task step in C with one OpenMP loop over the environments (oracle/dyn/oracle_step_omp.c)."""

    def __init__(self, model, params, props, cfg_struct, n_envs, threads=None, seed=42):
        path, self.flags = build_native()
        self.lib = C.CDLL(path)
        self.lib.orc_anymal_step_omp.restype = C.c_int
        self.lib.orc_omp_threads.restype = C.c_int
        self.model, self.params, self.props, self.cfg = model, params, props, cfg_struct
        self.n, self.nd, self.nb = n_envs, model.n_dof, model.n_bodies
        self.threads = int(threads or os.cpu_count() or 1)
        self.rng = np.random.default_rng(seed)
        n, nd, nb = self.n, self.nd, self.nb
        f32 = np.float32
        self.root = np.tile(np.array(list(cfg_struct.init_root), f32), (n, 1))
        self.dof = np.zeros((n, nd, 2), f32)
        self.commands = np.zeros((n, 3), f32)
        self.progress = np.zeros(n, np.int64)
        self.reset = np.ones(n, np.int64)
        self.torques = np.zeros((n, nd), f32)
        self.contact = np.zeros((n, nb, 3), f32)
        self.obs = np.zeros((n, 12 + 3 * nd), f32)
        self.obs_clamped = np.zeros_like(self.obs)
        self.rew = np.zeros(n, f32)
        self.timeout = np.zeros(n, np.int64)

    def step(self, actions, draws=None):
        if draws is None:
            draws = self.rng.random((self.n, 2 * self.nd + 3), dtype=np.float32)
        a = np.ascontiguousarray(actions, np.float32)
        d = np.ascontiguousarray(draws, np.float32)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int64)
        p = lambda x, t=fp: x.ctypes.data_as(t)
        rc = self.lib.orc_anymal_step_omp(C.byref(self.model), C.byref(self.params), C.byref(self.props), C.byref(self.cfg), C.c_int(self.n),
                                          C.c_int(self.threads), p(self.root), p(self.dof), p(a), p(d), p(self.commands), p(self.progress, ip),
                                          p(self.reset, ip), p(self.torques), p(self.contact), p(self.obs), p(self.obs_clamped), p(self.rew),
                                          p(self.timeout, ip))
        if rc != 0:
            raise RuntimeError("native CPU step failed")
        return self.obs, self.obs_clamped, self.rew, self.timeout


class CpuAnymalStep:
    def __init__(self, model, params, props, cfg, n_envs, threads=None, dtype=np.float32, seed=42):
        self.model, self.params, self.props, self.cfg = model, params, props, cfg
        self.n, self.nd, self.nb = n_envs, model.n_dof, model.n_bodies
        self.dtype = np.dtype(dtype)
        self.threads = int(threads or os.cpu_count() or 1)
        self.pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None
        self.rng = np.random.default_rng(seed)
        n, nd, nb = self.n, self.nd, self.nb
        self.state = dict(root=np.tile(cfg["init_root"].astype(self.dtype), (n, 1)), dof_pos=np.zeros((n, nd), self.dtype),
                          dof_vel=np.zeros((n, nd), self.dtype), torques=np.zeros((n, nd), self.dtype),
                          contact=np.zeros((n, nb, 3), self.dtype), commands=np.zeros((n, 3), self.dtype),
                          progress=np.zeros(n, np.int64), reset=np.ones(n, np.int64))
        self.dof = np.zeros((n, nd, 2), self.dtype)
        self.zero = np.zeros((n, nd), self.dtype)
        bounds = np.linspace(0, n, self.threads + 1).astype(int)
        self.slices = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]

    def _sim_slice(self, ab, tgt):
        a, b = ab
        f, c = O.simulate(self.model, self.params, self.props, self.state["root"][a:b], self.dof[a:b], tgt[a:b], self.zero[a:b])
        self.state["torques"][a:b] = f
        self.state["contact"][a:b] = c

    def step(self, actions):
        cfg, st = self.cfg, self.state
        a = np.clip(actions.astype(self.dtype), -1.0, 1.0)
        tgt = (self.dtype.type(0.5) * a + cfg["default_dof_pos"].astype(self.dtype)[None]).astype(self.dtype)
        self.dof[:, :, 0], self.dof[:, :, 1] = st["dof_pos"], st["dof_vel"]
        if self.pool:
            list(self.pool.map(lambda ab: self._sim_slice(ab, tgt), self.slices))
        else:
            self._sim_slice((0, self.n), tgt)
        st["dof_pos"][:], st["dof_vel"][:] = self.dof[:, :, 0], self.dof[:, :, 1]
        draws = self.rng.random((self.n, 2 * self.nd + 3), dtype=np.float32)
        return tm.anymal_post_physics(st, cfg, a, draws)
