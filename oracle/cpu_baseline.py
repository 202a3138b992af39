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

from . import dyn_oracle as O
from . import task_math as tm


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
