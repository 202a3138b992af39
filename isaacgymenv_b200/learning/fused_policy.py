"""Rollout-side policy evaluation through ``libb200gym``'s fused tcgen05 kernel (``b2g_policy_*`` in ``include/b200gym.h``).

The learner keeps its fp32 torch parameters (``learning/ppo.py``); after every optimiser update :meth:`FusedPolicy.sync`
re-packs them (bf16, tensor-core operand layout) and the rollout calls :meth:`FusedPolicy.forward` -- one kernel launch for
normalise -> 3 x (linear + ELU) -> mu / value heads -- instead of ~12 torch kernels.  There is no torch fallback: without the
CUDA library the constructor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

HIDDEN0, HIDDEN1, HIDDEN2, MU, VALUE = range(5)


class FusedPolicy:
    def __init__(self, num_obs: int, num_actions: int, units=(256, 128, 64), device="cuda:0"):
        if len(units) != 3:
            raise _lib.B2GError("FusedPolicy supports exactly three hidden layers")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B2GError("FusedPolicy needs a CUDA device (no CPU path)")
        self.num_obs, self.num_actions = int(num_obs), int(num_actions)
        self._h = C.c_void_p()
        arr = (C.c_int * 3)(*[int(u) for u in units])
        _lib.check(self.lib.b2g_policy_create(self.device.index or 0, self.num_obs, arr, self.num_actions, C.byref(self._h)),
                   "b2g_policy_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.b2g_policy_destroy(h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _f32(t):
        return t.detach().to(torch.float32).contiguous()

    def set_layer(self, layer: int, weight: torch.Tensor, bias: torch.Tensor):
        w, b = self._f32(weight), self._f32(bias)
        _lib.check(self.lib.b2g_policy_set_layer(self._h, layer, C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), self._stream()),
                   "b2g_policy_set_layer")

    def set_obs_norm(self, mean=None, var=None, eps=1e-5, clip=5.0):
        m = self._f32(mean) if mean is not None else None
        v = self._f32(var) if var is not None else None
        _lib.check(self.lib.b2g_policy_set_obs_norm(self._h, C.c_void_p(m.data_ptr() if m is not None else None),
                                                    C.c_void_p(v.data_ptr() if v is not None else None),
                                                    C.c_float(eps), C.c_float(clip), self._stream()), "b2g_policy_set_obs_norm")

    def sync(self, model, obs_rms=None, critic=False):
        """Pack the parameters of a ``learning.ppo.ActorCritic`` (and its observation normaliser) for the kernel.  ``critic=True``
        packs the critic tower of a ``separate`` network instead of the actor tower / shared trunk (the value head is the same
        module either way; the caller then ignores the action head of this instance, and the value head of the actor instance)."""
        trunk = model.critic_trunk if critic else model.trunk
        linears = [m for m in trunk if isinstance(m, torch.nn.Linear)]
        for i, lin in enumerate(linears):
            self.set_layer(HIDDEN0 + i, lin.weight, lin.bias)
        self.set_layer(MU, model.mu.weight, model.mu.bias)
        self.set_layer(VALUE, model.value.weight, model.value.bias)
        if obs_rms is not None:
            self.set_obs_norm(obs_rms.mean, obs_rms.var, obs_rms.eps, 5.0)

    def forward(self, obs: torch.Tensor, mu: torch.Tensor = None, value: torch.Tensor = None):
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape[1] == self.num_obs
        n = obs.shape[0]
        if mu is None:
            mu = torch.empty(n, self.num_actions, device=obs.device)
        if value is None:
            value = torch.empty(n, device=obs.device)
        _lib.check(self.lib.b2g_policy_forward(self._h, C.c_void_p(obs.data_ptr()), n, C.c_void_p(mu.data_ptr()),
                                               C.c_void_p(value.data_ptr()), self._stream()), "b2g_policy_forward")
        return mu, value

    @property
    def launches(self) -> int:
        return int(self.lib.b2g_policy_launch_count(self._h))
