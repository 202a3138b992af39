"""Learner-side kernels of ``libb200gym`` behind torch (``b2g_ppo_head`` / ``b2g_adam_clip_step`` in ``include/b200gym.h``):

* :class:`FlatParameters` -- the network's parameters (and their gradients) as views of ONE flat buffer each, so that the optimiser
  and the multi-GPU all-reduce work on a single vector (no concatenation, no scatter back);
* :func:`ppo_head_loss` -- the PPO loss head as one kernel (+ a one-block finalize) with closed-form gradients, exposed to autograd:
  ``loss.backward()`` continues through the torch network from ``d loss / d mu`` and ``d loss / d value``;
* :class:`FusedClipAdam` -- global-norm clipping + Adam in two launches, learning rate and step count in device memory;
* :func:`ppo_head_reference` -- the same loss in plain torch ops (what ``learning/ppo.py`` runs without the library): the parity
  reference of the kernel, and of the closed-form gradients through autograd.

Loss structure: rl_games' ``a2c_continuous`` (the fork's in-tree statement: ``learning/common_agent.py:312-400``)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

MAX_ACTIONS = 24
ADAM_MAX_PARTIALS = 512
_fp = C.POINTER(C.c_float)


class PpoHeadArgs(C.Structure):
    _fields_ = [("mu", C.c_void_p), ("value", C.c_void_p), ("log_std", C.c_void_p), ("index", C.c_void_p), ("actions", C.c_void_p),
                ("old_mu", C.c_void_p), ("old_neglogp", C.c_void_p), ("advantages", C.c_void_p), ("old_values", C.c_void_p), ("returns", C.c_void_p),
                ("n_rows", C.c_int32), ("n_actions", C.c_int32), ("e_clip", C.c_float), ("critic_coef", C.c_float), ("entropy_coef", C.c_float),
                ("bounds_loss_coef", C.c_float), ("mu_bound", C.c_float), ("grad_mu", C.c_void_p), ("grad_value", C.c_void_p),
                ("grad_log_std", C.c_void_p), ("out", C.c_void_p), ("partial", C.c_void_p)]


class AdamArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("n", C.c_int32),
                ("lr", C.c_void_p), ("step", C.c_void_p), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("max_grad_norm", C.c_float), ("grad_scale", C.c_float), ("partial", C.c_void_p), ("out_norm", C.c_void_p)]


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ppo_head_reference(mu, value, log_std, actions, old_mu, old_neglogp, advantages, old_values, returns, e_clip, critic_coef, entropy_coef,
                       bounds_loss_coef, mu_bound=1.1):
    """Plain torch statement of the loss head; returns (loss, a_loss, c_loss, b_loss, kl, entropy)."""
    ls = log_std.expand_as(mu)
    nlp = 0.5 * (((actions - mu) / ls.exp()) ** 2).sum(-1) + ls.sum(-1) + 0.5 * mu.shape[-1] * 1.8378770664093453
    ratio = torch.exp(old_neglogp - nlp)
    a_loss = torch.max(-advantages * ratio, -advantages * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)).mean()
    v_clip = old_values + (value - old_values).clamp(-e_clip, e_clip)
    c_loss = torch.max((value - returns) ** 2, (v_clip - returns) ** 2).mean()
    b_loss = (torch.clamp(mu - mu_bound, min=0.0) ** 2 + torch.clamp(-mu_bound - mu, min=0.0) ** 2).sum(-1).mean()
    entropy = (ls + 0.5 + 0.9189385332046727).sum(-1).mean()
    loss = a_loss + 0.5 * critic_coef * c_loss - entropy_coef * entropy + bounds_loss_coef * b_loss
    with torch.no_grad():
        kl = (((mu - old_mu) ** 2) / (2.0 * torch.exp(2.0 * ls))).sum(-1).mean()
    return loss, a_loss, c_loss, b_loss, kl, entropy


class _PpoHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, value, log_std, head):
        ctx.head = head
        head.launch(mu, value, log_std)
        return head.out[0].clone()

    @staticmethod
    def backward(ctx, g):
        h = ctx.head
        return h.grad_mu * g, h.grad_value * g, h.grad_log_std * g, None


class _PpoHeadSeed(torch.autograd.Function):
    """Backward seed of :meth:`PpoHead.backward_direct`: a dummy scalar whose backward hands (mu, value) the gradients the head kernel
    has already computed, unscaled (the caller's contract: ``.backward()`` with the default unit gradient)."""

    @staticmethod
    def forward(ctx, mu, value, head, grads):
        ctx.grads = grads
        return mu.new_empty(())

    @staticmethod
    def backward(ctx, g):
        return ctx.grads[0], ctx.grads[1], None, None


class PpoHead:
    """Static buffers + argument block of ``b2g_ppo_head`` for one minibatch size (graph-capturable: no allocation per call)."""

    def __init__(self, n_rows, n_actions, device, e_clip, critic_coef, entropy_coef, bounds_loss_coef, mu_bound=1.1):
        if n_actions > MAX_ACTIONS:
            raise _lib.B2GError(f"b2g_ppo_head supports at most {MAX_ACTIONS} actions")
        self.lib = _lib.load()
        self.lib.b2g_ppo_head.argtypes = [C.POINTER(PpoHeadArgs), C.c_void_p]
        self.lib.b2g_ppo_head.restype = C.c_int
        self.lib.b2g_ppo_head_workspace_floats.restype = C.c_int
        self.device = torch.device(device)
        self.n_rows, self.n_actions = int(n_rows), int(n_actions)
        self.grad_mu = torch.zeros(n_rows, n_actions, device=device)
        self.grad_value = torch.zeros(n_rows, device=device)
        self.grad_log_std = torch.zeros(n_actions, device=device)
        self.out = torch.zeros(6, device=device)
        self.partial = torch.zeros(int(self.lib.b2g_ppo_head_workspace_floats(int(n_rows))), device=device)
        self.hyper = (float(e_clip), float(critic_coef), float(entropy_coef), float(bounds_loss_coef), float(mu_bound))
        self.buffers = None

    def bind(self, index, actions, old_mu, old_neglogp, advantages, old_values, returns):
        """The rollout buffers (full size, contiguous float32) and the minibatch index vector (int64) this head reads."""
        for t in (actions, old_mu, old_neglogp, advantages, old_values, returns):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        assert index is None or (index.dtype == torch.int64 and index.is_contiguous() and index.numel() == self.n_rows)
        self.buffers = (index, actions, old_mu, old_neglogp, advantages, old_values, returns)

    def launch(self, mu, value, log_std, grad_log_std=None, grad_mu=None, grad_value=None):
        """``grad_log_std``: where d loss / d log_std is stored (default: the head's own buffer) -- pass the parameter's ``.grad`` view
        to skip the autograd edge (see :meth:`backward_direct`)."""
        assert self.buffers is not None, "bind() the rollout buffers first"
        mu, value, log_std = mu.contiguous(), value.contiguous(), log_std.contiguous()
        assert mu.shape == (self.n_rows, self.n_actions) and value.shape == (self.n_rows,)
        index, actions, old_mu, old_nlp, adv, old_val, ret = self.buffers
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        e, cc, ec, bc, mb = self.hyper
        a = PpoHeadArgs(p(mu), p(value), p(log_std), p(index), p(actions), p(old_mu), p(old_nlp), p(adv), p(old_val), p(ret), self.n_rows, self.n_actions,
                        e, cc, ec, bc, mb, p(self.grad_mu if grad_mu is None else grad_mu), p(self.grad_value if grad_value is None else grad_value),
                        p(self.grad_log_std if grad_log_std is None else grad_log_std), p(self.out), p(self.partial))
        self._keep = (mu, value, log_std)
        _lib.check(self.lib.b2g_ppo_head(C.byref(a), _stream(self.device)), "b2g_ppo_head")

    def backward_direct(self, mu, value, log_std):
        """Loss + backward without the scalar detour: the head kernel leaves d loss / d(mu, value) in its buffers and writes
        d loss / d log_std straight into ``log_std.grad``; the network's backward starts from (mu, value).  Saves the clone, the
        ones-like seed, three scale kernels and the log_std accumulation of ``loss(...).backward()``.  ``self.out`` holds the losses."""
        assert log_std.grad is not None and log_std.grad.is_contiguous()
        # gradients of (mu, value) in tensors allocated on the calling stream: handing the autograd engine buffers that were allocated on
        # another stream (the head's persistent ones) makes it wait on that stream, which a CUDA-graph capture rejects
        g_mu, g_v = torch.empty_like(mu), torch.empty_like(value)
        self.launch(mu, value, log_std, grad_log_std=log_std.grad, grad_mu=g_mu, grad_value=g_v)
        _PpoHeadSeed.apply(mu, value, self, (g_mu, g_v)).backward()

    def loss(self, mu, value, log_std):
        """Scalar loss with autograd edges to mu, value, log_std; ``self.out`` then holds (loss, a_loss, c_loss, b_loss, kl, entropy)."""
        return _PpoHead.apply(mu, value, log_std, self)


class RolloutSampleArgs(C.Structure):
    _fields_ = [("mu", C.c_void_p), ("value", C.c_void_p), ("log_std", C.c_void_p), ("value_mean", C.c_void_p), ("value_var", C.c_void_p),
                ("value_eps", C.c_float), ("action_clip", C.c_float), ("n_envs", C.c_int32), ("n_actions", C.c_int32), ("t", C.c_int32),
                ("seed", C.c_uint64), ("rollout_counter", C.c_void_p), ("b_actions", C.c_void_p), ("b_mu", C.c_void_p), ("b_neglogp", C.c_void_p),
                ("b_values", C.c_void_p), ("env_actions", C.c_void_p)]


class RolloutPostArgs(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("done", C.c_void_p), ("time_out", C.c_void_p), ("flag_bytes", C.c_int32), ("n_envs", C.c_int32), ("t", C.c_int32),
                ("reward_scale", C.c_float), ("gamma", C.c_float), ("b_values", C.c_void_p), ("b_rewards", C.c_void_p), ("b_dones", C.c_void_p),
                ("ep_reward", C.c_void_p), ("ep_length", C.c_void_p), ("finished", C.c_void_p)]


class GaeArgs(C.Structure):
    _fields_ = [("rewards", C.c_void_p), ("values", C.c_void_p), ("dones", C.c_void_p), ("v_last", C.c_void_p), ("horizon", C.c_int32), ("n_envs", C.c_int32),
                ("gamma", C.c_float), ("tau", C.c_float), ("value_eps", C.c_float), ("value_mean", C.c_void_p), ("value_var", C.c_void_p),
                ("value_count", C.c_void_p), ("adv", C.c_void_p), ("ret", C.c_void_p), ("f_ret", C.c_void_p), ("f_val", C.c_void_p), ("f_adv", C.c_void_p),
                ("partial", C.c_void_p)]


class RolloutKernels:
    """The learner's per-step bookkeeping through ``libb200gym`` (``b2g_running_stat_update``, ``b2g_normalize_store``,
    ``b2g_rollout_sample``, ``b2g_rollout_post``, ``b2g_gae_finish``): operates on the PPO object's static rollout buffers."""

    def __init__(self, ppo, seed):
        self.lib = lib = _lib.load()
        vp = C.c_void_p
        lib.b2g_running_stat_update.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, vp]
        lib.b2g_stat_workspace_doubles.argtypes = [C.c_int, C.c_int]
        lib.b2g_normalize_store.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, vp]
        lib.b2g_rollout_sample.argtypes = [C.POINTER(RolloutSampleArgs), vp]
        lib.b2g_rollout_counter_advance.argtypes = [vp, vp]
        lib.b2g_rollout_post.argtypes = [C.POINTER(RolloutPostArgs), vp]
        lib.b2g_gae_finish.argtypes = [C.POINTER(GaeArgs), vp]
        lib.b2g_random_permutation.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp]
        for f in (lib.b2g_running_stat_update, lib.b2g_stat_workspace_doubles, lib.b2g_normalize_store, lib.b2g_rollout_sample,
                  lib.b2g_rollout_counter_advance, lib.b2g_rollout_post, lib.b2g_gae_finish, lib.b2g_random_permutation):
            f.restype = C.c_int
        self.perm = None
        self.perm_counter = 0
        self.p = ppo
        env, cfg = ppo.env, ppo.cfg
        dev = self.device = torch.device(ppo.device)
        N, T, F, A = env.num_envs, cfg.horizon_length, env.num_obs, env.num_acts
        self.N, self.T, self.F, self.A = N, T, F, A
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.counter = torch.zeros((), device=dev, dtype=torch.int64)
        self.mean_f = torch.zeros(F, device=dev)
        self.inv_std_f = torch.ones(F, device=dev)
        self.obs_ws = torch.zeros(int(lib.b2g_stat_workspace_doubles(N, F)), device=dev, dtype=torch.float64)
        self.gae_ws = torch.zeros(2 * int(lib.b2g_stat_workspace_doubles(T * N, 1)), device=dev, dtype=torch.float64)      # returns' and advantages' partial moments
        self.env_actions = torch.zeros(N, A, device=dev)
        self.mu = torch.zeros(N, A, device=dev)
        self.mu_scratch = torch.zeros(N, A, device=dev)          # action head of the critic instance of a separate network (unused)
        self.value = torch.zeros(N, device=dev)
        self.adv = torch.zeros(T, N, device=dev)
        self.ret = torch.zeros(T, N, device=dev)
        self.nobs_last = torch.zeros(N, F, device=dev)

    def _s(self):
        return _stream(self.device)

    def permutation(self, n):
        """A fresh pseudo-random permutation of 0..n-1 (int64, device): the mini-epoch shuffle, one launch instead of torch.randperm's sort."""
        if self.perm is None or self.perm.numel() != n:
            self.perm = torch.empty(n, dtype=torch.int64, device=self.device)
        self.perm_counter += 1
        _lib.check(self.lib.b2g_random_permutation(C.c_void_p(self.perm.data_ptr()), int(n), C.c_uint64(self.seed ^ 0x5DEECE66D), C.c_uint64(self.perm_counter), self._s()),
                   "b2g_random_permutation")
        return self.perm

    def obs_stats(self, obs):
        """obs_rms.update(obs) + float32 mean / inverse std for the normalisation of this step."""
        r = self.p.obs_rms
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.b2g_running_stat_update(p(obs), self.N, self.F, p(r.mean), p(r.var), p(r.count), p(self.obs_ws), p(self.mean_f), p(self.inv_std_f),
                                                    C.c_float(r.eps), self._s()), "b2g_running_stat_update")

    def normalize(self, obs, out, clip=5.0):
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.b2g_normalize_store(p(obs), p(self.mean_f), p(self.inv_std_f), p(out), self.N, self.F, C.c_float(clip), self._s()), "b2g_normalize_store")

    def sample(self, t, mu, value, log_std):
        P, p = self.p, (lambda x: C.c_void_p(x.data_ptr()))
        a = RolloutSampleArgs(p(mu), p(value), p(log_std), p(P.val_rms.mean), p(P.val_rms.var), float(P.val_rms.eps), 1.0, self.N, self.A, int(t), self.seed,
                              p(self.counter), p(P.b_act), p(P.b_mu), p(P.b_nlp), p(P.b_val), p(self.env_actions))
        _lib.check(self.lib.b2g_rollout_sample(C.byref(a), self._s()), "b2g_rollout_sample")

    def post(self, t, rew, done, time_outs):
        P, p = self.p, (lambda x: C.c_void_p(x.data_ptr()))
        if done.dtype != time_outs.dtype or done.dtype not in (torch.int64, torch.bool, torch.uint8):
            done, time_outs = done.to(torch.int64), time_outs.to(torch.int64)
        nbytes = 8 if done.dtype == torch.int64 else 1
        self._keep = (rew, done, time_outs)
        a = RolloutPostArgs(p(rew), p(done), p(time_outs), nbytes, self.N, int(t), float(P.cfg.reward_scale), float(P.cfg.gamma), p(P.b_val), p(P.b_rew),
                            p(P.b_done), p(P.ep_rew), p(P.ep_len), p(P.fin))
        _lib.check(self.lib.b2g_rollout_post(C.byref(a), self._s()), "b2g_rollout_post")

    def gae_finish(self, v_last):
        P, p = self.p, (lambda x: C.c_void_p(x.data_ptr()))
        r = P.val_rms
        a = GaeArgs(p(P.b_rew), p(P.b_val), p(P.b_done), p(v_last), self.T, self.N, float(P.cfg.gamma), float(P.cfg.tau), 1e-5, p(r.mean), p(r.var),
                    p(r.count), p(self.adv), p(self.ret), p(P.f_ret), p(P.f_val), p(P.f_adv), p(self.gae_ws))
        _lib.check(self.lib.b2g_gae_finish(C.byref(a), self._s()), "b2g_gae_finish")
        _lib.check(self.lib.b2g_rollout_counter_advance(C.c_void_p(self.counter.data_ptr()), self._s()), "counter")


_MM_OUT_OK = None      # does torch.mm(..., out_dtype=float32, out=...) work in this torch? (probed once)


def _mm_f32(a16, b16, out=None):
    """bf16 x bf16 -> float32 (tensor cores, fp32 accumulation), optionally into ``out``; older torch returns bf16, widened afterwards."""
    global _MM_OUT_OK
    if out is not None:
        if _MM_OUT_OK is None:
            try:
                torch.mm(a16, b16, out_dtype=torch.float32, out=out)
                _MM_OUT_OK = True
                return out
            except (TypeError, RuntimeError):
                _MM_OUT_OK = False
        if _MM_OUT_OK:
            return torch.mm(a16, b16, out_dtype=torch.float32, out=out)
        return out.copy_(_mm_f32(a16, b16))
    try:
        return torch.mm(a16, b16, out_dtype=torch.float32)
    except TypeError:
        return torch.mm(a16, b16).float()


class _LinearELU(torch.autograd.Function):
    """h = elu(x W^T + b): the GEMMs stay with torch / cuBLAS, bias + ELU is one in-place pass over the GEMM output, and the backward
    pass fuses ELU' with the bias gradient (``b2g_mlp_bias_elu`` / ``b2g_mlp_elu_backward``) -- instead of add, elu, elu_backward and a
    32768-row column reduction as four separate torch kernels.  With ``x16`` (a bf16 copy of the input) the weight gradient
    dW = dZ^T X -- the K = minibatch GEMM, the slowest of the three -- runs on bf16 operands with float32 accumulation and output
    (the reference trains with ``mixed_precision: True``, cfg/train/AnymalPPO.yaml:44); the two kernels write the bf16 copies of the
    activations and of dZ on the way, so no separate cast pass exists.  Returns (h, bf16 copy of h or None)."""

    @staticmethod
    def forward(ctx, x, weight, bias, x16, direct):
        lib = _lib.load()
        ctx.direct = bool(direct)
        ctx.set_materialize_grads(False)      # the bf16 copy is not differentiable: no zero tensor for its gradient
        h = torch.mm(x, weight.t())
        h16 = torch.empty_like(h, dtype=torch.bfloat16) if x16 is not None else None
        _lib.check(lib.b2g_mlp_bias_elu(C.c_void_p(h.data_ptr()), C.c_void_p(bias.data_ptr()), h.shape[0], h.shape[1],
                                        C.c_void_p(h16.data_ptr()) if h16 is not None else None, _stream(h.device)), "b2g_mlp_bias_elu")
        ctx.save_for_backward(x, weight, h, x16, bias)
        ctx.mark_non_differentiable(*([h16] if h16 is not None else []))
        return h, h16

    @staticmethod
    def backward(ctx, dh, _unused=None):
        x, weight, h, x16, bias = ctx.saved_tensors
        lib = _lib.load()
        dh = dh.contiguous()
        rows, cols = h.shape
        need_dx = ctx.needs_input_grad[0]
        direct = ctx.direct and weight.grad is not None and bias.grad is not None      # store into the .grad views, return no gradient
        dz16 = torch.empty_like(h, dtype=torch.bfloat16) if x16 is not None else None
        dz = torch.empty_like(h) if (need_dx or dz16 is None) else None               # first layer with a bf16 copy: nobody reads the float32 dz
        db = bias.grad if direct else torch.empty(cols, device=h.device, dtype=h.dtype)
        ws = torch.empty(int(lib.b2g_mlp_elu_backward_workspace_floats(rows, cols)), device=h.device, dtype=h.dtype)
        _lib.check(lib.b2g_mlp_elu_backward(C.c_void_p(dh.data_ptr()), C.c_void_p(h.data_ptr()), C.c_void_p(dz.data_ptr()) if dz is not None else None, C.c_void_p(db.data_ptr()),
                                            C.c_void_p(ws.data_ptr()), rows, cols, C.c_void_p(dz16.data_ptr()) if dz16 is not None else None,
                                            _stream(h.device)), "b2g_mlp_elu_backward")
        dx = torch.mm(dz, weight) if need_dx else None
        dw = None
        if ctx.needs_input_grad[1]:
            if direct:
                _mm_f32(dz16.t(), x16, out=weight.grad) if x16 is not None else torch.mm(dz.t(), x, out=weight.grad)
            else:
                dw = _mm_f32(dz16.t(), x16) if x16 is not None else torch.mm(dz.t(), x)
        return dx, dw, (None if direct else db), None, None


def _protos():
    lib = _lib.load()
    if not getattr(lib, "_b2g_mlp_protos", False):
        vp = C.c_void_p
        lib.b2g_mlp_bias_elu.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        lib.b2g_mlp_bias_elu.restype = C.c_int
        lib.b2g_mlp_elu_backward.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
        lib.b2g_mlp_elu_backward.restype = C.c_int
        lib.b2g_mlp_elu_backward_workspace_floats.argtypes = [C.c_int, C.c_int]
        lib.b2g_mlp_elu_backward_workspace_floats.restype = C.c_int
        lib.b2g_mlp_heads_forward.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        lib.b2g_mlp_heads_forward.restype = C.c_int
        lib.b2g_mlp_heads_backward.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        lib.b2g_mlp_heads_backward.restype = C.c_int
        lib.b2g_mlp_heads_backward_workspace_floats.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.b2g_mlp_heads_backward_workspace_floats.restype = C.c_int
        lib.b2g_mlp_heads_backward_scatter.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        lib.b2g_mlp_heads_backward_scatter.restype = C.c_int
        lib.b2g_gather_rows.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
        lib.b2g_gather_rows.restype = C.c_int
        lib._b2g_mlp_protos = True
    return lib


def linear_elu(x, weight, bias, x16=None, direct_grad=False):
    """``F.elu(F.linear(x, weight, bias))`` through the fused kernels (CUDA float32, width a multiple of 4).  Returns (h, h16): h16 is a
    bf16 copy of h when a bf16 copy of the input (``x16``) was given -- pass it on to the next layer.  ``direct_grad``: the backward
    pass STORES d weight / d bias into the parameters' existing ``.grad`` tensors (overwrite, not accumulate: every parameter used once
    per backward, as in the PPO minibatch step) and hands autograd no gradient for them -- no accumulation kernels."""
    _protos()
    return _LinearELU.apply(x.contiguous(), weight, bias, x16, direct_grad)


class _Heads(torch.autograd.Function):
    """(mu, value) = (h W_mu^T + b_mu, h W_v^T + b_v); the backward of both heads is one kernel (``b2g_mlp_heads_backward``)."""

    @staticmethod
    def forward(ctx, h, w_mu, b_mu, w_v, b_v, direct):
        ctx.direct = bool(direct)
        lib = _lib.load()
        rows, hid = h.shape
        A = w_mu.shape[0]
        mu = torch.empty(rows, A, device=h.device, dtype=h.dtype)
        v = torch.empty(rows, device=h.device, dtype=h.dtype)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.b2g_mlp_heads_forward(p(h), p(w_mu), p(b_mu), p(w_v), p(b_v), rows, hid, A, p(mu), p(v), _stream(h.device)), "b2g_mlp_heads_forward")
        ctx.save_for_backward(h, w_mu, w_v, b_mu, b_v)
        return mu, v

    @staticmethod
    def backward(ctx, dmu, dv):
        h, w_mu, w_v, b_mu, b_v = ctx.saved_tensors
        lib = _lib.load()
        rows, hid = h.shape
        A = w_mu.shape[0]
        dmu, dv = dmu.contiguous(), dv.contiguous()
        dh = torch.empty_like(h)
        ws = torch.empty(int(lib.b2g_mlp_heads_backward_workspace_floats(rows, hid, A)), device=h.device, dtype=h.dtype)
        p = lambda t: C.c_void_p(t.data_ptr())
        if ctx.direct and all(t.grad is not None for t in (w_mu, b_mu, w_v)):      # store into the .grad views (see linear_elu)
            # ... except the one-element value bias, which still goes back through autograd: with NO gradient accumulation node left
            # in the backward pass, capturing it into a CUDA graph fails (torch 2.11: "dependency created on uncaptured work in
            # another stream"); one 1-element accumulation keeps the engine on the path it takes for every ordinary network
            db_v = torch.empty_like(b_v)
            _lib.check(lib.b2g_mlp_heads_backward_scatter(p(h), p(dmu), p(dv), p(w_mu), p(w_v), rows, hid, A, p(dh), p(w_mu.grad), p(b_mu.grad),
                                                          p(w_v.grad), p(db_v), p(ws), _stream(h.device)), "b2g_mlp_heads_backward_scatter")
            return dh, None, None, None, db_v, None
        cat = torch.empty(A + 1, hid + 1, device=h.device, dtype=h.dtype)
        _lib.check(lib.b2g_mlp_heads_backward(p(h), p(dmu), p(dv), p(w_mu), p(w_v), rows, hid, A, p(dh), p(cat), p(ws), _stream(h.device)),
                   "b2g_mlp_heads_backward")
        return dh, cat[:A, :hid], cat[:A, hid], cat[A:, :hid], cat[A:, hid], None


def heads(h, w_mu, b_mu, w_v, b_v, direct_grad=False):
    """Both output heads of the actor-critic on the last hidden layer (shared-trunk networks: one ``h`` feeds both)."""
    _protos()
    return _Heads.apply(h.contiguous(), w_mu, b_mu, w_v, b_v, direct_grad)


def gather_rows(src, index, out=None, out_bf16=None):
    """``src[index]`` for a (rows, cols) float32 CUDA matrix (cols % 4 == 0) into ``out`` (float32) and / or ``out_bf16`` in one pass."""
    lib = _protos()
    rows, cols = int(index.numel()), int(src.shape[1])
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.b2g_gather_rows(p(src), p(index), rows, cols, p(out), p(out_bf16), _stream(src.device)), "b2g_gather_rows")
    return out, out_bf16


class FlatParameters:
    """Re-homes every parameter of ``module`` into one flat float32 buffer (``.data`` becomes a view) and gives each a persistent
    ``.grad`` view of a second flat buffer: autograd accumulates straight into it (use ``zero_()`` on :attr:`grad`, never
    ``zero_grad(set_to_none=True)``)."""

    ALIGN = 32      # floats: every parameter starts on a 128-byte boundary (the kernels read rows of them as float4)

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        pad = lambda k: (k + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        n = sum(pad(p.numel()) for p in params)
        dev = params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)       # the gaps stay zero: zero gradient, zero Adam update
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            off += pad(k)
        self.params, self.n = params, n


class FusedClipAdam:
    """``clip_grad_norm_`` + ``torch.optim.Adam`` (no weight decay, no amsgrad) on a :class:`FlatParameters` vector through
    ``b2g_adam_clip_step``."""

    def __init__(self, flat: FlatParameters, lr, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0, grad_scale=1.0):
        self.lib = _lib.load()
        self.lib.b2g_adam_clip_step.argtypes = [C.POINTER(AdamArgs), C.c_void_p]
        self.lib.b2g_adam_clip_step.restype = C.c_int
        self.flat = flat
        dev = flat.flat.device
        self.device = dev
        self.lr = lr if torch.is_tensor(lr) else torch.tensor(float(lr), device=dev, dtype=torch.float32)
        self.step_count = torch.zeros((), device=dev, dtype=torch.int64)
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.partial = torch.zeros(ADAM_MAX_PARTIALS, device=dev)
        self.norm = torch.zeros(2, device=dev)          # gradient norm before clipping, clip coefficient
        self.betas, self.eps, self.max_grad_norm, self.grad_scale = betas, float(eps), float(max_grad_norm), float(grad_scale)
        # a param_groups-shaped view so that code written against torch optimisers (learning-rate schedules) keeps working
        self.param_groups = [{"lr": self.lr, "params": flat.params}]

    def step(self):
        p = lambda t: C.c_void_p(t.data_ptr())
        f = self.flat
        a = AdamArgs(p(f.flat), p(f.grad), p(self.exp_avg), p(self.exp_avg_sq), f.n, p(self.lr), p(self.step_count), self.betas[0], self.betas[1],
                     self.eps, self.max_grad_norm, self.grad_scale, p(self.partial), p(self.norm))
        _lib.check(self.lib.b2g_adam_clip_step(C.byref(a), _stream(self.device)), "b2g_adam_clip_step")

    def zero_grad(self, set_to_none=False):
        self.flat.grad.zero_()

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": int(self.step_count), "lr": float(self.lr)}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count.fill_(int(sd["step"]))
