"""Torch versions of the quaternion / sampling helpers the task code uses on its generic (un-fused)
path.  Same conventions as the reference's ``utils/torch_jit_utils.py`` (xyzw quaternions; ``quat_rotate``
:80-90, ``quat_rotate_inverse`` :93-103, ``quat_apply`` :70-77, ``normalize`` :65-67, ``torch_rand_float``
:215-218, ``get_axis_params`` :157-165) and ``tasks/anymal_terrain.py:676-687`` (``quat_apply_yaw``,
``wrap_to_pi`` with C ``fmod`` semantics).  The fused CUDA kernels carry their own copies of this math.
"""
import math

import numpy as np
import torch


def to_torch(x, dtype=torch.float, device="cuda:0", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def _rot(q, v, sign):
    w = q[:, 3:4]
    qv = q[:, :3]
    a = v * (2.0 * w * w - 1.0)
    b = torch.cross(qv, v, dim=-1) * w * 2.0
    c = qv * (qv * v).sum(dim=-1, keepdim=True) * 2.0
    return a + sign * b + c


def quat_rotate(q, v):
    return _rot(q, v, 1.0)


def quat_rotate_inverse(q, v):
    return _rot(q, v, -1.0)


def quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = torch.cross(xyz, b, dim=-1) * 2
    return (b + a[:, 3:] * t + torch.cross(xyz, t, dim=-1)).view(shape)


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def quat_apply_yaw(quat, vec):
    q = quat.clone().view(-1, 4)
    q[:, :2] = 0.0
    return quat_apply(normalize(q), vec)


def wrap_to_pi(angles):
    """fmod (sign of the dividend), then fold values above pi -- what the reference's TorchScript executes."""
    a = torch.fmod(angles, 2 * math.pi)
    return a - 2 * math.pi * (a > math.pi)


def torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower


def tensor_clamp(t, min_t, max_t):
    return torch.max(torch.min(t, max_t), min_t)


def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    zs = np.zeros((n_dims,))
    zs[axis_idx] = 1.0
    params = np.where(zs == 1.0, value, zs)
    params[0] = x_value
    return list(params.astype(dtype))
