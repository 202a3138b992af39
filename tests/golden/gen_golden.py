#!/usr/bin/env python
"""Generate golden vectors from the reference's own @torch.jit.script task functions.

Runs ONLY in the build container (needs /root/reference, imported through ref_loader's stubs).
Inputs are RNG-free (sin-based), so the fixtures do not depend on a torch RNG version.  Outputs are
committed as tests/golden/*.npz; tests compare the numpy oracle and the CUDA kernels against them.

    python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def sinfill(shape, a, b, scale=1.0):
    n = int(np.prod(shape))
    x = np.sin(a * np.arange(n, dtype=np.float64) + b) * scale
    return torch.tensor(x.reshape(shape), dtype=torch.float32)


def make_root(n, a=0.37, b=0.1, z0=0.55):
    root = sinfill((n, 13), a, b)
    root[:, 3:7] = root[:, 3:7] / root[:, 3:7].norm(dim=1, keepdim=True)
    root[:, 7:13] *= 1.5
    root[:, 2] = z0 + 0.1 * root[:, 2]
    return root


def gen_anymal(task_mod, prefix, n_bodies, knee, base, out):
    n = 64
    root = make_root(n)
    commands = sinfill((n, 3), 0.91, 0.2) * torch.tensor([2.0, 1.0, 1.0])
    torques = sinfill((n, 12), 0.43, 0.6, 40.0)
    contact = sinfill((n, n_bodies, 3), 0.67, 0.7)
    # per-env magnitude so that some envs sit below, near and above the 1 N threshold
    mag = torch.tensor([0.3, 0.55, 0.57, 0.58, 1.2, 0.3, 0.9, 0.05])[torch.arange(n) % 8]
    contact = contact * mag[:, None, None]
    progress = torch.tensor([0, 10, 2497, 2498, 2499, 2500, 1, 300], dtype=torch.int64)[torch.arange(n) % 8]
    progress = torch.roll(progress, 3)  # decorrelate from the contact magnitudes
    knee_idx = torch.tensor(knee, dtype=torch.int64)
    scales = {"lin_vel_xy": 1.0 * 0.02, "ang_vel_z": 0.5 * 0.02, "torque": -0.000025 * 0.02}
    rew_fn = getattr(task_mod, f"compute_{prefix}_reward")
    obs_fn = getattr(task_mod, f"compute_{prefix}_observations")
    rew, reset = rew_fn(root, commands, torques, contact, knee_idx, progress, scales, base, 2500)
    dof_pos = sinfill((n, 12), 0.29, 0.3, 0.8)
    default = sinfill((1, 12), 1.3, 0.5, 0.6).repeat(n, 1)
    dof_vel = sinfill((n, 12), 0.53, 0.9, 8.0)
    grav = torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1)
    actions = sinfill((n, 12), 0.77, 1.1)
    obs = obs_fn(root, commands, dof_pos, default, dof_vel, grav, actions, 2.0, 0.25, 1.0, 0.05)
    np.savez(out, root=root.numpy(), commands=commands.numpy(), torques=torques.numpy(), contact=contact.numpy(),
             progress=progress.numpy(), knee=np.array(knee), base=np.array(base), max_len=np.array(2500),
             scale_lin=np.float32(scales["lin_vel_xy"]), scale_ang=np.float32(scales["ang_vel_z"]), scale_torque=np.float32(scales["torque"]),
             rew=rew.numpy(), reset=reset.numpy(), dof_pos=dof_pos.numpy(), default=default.numpy(), dof_vel=dof_vel.numpy(),
             actions=actions.numpy(), obs=obs.numpy(), obs_scales=np.array([2.0, 0.25, 1.0, 0.05], dtype=np.float32))
    print(out, "reset count", int(reset.sum()), "rew range", float(rew.min()), float(rew.max()))


def gen_cartpole(out):
    mod = ref_loader.load("tasks.cartpole")
    n = 64
    ang = sinfill((n,), 0.7, 0.1, 2.0)
    pv = sinfill((n,), 0.31, 0.4, 6.0)
    cv = sinfill((n,), 0.57, 0.2, 3.0)
    cp = sinfill((n,), 0.23, 0.8, 3.5)
    reset_buf = torch.zeros(n, dtype=torch.int64)
    reset_buf[::7] = 1
    progress = (torch.arange(n, dtype=torch.int64) * 9) % 503
    rew, reset = mod.compute_cartpole_reward(ang, pv, cv, cp, 3.0, reset_buf, progress, 500.0)
    np.savez(out, pole_angle=ang.numpy(), pole_vel=pv.numpy(), cart_vel=cv.numpy(), cart_pos=cp.numpy(), reset_buf=reset_buf.numpy(),
             progress=progress.numpy(), rew=rew.numpy(), reset=reset.numpy())
    print(out, "reset count", int(reset.sum()))


def gen_utils(out):
    tj = ref_loader.load("utils.torch_jit_utils")
    at = ref_loader.load("tasks.anymal_terrain")
    n = 32
    q = sinfill((n, 4), 0.61, 0.3)
    q = q / q.norm(dim=1, keepdim=True)
    q2 = sinfill((n, 4), 0.47, 1.3)
    q2 = q2 / q2.norm(dim=1, keepdim=True)
    v = sinfill((n, 3), 0.83, 0.5, 2.0)
    ang = torch.tensor([0, 3.5, -3.5, 7, -7, -0.1, 3.1415927, -3.1415927, 6.2831855, 100.0, -100.0, 1e-6])
    np.savez(out, q=q.numpy(), q2=q2.numpy(), v=v.numpy(),
             quat_rotate=tj.quat_rotate(q, v).numpy(), quat_rotate_inverse=tj.quat_rotate_inverse(q, v).numpy(),
             quat_apply=tj.quat_apply(q, v).numpy(), quat_mul=tj.quat_mul(q, q2).numpy(), normalize=tj.normalize(v).numpy(),
             quat_apply_yaw=at.quat_apply_yaw(q.clone(), v).numpy(), angles=ang.numpy(), wrap_to_pi=at.wrap_to_pi(ang.clone()).numpy(),
             rand_u=sinfill((n, 3), 0.2, 0.1).abs().numpy(),
             torch_rand_float=((1.5 - 0.5) * sinfill((n, 3), 0.2, 0.1).abs() + 0.5).numpy())
    print(out)


if __name__ == "__main__":
    torch.set_num_threads(1)
    gen_anymal(ref_loader.load("tasks.anymal"), "anymal", 13, [2, 5, 8, 11], 0, os.path.join(HERE, "anymal_flat.npz"))
    gen_anymal(ref_loader.load("tasks.hound"), "hound", 17, [2, 6, 10, 14], 0, os.path.join(HERE, "hound_flat.npz"))
    gen_cartpole(os.path.join(HERE, "cartpole.npz"))
    gen_utils(os.path.join(HERE, "jit_utils.npz"))
