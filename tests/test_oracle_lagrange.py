"""Contact-free forward dynamics of the oracle against an INDEPENDENT derivation: the Euler-Lagrange equations of the same articulation,
formed by automatic differentiation (torch.func, float64) of a Lagrangian that is built from nothing but forward kinematics,

    L(q, qd) = sum_links [ 1/2 m |d/dt p_com|^2 + 1/2 w . (R I R^T) w ] + sum_dofs 1/2 armature qd^2 - sum_links m g z_com,
    w from (d/dt R) R^T,      M = d2L/dqd2,      M qdd = tau - (d(dL/dqd)/dq) qd + dL/dq.

No spatial algebra, no recursion, no articulated-body inertias: a different principle (variational, not Newton-Euler) and different code than
the oracle's ABA / CRBA / RNEA, which only check each other (tests/test_oracle_dynamics.py).  The floating base enters through local
coordinates (position, rotation vector about the current attitude): at rotation vector 0 its second derivative IS the angular acceleration
(the Jacobian of the exponential map is I + theta^ / 2, whose time derivative applied to theta' vanishes), so the six base rows compare with
the oracle's root acceleration directly.

north_star asks for "joint accelerations 1e-3 relative, contact-free" against PhysX; PhysX is a closed binary that is not here,
so this is the external pin that exists: rigid-body mechanics itself.  The CUDA kernels are held to the oracle by tests/test_kernels_gpu.py
(check_forward_dynamics), which closes the chain kernel -> oracle -> Euler-Lagrange.
"""
import numpy as np
import pytest
import torch

from isaacgymenv_b200 import _abi
from oracle import dyn_oracle as O
from tests.kernel_checks import flat_params, load_robot, random_flying_state

JOINT_REVOLUTE = 0      # model/urdf.py


def _quat_to_mat(q):
    x, y, z, w = q[0], q[1], q[2], q[3]
    return torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
                        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)]),
                        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)])])


def _hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def _axis_angle(axis, ang):
    k = _hat(axis)
    return torch.eye(3, dtype=axis.dtype) + torch.sin(ang) * k + (1 - torch.cos(ang)) * (k @ k)


class Lagrange:
    """Euler-Lagrange dynamics of a packed model (`_abi.Model`: the float32-rounded parameters the oracle and the kernels see)."""

    def __init__(self, art, m, gravity):
        t = lambda a: torch.tensor(np.array(a, dtype=np.float64))
        nd, nl = art.num_dofs, art.num_links
        self.nd, self.nl, self.fixed = nd, nl, bool(art.fixed_base)
        self.parent = [int(p) for p in art.link_parent]
        self.mass = t(list(m.link_mass)[:nl])
        self.com = t([list(m.link_com[l]) for l in range(nl)])
        i6 = np.array([list(m.link_inertia[l]) for l in range(nl)], dtype=np.float64)
        self.inertia = t([[[a[0], a[3], a[4]], [a[3], a[1], a[5]], [a[4], a[5], a[2]]] for a in i6])
        self.jtype = [int(m.joint_type[d]) for d in range(nd)]
        self.jpos = t([list(m.joint_pos[d]) for d in range(nd)])
        self.jrot = [_quat_to_mat(t(list(m.joint_quat[d]))) for d in range(nd)]
        self.jaxis = t([list(m.joint_axis[d]) for d in range(nd)])
        self.armature = t(list(m.armature)[:nd])
        self.g = t(gravity)

    def fk(self, x, root_pos, root_rot):
        """Generalised coordinates x = (base translation (3), base rotation vector (3), joints) -- the base part absent for a fixed base --
        -> COM positions (nl, 3) and link rotations (nl, 3, 3) in the world."""
        if self.fixed:
            q, p0, r0 = x, root_pos, root_rot
        else:
            q = x[6:]
            # exp(theta^) about WORLD axes on top of the current attitude, as its power series: the derivatives this test takes are evaluated
            # at theta = 0 and reach the second order, where sin(a) / a = 1 - a^2/6 + a^4/120 and (1 - cos a) / a^2 = 1/2 - a^2/24 + a^4/720
            # are exact; the closed form's a = |theta| is not differentiable there
            th = x[3:6]
            s2 = th @ th
            k = _hat(th)
            r0 = (torch.eye(3, dtype=th.dtype) + (1 - s2 / 6 + s2 * s2 / 120) * k + (0.5 - s2 / 24 + s2 * s2 / 720) * (k @ k)) @ root_rot
            p0 = root_pos + x[:3]
        pos, rot = [p0], [r0]
        for d in range(self.nd):
            par = self.parent[d + 1]
            rj = rot[par] @ self.jrot[d]
            pj = pos[par] + rot[par] @ self.jpos[d]
            if self.jtype[d] == JOINT_REVOLUTE:
                rot.append(rj @ _axis_angle(self.jaxis[d], q[d]))
                pos.append(pj)
            else:
                rot.append(rj)
                pos.append(pj + rj @ (self.jaxis[d] * q[d]))
        rot = torch.stack(rot)
        com = torch.stack(pos) + torch.einsum("lij,lj->li", rot, self.com)
        return com, rot

    def lagrangian(self, x, xd, root_pos, root_rot):
        (com, rot), (comd, rotd) = torch.func.jvp(lambda y: self.fk(y, root_pos, root_rot), (x,), (xd,))
        W = rotd @ rot.transpose(-1, -2)
        w = torch.stack([W[:, 2, 1], W[:, 0, 2], W[:, 1, 0]], dim=-1)
        iw = rot @ self.inertia @ rot.transpose(-1, -2)
        qd = xd if self.fixed else xd[6:]
        kinetic = 0.5 * (self.mass * (comd * comd).sum(-1)).sum() + 0.5 * torch.einsum("li,lij,lj->", w, iw, w) + 0.5 * (self.armature * qd * qd).sum()
        potential = -(self.mass * (com @ self.g)).sum()
        return kinetic - potential, kinetic, potential

    def accelerations(self, root13, dof, tau):
        """(generalised accelerations, kinetic energy): base rows (linear, angular) first for a floating base."""
        root13, dof = torch.tensor(root13, dtype=torch.float64), torch.tensor(dof, dtype=torch.float64)
        root_pos, root_rot = root13[:3], _quat_to_mat(root13[3:7])
        if self.fixed:
            x, xd, f = dof[:, 0].clone(), dof[:, 1].clone(), torch.tensor(tau, dtype=torch.float64)
        else:
            x = torch.cat([torch.zeros(6, dtype=torch.float64), dof[:, 0]])
            xd = torch.cat([root13[7:10], root13[10:13], dof[:, 1]])
            f = torch.cat([torch.zeros(6, dtype=torch.float64), torch.tensor(tau, dtype=torch.float64)])
        lag = lambda a, b: self.lagrangian(a, b, root_pos, root_rot)[0]
        momentum = torch.func.grad(lag, argnums=1)
        M = torch.func.jacrev(momentum, argnums=1)(x, xd)
        dp_dx = torch.func.jacrev(momentum, argnums=0)(x, xd)
        dl_dx = torch.func.grad(lag, argnums=0)(x, xd)
        xdd = torch.linalg.solve(M, f - dp_dx @ xd + dl_dx)
        return xdd.numpy(), float(self.lagrangian(x, xd, root_pos, root_rot)[1]), M.numpy()


# robots whose joint frames are axis-aligned in the URDF: their float32-rounded joint quaternions and axes are exactly unit, and the two
# derivations agree to float64 rounding.  The others (rpy-rotated joint frames) carry quaternions / axes that are unit only to float32
# precision, which the recursion and the exponential-map kinematics turn into slightly different (1e-7) rotations.
EXACT_FRAMES = {"hound", "useful_hound", "houndarm"}


@pytest.mark.parametrize("robot", ["anymal", "hound", "useful_hound", "houndarm", "manipulator", "cartpole"])
def test_aba_equals_euler_lagrange(robot):
    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    rng = np.random.default_rng(11)
    n = 3
    root, dof = random_flying_state(art, n, rng)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    root[:, 3:7] /= np.linalg.norm(root[:, 3:7], axis=1, keepdims=True)      # unit in float64, not only in float32
    tau = rng.normal(size=(n, art.num_dofs)) * 10
    qdd, a0 = O.forward_dynamics(m, sp, root, dof, tau)
    L = Lagrange(art, m, [sp.gravity[0], sp.gravity[1], sp.gravity[2]])
    tol = 1e-9 if robot in EXACT_FRAMES else 2e-6
    for e in range(n):
        xdd, ke, M = L.accelerations(root[e], dof[e], tau[e])
        assert np.abs(M - M.T).max() < 1e-9 * np.abs(M).max() and np.linalg.eigvalsh(M).min() > 0
        ke_o, _, _ = O.energy_momentum(m, sp, root[e], dof[e])
        assert abs(ke - ke_o) < max(tol, 1e-12) * max(1.0, abs(ke_o)), (ke, ke_o)
        nb = 0 if art.fixed_base else 6
        scale = max(1.0, float(np.abs(qdd[e]).max()))
        assert np.abs(xdd[nb:] - qdd[e]).max() < tol * scale, (robot, e, np.abs(xdd[nb:] - qdd[e]).max(), scale)
        if nb:
            # the oracle reports the root's SPATIAL acceleration (Featherstone), angular part first: the angular part is the classical
            # angular acceleration, the linear part is d/dt(velocity of the root origin) - w x v
            v, w = root[e][7:10], root[e][10:13]
            assert np.abs(xdd[3:6] - a0[e][:3]).max() < tol * scale, (robot, e, "angular")
            assert np.abs(xdd[:3] - np.cross(w, v) - a0[e][3:]).max() < tol * scale, (robot, e, "linear")


@pytest.mark.parametrize("robot", ["anymal", "hound", "useful_hound", "houndarm", "manipulator", "cartpole"])
def test_kernel_code_equals_euler_lagrange(robot):
    """The same comparison without the oracle in between: the step kernels' own dynamics code (float32, run on the host lane emulator --
    tests/emu compiles csrc/b2g_dynamics.cuh for host threads in lock-step) against the float64 Euler-Lagrange accelerations, at the
    tolerance north_star names for contact-free single-step dynamics (1e-3 relative; measured 2e-7 ... 5e-5, asserted 2e-4).  The GPU build of the same code is held to the oracle
    at the same tolerance by tests/test_kernels_gpu.py."""
    from tests.backends import EmuBackend

    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    props = _abi.default_dof_props(art)
    rng = np.random.default_rng(5)
    n = 4
    root, dof = random_flying_state(art, n, rng)
    q = root[:, 3:7].astype(np.float64)
    tau = (rng.normal(size=(n, art.num_dofs)) * 20).astype(np.float32)
    be = EmuBackend(art, sp, props, n)
    try:
        be.set_state(root, dof)
        qdd, a0 = be.forward_dynamics(tau)
    finally:
        be.close()
    L = Lagrange(art, m, [sp.gravity[0], sp.gravity[1], sp.gravity[2]])
    for e in range(n):
        r = root[e].astype(np.float64)
        r[3:7] = q[e] / np.linalg.norm(q[e])
        xdd, _, _ = L.accelerations(r, dof[e].astype(np.float64), tau[e].astype(np.float64))
        nb = 0 if art.fixed_base else 6
        err = np.abs(xdd[nb:] - qdd[e]).max() / np.abs(xdd[nb:]).max()
        assert err < 2e-4, (robot, e, err)
        if nb:
            ang = np.abs(xdd[3:6] - a0[e][:3]).max() / max(1.0, np.abs(xdd[3:6]).max())
            lin = np.abs(xdd[:3] - np.cross(r[10:13], r[7:10]) - a0[e][3:]).max() / max(1.0, np.abs(xdd[:3]).max())
            assert ang < 2e-4 and lin < 2e-4, (robot, e, ang, lin)


@pytest.mark.parametrize("robot", ["houndarm", "manipulator"])
def test_mass_matrix_equals_the_lagrangian_hessian(robot):
    """The joint-space mass matrix the arm tasks' operational-space law inverts (acquire_mass_matrix_tensor; oracle: CRBA) is the Hessian of
    the kinetic energy in the joint velocities, and the bias the RNEA returns is the rest of the Euler-Lagrange equation."""
    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    rng = np.random.default_rng(3)
    root, dof = random_flying_state(art, 2, rng)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    root[:, 3:7] /= np.linalg.norm(root[:, 3:7], axis=1, keepdims=True)
    L = Lagrange(art, m, [sp.gravity[0], sp.gravity[1], sp.gravity[2]])
    tol = 1e-9 if robot in EXACT_FRAMES else 2e-6
    for e in range(2):
        H, Cb = O.crba_rnea(m, sp, root[e], dof[e])
        xdd, _, M = L.accelerations(root[e], dof[e], np.zeros(art.num_dofs))
        assert np.abs(H - M).max() < tol * np.abs(M).max()
        np.testing.assert_allclose(-M @ xdd, Cb, rtol=0, atol=tol * max(1.0, np.abs(Cb).max()))      # tau = 0: M qdd = -bias
