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


@pytest.mark.parametrize("robot", ["houndarm", "manipulator", "useful_hound"])
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
        if art.fixed_base:
            assert np.abs(H - M).max() < tol * np.abs(M).max()
            np.testing.assert_allclose(-M @ xdd, Cb, rtol=0, atol=tol * max(1.0, np.abs(Cb).max()))      # tau = 0: M qdd = -bias
        else:
            # floating base (UsefulHound: the arm block of this matrix is what its operational-space law inverts, useful_hound.py:455, `mm[:, -6:, -6:]`):
            # the joint-joint block does not depend on how the six base coordinates are ordered or expressed
            assert np.abs(H[6:, 6:] - M[6:, 6:]).max() < tol * np.abs(M[6:, 6:]).max()


def _quat_mul(a, b):
    x1, y1, z1, w1 = a
    x2, y2, z2, w2 = b
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                     w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


def _rk4(m, sp, root, dof, torque_of, horizon, h):
    """Classical Runge-Kutta on the equations of motion whose right-hand side the tests above pin to the Euler-Lagrange equations."""

    def rhs(r, d):
        qdd, a0 = O.forward_dynamics(m, sp, r[None], d[None], torque_of(d)[None])
        v, w = r[7:10], r[10:13]
        dr, dd = np.zeros(13), np.zeros_like(d)
        dr[:3] = v
        dr[3:7] = 0.5 * _quat_mul(np.array([w[0], w[1], w[2], 0.0]), r[3:7])
        dr[7:10] = a0[0][3:] + np.cross(w, v)      # spatial -> classical acceleration of the root origin
        dr[10:13] = a0[0][:3]
        dd[:, 0], dd[:, 1] = d[:, 1], qdd[0]
        return dr, dd

    root, dof = root.copy(), dof.copy()
    for _ in range(int(round(horizon / h))):
        k1 = rhs(root, dof)
        k2 = rhs(root + 0.5 * h * k1[0], dof + 0.5 * h * k1[1])
        k3 = rhs(root + 0.5 * h * k2[0], dof + 0.5 * h * k2[1])
        k4 = rhs(root + h * k3[0], dof + h * k3[1])
        root = root + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        dof = dof + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        root[3:7] /= np.linalg.norm(root[3:7])
    return root, dof


@pytest.mark.parametrize("robot,mode", [("anymal", "pd"), ("anymal", "effort"), ("hound", "pd")])
def test_simulate_is_a_first_order_integrator_of_the_verified_equations(robot, mode):
    """`simulate` (semi-implicit sub-steps, implicit PD drives) over north_star's 10-step horizon, contact-free, against a 4th-order
    integration of the pinned equations of motion with the same drive law: the state error halves with the step -- a consistent first-order
    scheme, no missing or doubled term -- and at the production step (dt 0.02 s, 2 sub-steps) ends at 1-2e-2 rad / 1e-2 m after 0.2 s,
    the 1/2 g T h position lag of semi-implicit Euler included."""
    from tests.kernel_checks import default_pose

    art = load_robot(robot)
    m = _abi.pack_model(art)
    rng = np.random.default_rng(2)
    root, dof = random_flying_state(art, 1, rng, scale_qd=0.5)
    root, dof = root[0].astype(np.float64), dof[0].astype(np.float64)
    root[3:7] /= np.linalg.norm(root[3:7])
    root[7:13] *= 0.5
    nd = art.num_dofs
    target = default_pose(art) + rng.uniform(-0.2, 0.2, nd)
    tau_c = rng.normal(size=nd) * 0.2
    kp, kd = 85.0, 2.0      # cfg/task/Anymal.yaml:55-56
    if mode == "pd":
        props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, kp, kd)
        lim = np.array(art.effort, dtype=np.float64)
        torque_of = lambda d: np.clip(kp * (target - d[:, 0]) - kd * d[:, 1], -lim, lim)
    else:
        props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
        torque_of = lambda d: tau_c
    horizon = 0.2
    ref_root, ref_dof = _rk4(m, flat_params(ground=False), root, dof, torque_of, horizon, 1e-3)
    errs = []
    for dt in (0.02, 0.01, 0.005):
        sp = flat_params(dt=dt, substeps=2, ground=False)
        r, d = root[None].copy(), dof[None].copy()
        for _ in range(int(round(horizon / dt))):
            O.simulate(m, sp, props, r, d, target[None] if mode == "pd" else np.zeros((1, nd)), tau_c[None] if mode == "effort" else np.zeros((1, nd)))
        errs.append((np.abs(d[0][:, 0] - ref_dof[:, 0]).max(), np.abs(r[0][:3] - ref_root[:3]).max(), np.abs(r[0][10:13] - ref_root[10:13]).max()))
    (q0, p0, w0), (q1, p1, w1), (q2, p2, w2) = errs
    assert q0 < 3e-2 and p0 < 1.2e-2, errs                       # production step, 10 steps
    for a, b in ((q0, q1), (q1, q2), (p0, p1), (p1, p2), (w0, w1), (w1, w2)):
        assert 1.5 < a / b < 2.5, errs                            # first order
    assert q2 < 1e-2 and p2 < 3e-3, errs


def _implicit_pd_case(robot, h=0.01, kp=85.0, kd=2.0):
    from tests.kernel_checks import default_pose

    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(dt=h, substeps=1, ground=False)
    rng = np.random.default_rng(7)
    nd = art.num_dofs
    root, dof = random_flying_state(art, 1, rng, scale_qd=0.5)
    root, dof = root[0].astype(np.float64), dof[0].astype(np.float64)
    root[3:7] /= np.linalg.norm(root[3:7])
    root[7:13] *= 0.5
    target = default_pose(art) + rng.uniform(-0.2, 0.2, nd)
    L = Lagrange(art, m, [sp.gravity[0], sp.gravity[1], sp.gravity[2]])
    nb = 0 if art.fixed_base else 6
    q, qd = dof[:, 0], dof[:, 1]
    # the drive law linearised about the end-of-step state: tau(q+, v+) = kp (target - q - h v+) - kd v+
    #   => (M + h kd + h^2 kp) (v+ - v) = h [ kp (target - q - h v) - kd v - bias(q, v) ]
    xdd, _, M = L.accelerations(root, dof, kp * (target - q - h * qd) - kd * qd)
    D = np.zeros(nb + nd)
    D[nb:] = h * kd + h * h * kp
    xd = np.concatenate([root[7:10], root[10:13], qd]) if nb else qd.copy()
    v_new = xd + np.linalg.solve(M + np.diag(D), h * (M @ xdd))
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, kp, kd)
    return art, m, sp, props, root, dof, target, v_new, nb


@pytest.mark.parametrize("robot", ["hound", "anymal", "houndarm"])
def test_implicit_pd_drive_is_the_linearly_implicit_euler_step(robot):
    """Position drives (the flat tasks, cfg/task/Anymal.yaml:55-56) are integrated implicitly: one sub-step of the oracle equals the linearly
    implicit Euler step of the Euler-Lagrange model with the drive's (diagonal) stiffness and damping on the left-hand side -- to rounding,
    not to O(h)."""
    art, m, sp, props, root, dof, target, v_new, nb = _implicit_pd_case(robot)
    h = sp.dt
    r, d = root[None].copy(), dof[None].copy()
    O.simulate(m, sp, props, r, d, target[None], np.zeros((1, art.num_dofs)))
    got = np.concatenate([r[0][7:10], r[0][10:13], d[0][:, 1]]) if nb else d[0][:, 1]
    tol = 2e-7 if robot in EXACT_FRAMES else 2e-6
    assert np.abs(got - v_new).max() < tol * max(1.0, np.abs(v_new).max())
    assert np.abs(d[0][:, 0] - (dof[:, 0] + h * v_new[nb:])).max() < tol


@pytest.mark.parametrize("robot", ["hound", "anymal", "houndarm"])
def test_kernel_code_implicit_pd_step(robot):
    """The same step through the step kernels' code (float32, host lane emulator)."""
    from tests.backends import EmuBackend

    art, m, sp, props, root, dof, target, v_new, nb = _implicit_pd_case(robot)
    be = EmuBackend(art, sp, props, 1)
    try:
        be.set_state(root[None].astype(np.float32), dof[None].astype(np.float32))
        be.simulate(target[None].astype(np.float32), np.zeros((1, art.num_dofs), np.float32))
        r, d = be.get_state()
    finally:
        be.close()
    got = (np.concatenate([r[0][7:10], r[0][10:13], d[0][:, 1]]) if nb else d[0][:, 1]).astype(np.float64)
    assert np.abs(got - v_new).max() < 1e-4 * max(1.0, np.abs(v_new).max()), np.abs(got - v_new).max()
