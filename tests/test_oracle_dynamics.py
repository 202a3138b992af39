"""The dynamics oracle validates itself by invariants (the reference offers no golden vector at this boundary --
Isaac Gym/PhysX is a closed, absent binary: PARITY UNPINNED against PhysX; the contact-free dynamics are pinned against an
independent Euler-Lagrange derivation in tests/test_oracle_lagrange.py):
  * ABA forward dynamics == solve(CRBA mass matrix, tau - RNEA bias)          (two independent code paths)
  * kinetic energy from the spatial inertias == finite-difference FK energy   (independent numpy kinematics)
  * free flight: linear momentum (x, y), angular momentum (z) and energy conservation
  * a robot dropped on the plane settles with the contact forces summing to m g
  * float32 instantiation tracks float64
"""
import numpy as np
import pytest

from isaacgymenv_b200 import _abi
from isaacgymenv_b200.model import urdf
from oracle import dyn_oracle as O
from tests.kernel_checks import default_pose, flat_params, load_robot, random_flying_state


@pytest.mark.parametrize("robot", ["anymal", "hound", "useful_hound", "cartpole"])
def test_aba_equals_crba_rnea(robot):
    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    rng = np.random.default_rng(0)
    root, dof = random_flying_state(art, 4, rng)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    tau = rng.normal(size=(4, art.num_dofs)) * 10
    qdd, a0 = O.forward_dynamics(m, sp, root, dof, tau)
    for e in range(4):
        H, Cb = O.crba_rnea(m, sp, root[e], dof[e])
        assert np.abs(H - H.T).max() < 1e-12
        assert np.linalg.eigvalsh(H).min() > 0
        nb = 0 if art.fixed_base else 6
        x = np.linalg.solve(H, np.concatenate([np.zeros(nb), tau[e]]) - Cb)
        np.testing.assert_allclose(x[nb:], qdd[e], rtol=1e-8, atol=1e-8)
        if nb:
            np.testing.assert_allclose(x[:6], a0[e], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("robot", ["anymal", "useful_hound"])
def test_kinetic_energy_against_finite_difference_kinematics(robot):
    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    rng = np.random.default_rng(1)
    root, dof = random_flying_state(art, 1, rng)
    root, dof = root[0].astype(np.float64), dof[0].astype(np.float64)
    ke, pe, mom = O.energy_momentum(m, sp, root, dof)

    def advance(eps):
        r = root.copy()
        r[:3] += eps * root[7:10]
        w = root[10:13]
        ang = np.linalg.norm(w) * eps
        ax = w / np.linalg.norm(w)
        x1, y1, z1 = ax * np.sin(ang / 2)
        w1 = np.cos(ang / 2)
        x2, y2, z2, w2 = root[3:7]
        r[3:7] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                  w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
        return urdf.forward_kinematics(art, dof[:, 0] + eps * dof[:, 1], r[:3], r[3:7])

    eps = 1e-6
    (pp, rp), (pm, rm), (p0, r0) = advance(eps), advance(-eps), advance(0.0)
    T = 0.0
    m32 = np.array(list(m.link_mass)[:art.num_links], dtype=np.float64)     # the oracle sees float32-rounded parameters
    for l in range(art.num_links):
        com = np.array(m.link_com[l][:], dtype=np.float64)
        i6 = np.array(m.link_inertia[l][:], dtype=np.float64)
        il = np.array([[i6[0], i6[3], i6[4]], [i6[3], i6[1], i6[5]], [i6[4], i6[5], i6[2]]])
        v = ((pp[l] + rp[l] @ com) - (pm[l] + rm[l] @ com)) / (2 * eps)
        W = ((rp[l] - rm[l]) / (2 * eps)) @ r0[l].T
        w = np.array([W[2, 1], W[0, 2], W[1, 0]])
        T += 0.5 * m32[l] * v @ v + 0.5 * w @ (r0[l] @ il @ r0[l].T) @ w
    assert abs(T - ke) / ke < 1e-6


def test_free_flight_conservation():
    art = load_robot("anymal")
    m = _abi.pack_model(art)
    sp = _abi.SimParams(dt=0.0005, substeps=1, num_position_iterations=0, num_velocity_iterations=0, joint_limit_stiffness=2000.0, joint_limit_damping=20.0)
    sp.gravity[2] = -9.81
    props = _abi.default_dof_props(art)
    for d in range(art.num_dofs):
        props.lower[d], props.upper[d], props.velocity[d] = -3e38, 3e38, 0
    rng = np.random.default_rng(2)
    root, dof = random_flying_state(art, 1, rng)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    ke0, pe0, mom0 = O.energy_momentum(m, sp, root[0], dof[0])
    z = np.zeros((1, art.num_dofs))
    steps = 1000
    for _ in range(steps):
        O.simulate(m, sp, props, root, dof, z, z)
    ke1, pe1, mom1 = O.energy_momentum(m, sp, root[0], dof[0])
    t = steps * sp.dt
    assert abs((ke1 + pe1) - (ke0 + pe0)) / abs(ke0 + pe0) < 2e-3          # first-order integrator drift
    np.testing.assert_allclose(mom1[3:5], mom0[3:5], rtol=1e-3, atol=1e-2)    # linear x, y
    np.testing.assert_allclose(mom1[5] - mom0[5], -art.total_mass * 9.81 * t, rtol=1e-3)
    np.testing.assert_allclose(mom1[2], mom0[2], rtol=1e-3, atol=1e-2)        # angular z (gravity has no z moment)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_drop_and_stand(dtype):
    art = load_robot("anymal")
    m = _abi.pack_model(art)
    sp = flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    q0 = default_pose(art)
    root = np.zeros((1, 13), dtype)
    root[0, 2], root[0, 6] = 0.62, 1
    dof = np.zeros((1, 12, 2), dtype)
    dof[0, :, 0] = q0
    z = np.zeros((1, 12), dtype)
    for _ in range(150):
        f, c = O.simulate(m, sp, props, root, dof, q0[None].astype(dtype), z)
    mg = float(sum(list(m.link_mass)[:13])) * 9.81
    assert abs(c[0, :, 2].sum() - mg) / mg < 2e-3
    feet = [i for i, n in enumerate(art.body_names) if "SHANK" in n]
    assert (c[0, feet, 2] > 100).all() and np.abs(np.delete(c[0], feet, axis=0)).max() == 0
    assert 0.45 < root[0, 2] < 0.56 and abs(root[0, 9]) < 1e-2 and np.abs(root[0, 3:6]).max() < 0.02
    assert np.abs(f).max() < 80.0
    assert np.abs(dof[0, :, 1]).max() < 0.05


def test_joint_limits_hold():
    art = load_robot("anymal")
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    root = np.zeros((1, 13))
    root[0, 2], root[0, 6] = 3.0, 1
    dof = np.zeros((1, 12, 2))
    dof[0, :, 0] = default_pose(art)
    act = np.zeros((1, 12))
    act[0, 0] = 60.0       # push LF_HAA into its upper limit (0.49 rad)
    for _ in range(40):
        O.simulate(m, sp, props, root, dof, np.zeros((1, 12)), act)
    assert dof[0, 0, 0] < 0.49 + 0.05
    assert dof[0, 0, 0] > 0.40
