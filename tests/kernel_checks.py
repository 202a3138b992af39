"""Parity checks shared by the CPU (host lane emulator) and GPU (libb200gym.so through its C ABI) test
modules.  Every check compares the kernel code against the oracle (oracle/) or the committed golden vectors
(tests/golden/, outputs of the reference's own @torch.jit.script functions).

Tolerances (north_star): reward/obs math 1e-5 relative on identical state tensors, integer/boolean masks
bit-exact; contact-free joint accelerations 1e-3 relative; contact states within 1e-2 over a 10-step horizon.
"""
from __future__ import annotations

import os

import numpy as np

from isaacgymenv_b200 import _abi
from isaacgymenv_b200.model.store import COMPILED_DIR, load_articulation
from oracle import dyn_oracle as O
from oracle import task_math as tm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ANYMAL_DEFAULT = {"LF_HAA": .03, "LH_HAA": .03, "RF_HAA": -.03, "RH_HAA": -.03, "LF_HFE": .4, "LH_HFE": -.4, "RF_HFE": .4, "RH_HFE": -.4,
                  "LF_KFE": -.8, "LH_KFE": .8, "RF_KFE": -.8, "RH_KFE": .8}
HOUND_DEFAULT = {"roll": 0.0, "hip": 0.7854, "knee": -1.5708}


def load_robot(name):
    files = {"anymal": "urdf__anymal_c__urdf__anymal.c1k1f0.json", "anymal_minimal": "urdf__anymal_c__urdf__anymal_minimal.c1k1f0.json",
             "hound": "urdf__Hound_new__Hound.c0k0f0.json", "useful_hound": "urdf__UsefulHound__urdf__Hound.c0k0f0.json",
             "cartpole": "urdf__cartpole.c0k0f1.json", "houndarm": "urdf__open_manipulator_p_gazebo__urdf__open_manipulator_p.c0k0f1.json",
             "manipulator": "urdf__franka_description__robots__franka_panda_manipulator.c0k0f1.json"}
    return load_articulation(os.path.join(COMPILED_DIR, files[name]))


def default_pose(art):
    if art.dof_names[0] in ANYMAL_DEFAULT:
        return np.array([ANYMAL_DEFAULT[n] for n in art.dof_names])
    out = []
    for n in art.dof_names:
        out.append(next((v for k, v in HOUND_DEFAULT.items() if k in n), 0.0))
    return np.array(out)


def flat_params(dt=0.02, substeps=2, npos=4, nvel=1, ground=True, slots=0):
    sp = _abi.SimParams(max_contacts_per_chain=slots, dt=dt, substeps=substeps, num_position_iterations=npos, num_velocity_iterations=nvel, contact_offset=0.02,
                        rest_offset=0.0, bounce_threshold_velocity=0.2, max_depenetration_velocity=100.0, plane_static_friction=1.0,
                        plane_dynamic_friction=1.0, plane_restitution=0.0, has_ground=1 if ground else 0, joint_limit_stiffness=2000.0,
                        joint_limit_damping=20.0)
    sp.gravity[2] = -9.81
    return sp


def random_flying_state(art, n, rng, scale_qd=2.0):
    nd = art.num_dofs
    root = np.zeros((n, 13), np.float32)
    root[:, 2] = 5.0
    q = rng.normal(size=(n, 4))
    root[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    root[:, 7:13] = rng.normal(size=(n, 6))
    dof = np.zeros((n, nd, 2), np.float32)
    dof[:, :, 0] = default_pose(art) + rng.uniform(-.3, .3, (n, nd))
    dof[:, :, 1] = rng.normal(size=(n, nd)) * scale_qd
    return root, dof


def standing_state(art, n, rng, z0=0.55):
    nd = art.num_dofs
    root = np.zeros((n, 13), np.float32)
    root[:, 2] = z0 + rng.uniform(0, .1, n)
    root[:, 3:6] = rng.normal(size=(n, 3)) * 0.05
    root[:, 6] = 1
    root[:, 3:7] /= np.linalg.norm(root[:, 3:7], axis=1, keepdims=True)
    dof = np.zeros((n, nd, 2), np.float32)
    dof[:, :, 0] = default_pose(art) * rng.uniform(.5, 1.5, (n, nd))
    dof[:, :, 1] = rng.uniform(-.1, .1, (n, nd))
    return root, dof


# ------------------------------------------------------------------------------------------------
def check_forward_dynamics(make_backend, robot="anymal", n=16, seed=0):
    """Contact-free joint and root accelerations vs the float64 oracle: <= 1e-3 relative (north_star)."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    sp = flat_params(ground=False)
    props = _abi.default_dof_props(art)
    root, dof = random_flying_state(art, n, rng)
    tau = (rng.normal(size=(n, art.num_dofs)) * 20).astype(np.float32)
    be = make_backend(art, sp, props, n)
    try:
        be.set_state(root, dof)
        qdd, a0 = be.forward_dynamics(tau)
    finally:
        be.close()
    m = _abi.pack_model(art)
    qo, ao = O.forward_dynamics(m, sp, root.astype(np.float64), dof.astype(np.float64), tau.astype(np.float64))
    if art.fixed_base:
        ao = np.zeros_like(ao)
    err_q = np.abs(qdd - qo).max(axis=1) / np.abs(qo).max(axis=1)
    assert err_q.max() < 1e-3, f"joint acceleration relative error {err_q.max():.2e}"
    if not art.fixed_base:
        err_a = np.abs(a0 - ao).max(axis=1) / np.abs(ao).max(axis=1)
        assert err_a.max() < 1e-3, f"root acceleration relative error {err_a.max():.2e}"
    return float(err_q.max())


def scaled_model_props(m, props, mass, kp, kd):
    """Copies of the packed model / DOF properties with every link mass + inertia scaled by ``mass`` and the drive gains by
    ``kp`` / ``kd``: what one environment of a domain-randomised sim simulates."""
    m2 = _abi.Model.from_buffer_copy(m)
    p2 = _abi.DofProps.from_buffer_copy(props)
    for i in range(_abi.MAX_LINKS):
        m2.link_mass[i] = m.link_mass[i] * mass
        for k in range(6):
            m2.link_inertia[i][k] = m.link_inertia[i][k] * mass
    for i in range(_abi.MAX_DOF):
        p2.stiffness[i] = props.stiffness[i] * kp
        p2.damping[i] = props.damping[i] * kd
    return m2, p2


def link_scaled_model_props(m, props, link_scale):
    """Same with one [mass, stiffness, damping] triple per link (row l of ``link_scale``: link 0 = root, 1 + d = child of DOF d)."""
    m2 = _abi.Model.from_buffer_copy(m)
    p2 = _abi.DofProps.from_buffer_copy(props)
    for l in range(link_scale.shape[0]):
        m2.link_mass[l] = m.link_mass[l] * float(link_scale[l, 0])
        for k in range(6):
            m2.link_inertia[l][k] = m.link_inertia[l][k] * float(link_scale[l, 0])
        if l > 0:
            p2.stiffness[l - 1] = props.stiffness[l - 1] * float(link_scale[l, 1])
            p2.damping[l - 1] = props.damping[l - 1] * float(link_scale[l, 2])
            p2.lower[l - 1] = props.lower[l - 1] + float(link_scale[l, 3])
            p2.upper[l - 1] = props.upper[l - 1] + float(link_scale[l, 4])
    return m2, p2


def check_link_scale(make_backend, robot="anymal", n=6, steps=8, seed=23):
    """Per-link / per-DOF domain randomisation (B2G_T_LINK_SCALE: the reference draws every body mass and every DOF stiffness / damping
    on its own, utils/dr_utils.py:135-238): each environment against the float64 oracle on its individually scaled model."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    sp = flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    ls = np.zeros((n, nd + 1, _abi.LINK_SCALE_COLS), np.float32)
    ls[:, :, :3] = rng.uniform(0.6, 1.4, (n, nd + 1, 3))
    ls[:, :, 3] = rng.uniform(0.0, 0.3, (n, nd + 1))       # lower limits move up, upper limits down: the stops really bite
    ls[:, :, 4] = -rng.uniform(0.0, 0.3, (n, nd + 1))
    ls[0, :, :3], ls[0, :, 3:] = 1.0, 0.0
    m = _abi.pack_model(art)
    be = make_backend(art, flat_params(ground=False), _abi.default_dof_props(art), n)
    try:
        fr, fd = random_flying_state(art, n, rng)
        tau = (rng.normal(size=(n, nd)) * 20).astype(np.float32)
        be.set_state(fr, fd)
        be.set_link_scale(ls)
        qdd, _ = be.forward_dynamics(tau)
    finally:
        be.close()
    for e in range(n):
        m2, _ = link_scaled_model_props(m, props, ls[e])
        qo, _ = O.forward_dynamics(m2, flat_params(ground=False), fr[e:e + 1].astype(np.float64), fd[e:e + 1].astype(np.float64), tau[e:e + 1].astype(np.float64))
        err = np.abs(qdd[e] - qo[0]).max() / np.abs(qo[0]).max()
        assert err < 1e-3, f"env {e}: joint acceleration relative error {err:.2e} with per-link mass scales"
    root, dof = standing_state(art, n, rng, 0.55 if "anymal" in robot else 0.5)
    be = make_backend(art, sp, props, n)
    r64, d64 = root.astype(np.float64), dof.astype(np.float64)
    q0 = default_pose(art)
    worst = 0.0
    try:
        be.set_state(root, dof)
        be.set_link_scale(ls)
        for _ in range(steps):
            tgt = q0 + 0.5 * rng.uniform(-1, 1, (n, nd))
            act = np.zeros((n, nd))
            f_dev, _ = be.simulate(tgt, act)
            for e in range(n):
                m2, p2 = link_scaled_model_props(m, props, ls[e])
                re, de = r64[e:e + 1].copy(), d64[e:e + 1].copy()
                f_o, _ = O.simulate(m2, sp, p2, re, de, tgt[e:e + 1].astype(np.float64), act[e:e + 1].astype(np.float64))
                r64[e], d64[e] = re[0], de[0]
            rb, db = be.get_state()
            worst = max(worst, float(np.abs(rb - r64).max()), float(np.abs(db[:, :, 0] - d64[:, :, 0]).max()))
    finally:
        be.close()
    assert worst < 1e-2, f"state deviation {worst:.2e} over {steps} steps with per-link scales"
    return worst


def check_env_scale(make_backend, robot="anymal", n=8, steps=10, seed=17):
    """Tensorised domain randomisation (B2G_T_ENV_SCALE + B2G_T_FRICTION): every environment simulates its own mass / drive
    gain / friction scales; each one is checked against the float64 oracle run on the correspondingly scaled model."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    sp = flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    scale = np.ones((n, 4), np.float32)
    scale[:, 0] = rng.uniform(0.5, 1.5, n)
    scale[:, 1] = rng.uniform(0.5, 1.5, n)
    scale[:, 2] = rng.uniform(0.5, 1.5, n)
    scale[0, :3] = 1.0
    fric = rng.uniform(0.4, 1.3, n).astype(np.float32)
    root, dof = standing_state(art, n, rng, 0.55 if "anymal" in robot else 0.5)
    m = _abi.pack_model(art)
    # contact-free accelerations first (mass scale only matters)
    be = make_backend(art, flat_params(ground=False), _abi.default_dof_props(art), n)
    try:
        fr, fd = random_flying_state(art, n, rng)
        tau = (rng.normal(size=(n, nd)) * 20).astype(np.float32)
        be.set_state(fr, fd)
        be.set_env_scale(scale)
        qdd, _ = be.forward_dynamics(tau)
    finally:
        be.close()
    for e in range(n):
        m2, _ = scaled_model_props(m, props, float(scale[e, 0]), 1.0, 1.0)
        qo, _ = O.forward_dynamics(m2, flat_params(ground=False), fr[e:e + 1].astype(np.float64), fd[e:e + 1].astype(np.float64), tau[e:e + 1].astype(np.float64))
        err = np.abs(qdd[e] - qo[0]).max() / np.abs(qo[0]).max()
        assert err < 1e-3, f"env {e}: joint acceleration relative error {err:.2e} with mass scale {scale[e, 0]:.2f}"
    assert np.abs(qdd[1] - qdd[0]).max() > 1e-3 or abs(scale[1, 0] - 1) < 1e-3
    # then a contact horizon with every scale + friction active
    be = make_backend(art, sp, props, n)
    r64, d64 = root.astype(np.float64), dof.astype(np.float64)
    q0 = default_pose(art)
    worst = 0.0
    try:
        be.set_state(root, dof)
        be.set_env_scale(scale, fric)
        for _ in range(steps):
            tgt = q0 + 0.5 * rng.uniform(-1, 1, (n, nd))
            act = np.zeros((n, nd))
            be.simulate(tgt, act)
            for e in range(n):
                m2, p2 = scaled_model_props(m, props, *[float(x) for x in scale[e, :3]])
                re, de = r64[e:e + 1].copy(), d64[e:e + 1].copy()
                O.simulate(m2, sp, p2, re, de, tgt[e:e + 1].astype(np.float64), act[e:e + 1].astype(np.float64), friction=fric[e:e + 1].astype(np.float64))
                r64[e], d64[e] = re[0], de[0]
            rb, db = be.get_state()
            worst = max(worst, float(np.abs(rb - r64).max()), float(np.abs(db[:, :, 0] - d64[:, :, 0]).max()))
    finally:
        be.close()
    assert worst < 1e-2, f"state deviation {worst:.2e} over {steps} steps with per-env scales"
    return worst


def check_simulate_horizon(make_backend, robot="anymal", n=16, steps=10, seed=1, drive="pos", tol=1e-2):
    """gym.simulate with ground contact for `steps` steps vs the float64 oracle: states within 1e-2."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    if drive == "pos":
        sp = flat_params()
        props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    else:
        sp = flat_params(dt=0.005, substeps=1)
        props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    z0 = 0.55 if "anymal" in robot else 0.5
    root, dof = standing_state(art, n, rng, z0)
    be = make_backend(art, sp, props, n)
    m = _abi.pack_model(art)
    r64, d64 = root.astype(np.float64), dof.astype(np.float64)
    q0 = default_pose(art)
    worst = 0.0
    saw_contact = False
    try:
        be.set_state(root, dof)
        for _ in range(steps):
            if drive == "pos":
                tgt = q0 + 0.5 * rng.uniform(-1, 1, (n, nd))
                act = np.zeros((n, nd))
            else:
                rb, db = be.get_state()
                tgt = np.zeros((n, nd))
                act = np.clip(80.0 * (q0 + 0.5 * rng.uniform(-1, 1, (n, nd)) - db[:, :, 0]) - 2.0 * db[:, :, 1], -80, 80)
            f, c = be.simulate(tgt, act)
            f64, c64 = O.simulate(m, sp, props, r64, d64, tgt.astype(np.float64), act.astype(np.float64))
            rb, db = be.get_state()
            worst = max(worst, float(np.abs(rb - r64).max()), float(np.abs(db[:, :, 0] - d64[:, :, 0]).max()))
            saw_contact = saw_contact or bool((np.abs(c64) > 1.0).any())
            cn = np.abs(c - c64).max() / max(1.0, np.abs(c64).max())
            assert cn < 5e-2, f"contact force relative deviation {cn:.2e}"
            assert np.abs(db[:, :, 1] - d64[:, :, 1]).max() < 10 * tol * max(1.0, np.abs(d64[:, :, 1]).max())
    finally:
        be.close()
    assert saw_contact, "test never reached ground contact"
    assert worst < tol, f"state deviation {worst:.2e} over {steps} steps"
    return worst


def check_root_velocity_limits(make_backend, robot="anymal", n=8, seed=4):
    """b2g_sim_params::max_linear_velocity / max_angular_velocity (AssetOptions defaults 1000 m/s, 64 rad/s): a robot in free flight that
    spins at 150 rad/s and moves at 2500 m/s leaves the step at exactly the limits; one inside them is untouched; limits of 0 switch the
    clamp off.  Kernel == oracle in all three cases."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    m = _abi.pack_model(art)
    q0 = default_pose(art)
    for lim_v, lim_w in ((1000.0, 64.0), (0.0, 0.0)):
        sp = flat_params(ground=False)
        sp.gravity[2] = 0.0
        sp.max_linear_velocity, sp.max_angular_velocity = lim_v, lim_w
        root, dof = standing_state(art, n, rng, 5.0)
        dof[:, :, 1] = 0.0
        fast = np.arange(n) % 2 == 0
        root[fast, 7:10] = np.array([2500.0, 0.0, 0.0], np.float32)
        root[fast, 10:13] = np.array([0.0, 0.0, 150.0], np.float32)
        root[~fast, 7:10] = np.array([1.0, -2.0, 0.5], np.float32)
        root[~fast, 10:13] = np.array([0.3, 0.2, -0.4], np.float32)
        be = make_backend(art, sp, props, n)
        try:
            be.set_state(root, dof)
            r64, d64 = root.astype(np.float64), dof.astype(np.float64)
            tgt = np.tile(q0, (n, 1))
            be.simulate(tgt, np.zeros((n, nd)))
            O.simulate(m, sp, props, r64, d64, tgt.astype(np.float64), np.zeros((n, nd)))
            rb, _ = be.get_state()
        finally:
            be.close()
        speed, spin = np.linalg.norm(rb[:, 7:10], axis=1), np.linalg.norm(rb[:, 10:13], axis=1)
        if lim_v > 0:
            assert np.all(speed[fast] <= lim_v * (1 + 1e-5)) and np.all(speed[fast] > 0.99 * lim_v), speed
            assert np.all(spin[fast] <= lim_w * (1 + 1e-5)) and np.all(spin[fast] > 0.9 * lim_w), spin      # (momentum moves between the base and the legs inside the step)
        else:
            assert np.all(speed[fast] > 2000.0) and np.all(spin[fast] > 100.0)
        assert np.all(speed[~fast] < 3.0) and np.all(spin[~fast] < 1.0)
        np.testing.assert_allclose(rb[:, 7:13], r64[:, 7:13], rtol=2e-3, atol=2e-3)
        np.testing.assert_allclose(rb[~fast, :3], r64[~fast, :3], atol=1e-4)


def check_drive_saturation(make_backend, robot="anymal", n=6, seed=12):
    """An implicit position drive whose torque would exceed the DOF's effort limit acts as a constant torque of that size (PhysX clamps the
    drive force to maxForce): with every target 2 rad away (85 N m / rad x 2 rad against 80 N m) a step in position mode equals the step
    of an effort-mode twin fed +-effort, on the kernel and on the oracle; with targets 0.2 rad away nothing saturates and the two differ."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    m = _abi.pack_model(art)
    sp = flat_params(ground=False)
    sp.gravity[2] = 0.0
    effort = np.array(art.effort, np.float64)
    assert (effort > 0).all() and (effort < 85.0 * 1.5).all()
    root, dof = standing_state(art, n, rng, 3.0)
    sign = np.where(rng.uniform(size=(n, nd)) < 0.5, -1.0, 1.0)
    out = {}
    for name, mode, tgt, act in (("pos_far", _abi.DOF_MODE_POS, dof[:, :, 0] + 2.0 * sign, np.zeros((n, nd))),
                                 ("effort", _abi.DOF_MODE_EFFORT, np.zeros((n, nd)), sign * effort),
                                 ("pos_near", _abi.DOF_MODE_POS, dof[:, :, 0] + 0.2 * sign, np.zeros((n, nd)))):
        props = _abi.default_dof_props(art, mode, 85.0 if mode == _abi.DOF_MODE_POS else 0.0, 2.0 if mode == _abi.DOF_MODE_POS else 0.0)
        be = make_backend(art, sp, props, n)
        try:
            be.set_state(root.copy(), dof.copy())
            r64, d64 = root.astype(np.float64), dof.astype(np.float64)
            f, _ = be.simulate(tgt, act)
            f64, _ = O.simulate(m, sp, props, r64, d64, tgt.astype(np.float64), act.astype(np.float64))
            rb, db = be.get_state()
        finally:
            be.close()
        np.testing.assert_allclose(db[:, :, 0], d64[:, :, 0], atol=2e-4, err_msg=name)
        np.testing.assert_allclose(db[:, :, 1], d64[:, :, 1], rtol=2e-3, atol=2e-3, err_msg=name)
        out[name] = (rb, db, f)
    np.testing.assert_allclose(out["pos_far"][1], out["effort"][1], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out["pos_far"][0], out["effort"][0], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.abs(out["pos_far"][2]), np.tile(effort, (n, 1)), rtol=1e-6)      # the reported DOF force sits on the limit
    assert np.abs(out["pos_near"][1][:, :, 1] - out["effort"][1][:, :, 1]).max() > 1.0


def root_box(art):
    """The base body's bounding box (centre, half extents; root frame) around its own contact candidates and their radii."""
    pts = np.array([p for l, b, p in zip(art.cp_link, art.cp_body, art.cp_pos) if l == 0 and b == 0], np.float64)
    rad = np.array([r for l, b, r in zip(art.cp_link, art.cp_body, art.cp_radius) if l == 0 and b == 0], np.float64)
    lo, hi = (pts - rad[:, None]).min(0), (pts + rad[:, None]).max(0)
    return 0.5 * (lo + hi), 0.5 * (hi - lo)


def self_penetration(art, root, dof):
    """Deepest penetration (m, >= 0) of any candidate sphere of a link that does not hang off the root directly into the root's box."""
    from isaacgymenv_b200.model.urdf import forward_kinematics as link_poses

    bc, bh = root_box(art)
    first = set(int(s) + 1 for s in art.chain_start)      # links attached to the root
    worst = np.zeros(root.shape[0])
    for e in range(root.shape[0]):
        pos, rot = link_poses(art, dof[e, :, 0].astype(np.float64), np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0]))      # in the root frame
        for l, cp, r in zip(art.cp_link, art.cp_pos, art.cp_radius):
            if l == 0 or int(l) in first:
                continue
            q = pos[l] + rot[l] @ cp - bc
            gap = (np.abs(q) - bh - r).max()
            worst[e] = max(worst[e], -gap)
    return worst


def check_self_collision(make_backend, robot="useful_hound", n=8, steps=30, seed=9):
    """b2g_sim_params::self_collision: robots in free flight (no ground, no gravity) whose joints are driven towards targets that fold the legs /
    the arm INTO the base.  (a) kernel == float64 oracle along the rollout; (b) with the flag the links stop at the base's box (a few
    millimetres of solver slack), without it they pass through by centimetres; (c) the contact is an internal force pair: the force
    reported on the root body is the opposite of the sum over the links."""
    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 60.0, 3.0)
    m = _abi.pack_model(art)
    q0 = default_pose(art)
    lower, upper = np.array(art.lower, np.float64), np.array(art.upper, np.float64)
    tgt = np.clip(q0 + rng.uniform(-1.6, 1.6, (n, nd)), np.where(np.isfinite(lower), lower, -3.0), np.where(np.isfinite(upper), upper, 3.0))
    pen = {}
    for flag in (1, 0):
        sp = flat_params(dt=0.01, substeps=2, ground=False)
        sp.gravity[2] = 0.0
        sp.self_collision = flag
        root = np.zeros((n, 13), np.float32)
        root[:, 2], root[:, 6] = 3.0, 1.0
        dof = np.zeros((n, nd, 2), np.float32)
        dof[:, :, 0] = q0
        be = make_backend(art, sp, props, n)
        try:
            be.set_state(root, dof)
            r64, d64 = root.astype(np.float64), dof.astype(np.float64)
            saw, worst_dev = False, 0.0
            for k in range(steps):
                f, c = be.simulate(tgt, np.zeros((n, nd)))
                f64, c64 = O.simulate(m, sp, props, r64, d64, tgt.astype(np.float64), np.zeros((n, nd)))
                rb, db = be.get_state()
                worst_dev = max(worst_dev, float(np.abs(rb - r64).max()), float(np.abs(db[:, :, 0] - d64[:, :, 0]).max()))
                c3, c643 = c.reshape(n, -1, 3), c64.reshape(n, -1, 3)
                if flag:
                    saw = saw or bool((np.abs(c643) > 1.0).any())
                    np.testing.assert_allclose(c3.sum(1), 0.0, atol=2e-2 * max(1.0, np.abs(c3).max()))        # (c): action = -reaction
                    assert np.abs(c3 - c643).max() < 5e-2 * max(1.0, np.abs(c643).max()) + 0.5
                else:
                    assert np.abs(c3).max() == 0.0
                r64, d64 = rb.astype(np.float64), db.astype(np.float64)      # stay on the kernel's trajectory
            assert worst_dev < 2e-3, f"self_collision={flag}: kernel deviates from the oracle by {worst_dev:.2e} within a step"
            if flag:
                assert saw, "the folded pose never produced a self contact"
            pen[flag] = self_penetration(art, r64, d64)
        finally:
            be.close()
    assert pen[0].max() > 0.02, f"without the flag the test pose does not fold into the base ({pen[0].max():.3f} m)"
    assert pen[1].max() < 0.012, f"links penetrate the base by {pen[1].max():.3f} m with self-collision on"
    return pen


# ------------------------------------------------------------------------------------------------
def anymal_cfg(art, seed=42, robot="anymal"):
    c = _abi.AnymalCfg()
    c.lin_vel_scale, c.ang_vel_scale, c.dof_pos_scale, c.dof_vel_scale, c.action_scale = 2.0, 0.25, 1.0, 0.05, 0.5
    c.rew_lin_vel_xy, c.rew_ang_vel_z, c.rew_torque = 1.0 * 0.02, 0.5 * 0.02, -0.000025 * 0.02
    c.clip_obs, c.clip_actions = 5.0, 1.0
    c.cmd_x[0], c.cmd_x[1], c.cmd_y[0], c.cmd_y[1], c.cmd_yaw[0], c.cmd_yaw[1] = -2, 2, -1, 1, -1, 1
    for i, v in enumerate(default_pose(art)):
        c.default_dof_pos[i] = v
    for i, v in enumerate([0, 0, 0.62, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]):
        c.init_root[i] = v
    key = "THIGH" if robot == "anymal" else "thigh"
    knees = [i for i, n in enumerate(art.body_names) if key in n]
    c.base_body, c.n_knee = 0, len(knees)
    for i, k in enumerate(knees):
        c.knee_bodies[i] = k
    c.max_episode_length = 2500
    c.seed = seed
    return c


def cfg_dict(c, nd):
    return dict(default_dof_pos=np.array(list(c.default_dof_pos)[:nd], np.float32), init_root=np.array(list(c.init_root), np.float32),
                cmd_x=list(c.cmd_x), cmd_y=list(c.cmd_y), cmd_yaw=list(c.cmd_yaw), lin_vel_scale=c.lin_vel_scale,
                ang_vel_scale=c.ang_vel_scale, dof_pos_scale=c.dof_pos_scale, dof_vel_scale=c.dof_vel_scale,
                rew_scales={"lin_vel_xy": c.rew_lin_vel_xy, "ang_vel_z": c.rew_ang_vel_z, "torque": c.rew_torque},
                knee_bodies=np.array(list(c.knee_bodies)[:c.n_knee]), base_body=c.base_body, max_episode_length=int(c.max_episode_length),
                clip_obs=c.clip_obs)


def check_post_physics_golden(make_backend, robot="anymal"):
    """The kernel's post_physics_step (obs + reward + reset mask) on the golden inputs vs the reference's own
    outputs: 1e-5 relative, masks bit-exact."""
    g = np.load(os.path.join(GOLDEN, "anymal_flat.npz" if robot == "anymal" else "hound_flat.npz"))
    art = load_robot(robot)
    n, nd = g["root"].shape[0], art.num_dofs
    sp = flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    c = anymal_cfg(art, robot=robot)
    c.clip_obs = 3.0e38
    for i, v in enumerate(g["default"][0]):
        c.default_dof_pos[i] = float(v)
    c.base_body = int(g["base"])
    for i, k in enumerate(g["knee"]):
        c.knee_bodies[i] = int(k)
    c.n_knee = len(g["knee"])
    c.rew_lin_vel_xy, c.rew_ang_vel_z, c.rew_torque = float(g["scale_lin"]), float(g["scale_ang"]), float(g["scale_torque"])
    c.max_episode_length = int(g["max_len"])
    be = make_backend(art, sp, props, n)
    try:
        be.anymal_create(c)
        dof = np.stack([g["dof_pos"], g["dof_vel"]], axis=2)
        be.set_state(g["root"], dof)
        # progress is incremented before use (tasks/anymal.py:232); reset_buf = 0 so no env is reset first
        be.set_task(commands=g["commands"], progress=g["progress"] - 1, reset=np.zeros(n, np.int64), dof_force=g["torques"], contact=g["contact"])
        be.anymal_post_only(g["actions"])
        out = be.get_task()
    finally:
        be.close()
    np.testing.assert_allclose(out["obs"], g["obs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["rew"], g["rew"], rtol=1e-5, atol=1e-8)
    assert np.array_equal(out["reset"], g["reset"].astype(np.int64)), "reset mask differs from the reference"
    want_timeout = ((g["progress"] >= int(g["max_len"]) - 1) & g["reset"]).astype(np.int64)
    assert np.array_equal(out["timeout"], want_timeout)
    assert np.array_equal(out["progress"], g["progress"])


def check_reset_draws(make_backend, robot="anymal", n=32, seed=1234):
    """reset_idx(all) with the in-kernel Philox stream == the numpy restatement of the stream, bit for bit."""
    art = load_robot(robot)
    nd = art.num_dofs
    sp = flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    c = anymal_cfg(art, seed=seed, robot=robot)
    be = make_backend(art, sp, props, n)
    try:
        be.anymal_create(c)
        be.anymal_reset_all()
        root, dof = be.get_state()
        t = be.get_task()
    finally:
        be.close()
    u = tm.philox_uniform(seed, np.arange(n), np.zeros(n), 2 * nd + 3)
    d0 = np.array(list(c.default_dof_pos)[:nd], np.float32)
    assert np.array_equal(dof[:, :, 0], d0[None] * tm.torch_rand_float(np.float32(0.5), np.float32(1.5), u[:, :nd]))
    assert np.array_equal(dof[:, :, 1], tm.torch_rand_float(np.float32(-0.1), np.float32(0.1), u[:, nd:2 * nd]))
    assert np.array_equal(t["commands"][:, 0], tm.torch_rand_float(np.float32(-2), np.float32(2), u[:, 2 * nd]))
    assert np.array_equal(t["commands"][:, 2], tm.torch_rand_float(np.float32(-1), np.float32(1), u[:, 2 * nd + 2]))
    assert np.array_equal(root, np.tile(np.array(list(c.init_root), np.float32), (n, 1)))
    assert (t["reset"] == 1).all() and (t["progress"] == 0).all()


def check_fused_step(make_backend, robot="anymal", n=16, steps=25, seed=3):
    """VecTask.step fused in one kernel vs the oracle composition (float32 dynamics oracle + numpy task math in the
    reference's order, same injected reset draws).  Physics drift is bounded loosely; the task math is then
    re-checked at 1e-5 on the kernel's own state tensors, masks bit-exact."""
    art = load_robot(robot)
    nd, nb = art.num_dofs, art.num_bodies
    rng = np.random.default_rng(seed)
    sp = flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    c = anymal_cfg(art, robot=robot)
    c.max_episode_length = 12      # make time-outs happen inside the test
    cd = cfg_dict(c, nd)
    m = _abi.pack_model(art)
    be = make_backend(art, sp, props, n)
    n_draws = 2 * nd + 3
    resets_seen = timeouts_seen = 0
    try:
        be.anymal_create(c)
        draws = rng.uniform(0, 1, (n, n_draws)).astype(np.float32)
        be.anymal_reset_all(draws)
        root, dof = be.get_state()
        t0 = be.get_task()
        st = dict(root=root.copy(), dof_pos=dof[:, :, 0].copy(), dof_vel=dof[:, :, 1].copy(), torques=np.zeros((n, nd), np.float32),
                  contact=np.zeros((n, nb, 3), np.float32), commands=t0["commands"].copy(), progress=t0["progress"].copy(), reset=t0["reset"].copy())
        for k in range(steps):
            actions = rng.uniform(-1.3, 1.3, (n, nd)).astype(np.float32)
            draws = rng.uniform(0, 1, (n, n_draws)).astype(np.float32)
            be.anymal_step(actions, draws)
            # oracle: pre_physics + simulate + post_physics
            a = np.clip(actions, -1.0, 1.0)
            tgt = (np.float32(0.5) * a + cd["default_dof_pos"][None]).astype(np.float32)
            dof_o = np.stack([st["dof_pos"], st["dof_vel"]], axis=2).astype(np.float32)
            f, cf = O.simulate(m, sp, props, st["root"], dof_o, tgt, np.zeros((n, nd), np.float32))
            st["dof_pos"], st["dof_vel"], st["torques"], st["contact"] = dof_o[:, :, 0].copy(), dof_o[:, :, 1].copy(), f, cf
            was_reset = st["reset"].copy()
            obs_o, obsc_o, rew_o, to_o = tm.anymal_post_physics(st, cd, a, draws)
            root_k, dof_k = be.get_state()
            tk = be.get_task()
            # (1) end-to-end agreement, physics tolerance: the two trajectories run free for up to 12 steps (an episode here); the bar is
            # north_star's 1e-2 over a 10-step horizon, stick / slip decisions of the friction cone are where float32 rounding shows
            assert np.abs(root_k - st["root"]).max() < 5e-3, f"step {k}: root deviates {np.abs(root_k - st['root']).max():.2e}"
            assert np.abs(dof_k[:, :, 0] - st["dof_pos"]).max() < 5e-3
            np.testing.assert_allclose(tk["obs"], obs_o, rtol=0, atol=2e-2)
            assert np.array_equal(tk["progress"], st["progress"])
            assert np.array_equal(tk["actions"], a)
            # envs that were reset this step carry exactly the injected draws
            ids = np.nonzero(was_reset)[0]
            if len(ids):
                assert np.array_equal(dof_k[ids, :, 0], st["dof_pos"][ids])
                assert np.array_equal(tk["commands"][ids], st["commands"][ids])
            # (2) task math at 1e-5 on the kernel's own state tensors
            grav = np.tile(np.array([[0, 0, -1]], np.float32), (n, 1))
            obs_s = tm.compute_anymal_observations(root_k, tk["commands"], dof_k[:, :, 0], np.tile(cd["default_dof_pos"], (n, 1)), dof_k[:, :, 1],
                                                   grav, a, 2.0, 0.25, 1.0, 0.05)
            rew_s, reset_s = tm.compute_anymal_reward(root_k, tk["commands"], tk["dof_force"], tk["contact"], cd["knee_bodies"], tk["progress"],
                                                      cd["rew_scales"], cd["base_body"], cd["max_episode_length"])
            np.testing.assert_allclose(tk["obs"], obs_s, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(tk["obs_clamped"], np.clip(obs_s, -5, 5), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(tk["rew"], rew_s, rtol=1e-5, atol=1e-8)
            # contact-force norms within 1e-4 N of the 1 N threshold are legitimately ambiguous in float32
            fn = np.concatenate([np.linalg.norm(tk["contact"][:, [cd["base_body"]], :], axis=2), np.linalg.norm(tk["contact"][:, cd["knee_bodies"], :], axis=2)], axis=1)
            sure = (np.abs(fn - 1.0) > 1e-4).all(axis=1)
            assert np.array_equal(tk["reset"][sure], reset_s.astype(np.int64)[sure])
            to_s = ((tk["progress"] >= cd["max_episode_length"] - 1) & (tk["reset"] != 0)).astype(np.int64)
            assert np.array_equal(tk["timeout"], to_s)
            # keep the oracle in lock-step with the kernel's discrete decisions
            assert np.array_equal(tk["reset"][sure], st["reset"][sure]), f"step {k}: reset decisions diverged"
            st["reset"][:] = tk["reset"]
            resets_seen += int(tk["reset"].sum())
            timeouts_seen += int(tk["timeout"].sum())
    finally:
        be.close()
    assert resets_seen > 0 and timeouts_seen > 0, "test must exercise resets and time-outs"


# ------------------------------------------------------------------------------------------------
def cartpole_params():
    sp = _abi.SimParams(dt=0.0166, substeps=2, num_position_iterations=4, num_velocity_iterations=0, contact_offset=0.02, rest_offset=0.001,
                        bounce_threshold_velocity=0.2, max_depenetration_velocity=100.0, plane_static_friction=1.0, plane_dynamic_friction=1.0,
                        plane_restitution=0.0, has_ground=1, joint_limit_stiffness=2000.0, joint_limit_damping=20.0)
    sp.gravity[2] = -9.81
    return sp


def cartpole_cfg(seed=42):
    return _abi.CartpoleCfg(reset_dist=3.0, max_push_effort=400.0, clip_obs=5.0, clip_actions=1.0, max_episode_length=500, seed=seed)


def cartpole_props(art):
    p = _abi.default_dof_props(art, _abi.DOF_MODE_NONE, 0.0, 0.0)
    p.drive_mode[0] = _abi.DOF_MODE_EFFORT
    return p


def check_cartpole_golden(make_backend):
    """Cartpole reward / reset (tasks/cartpole.py:180-196) through the kernel vs the reference's own outputs."""
    g = np.load(os.path.join(GOLDEN, "cartpole.npz"))
    art = load_robot("cartpole")
    n = g["pole_angle"].shape[0]
    be = make_backend(art, cartpole_params(), cartpole_props(art), n)
    try:
        c = cartpole_cfg()
        c.clip_obs = 3.0e38
        be.cartpole_create(c)
        root = np.zeros((n, 13), np.float32)
        root[:, 2], root[:, 6] = 2.0, 1.0
        dof = np.zeros((n, 2, 2), np.float32)
        dof[:, 0, 0], dof[:, 0, 1], dof[:, 1, 0], dof[:, 1, 1] = g["cart_pos"], g["cart_vel"], g["pole_angle"], g["pole_vel"]
        be.set_state(root, dof)
        be.set_task(progress=g["progress"] - 1, reset=np.zeros(n, np.int64))
        be.task_step(np.zeros((n, 1), np.float32), post_only=True)
        out = be.get_task()
    finally:
        be.close()
    keep = g["reset_buf"] == 0        # envs whose incoming reset_buf was 1 would be re-initialised first by post_physics_step
    np.testing.assert_allclose(out["rew"], g["rew"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out["reset"][keep], g["reset"][keep])
    np.testing.assert_allclose(out["obs"][:, 0], g["cart_pos"], rtol=0, atol=0)
    np.testing.assert_allclose(out["obs"][:, 2], g["pole_angle"], rtol=0, atol=0)


def check_cartpole_step(make_backend, n=32, steps=60, seed=5):
    """Fused Cartpole step vs oracle composition (float32 dynamics oracle + numpy reward in the reference's order)."""
    art = load_robot("cartpole")
    sp, props, c = cartpole_params(), cartpole_props(art), cartpole_cfg()
    c.max_episode_length = 25
    m = _abi.pack_model(art)
    rng = np.random.default_rng(seed)
    be = make_backend(art, sp, props, n)
    resets = 0
    try:
        be.cartpole_create(c)
        root = np.zeros((n, 13), np.float32)
        root[:, 2], root[:, 6] = 2.0, 1.0
        dof = np.zeros((n, 2, 2), np.float32)
        be.set_state(root, dof)
        progress, reset = np.zeros(n, np.int64), np.ones(n, np.int64)
        for k in range(steps):
            actions = rng.uniform(-1.2, 1.2, (n, 1)).astype(np.float32)
            draws = rng.uniform(0, 1, (n, 4)).astype(np.float32)
            be.task_step(actions, draws)
            a = np.clip(actions, -1, 1)
            act = np.zeros((n, 2), np.float32)
            act[:, 0] = a[:, 0] * np.float32(400.0)
            O.simulate(m, sp, props, root, dof, np.zeros((n, 2), np.float32), act)
            progress += 1
            ids = np.nonzero(reset)[0]
            dof[ids, :, 0] = np.float32(0.2) * (draws[ids, 0:2] - np.float32(0.5))
            dof[ids, :, 1] = np.float32(0.5) * (draws[ids, 2:4] - np.float32(0.5))
            reset[ids], progress[ids] = 0, 0
            rew, reset = tm.compute_cartpole_reward(dof[:, 1, 0], dof[:, 1, 1], dof[:, 0, 1], dof[:, 0, 0], 3.0, reset, progress, 25.0)
            rk, dk = be.get_state()
            t = be.get_task()
            assert np.abs(dk - dof).max() < 2e-3, f"step {k}: dof state deviates {np.abs(dk - dof).max():.2e}"
            assert np.array_equal(dk[ids], dof[ids])          # reset draws applied exactly
            obs = np.stack([dk[:, 0, 0], dk[:, 0, 1], dk[:, 1, 0], dk[:, 1, 1]], axis=1)
            np.testing.assert_array_equal(t["obs"], obs)
            np.testing.assert_array_equal(t["obs_clamped"], np.clip(obs, -5, 5))
            rew_s, reset_s = tm.compute_cartpole_reward(dk[:, 1, 0], dk[:, 1, 1], dk[:, 0, 1], dk[:, 0, 0], 3.0, np.zeros(n, np.int64), t["progress"], 25.0)
            np.testing.assert_allclose(t["rew"], rew_s, rtol=1e-5, atol=1e-6)
            assert np.array_equal(t["reset"], reset_s) and np.array_equal(t["progress"], progress)
            assert np.array_equal(t["timeout"], ((progress >= 24) & (t["reset"] != 0)).astype(np.int64))
            reset = t["reset"].copy()
            resets += int(reset.sum())
    finally:
        be.close()
    assert resets > 0


# ------------------------------------------------------------------------------------------------
# rough-terrain tasks
# ------------------------------------------------------------------------------------------------
def terrain_params():
    sp = flat_params(dt=0.005, substeps=1)
    return sp


def terrain_cfg_struct(cfg, nd=12, seed=42):
    """oracle cfg dict (tests/test_oracle_terrain.terrain_case) -> C b2g_terrain_cfg."""
    c = _abi.TerrainCfg()
    c.lin_vel_scale, c.ang_vel_scale, c.dof_pos_scale, c.dof_vel_scale = cfg["lin_vel_scale"], cfg["ang_vel_scale"], cfg["dof_pos_scale"], cfg["dof_vel_scale"]
    c.height_meas_scale, c.action_scale = cfg["height_meas_scale"], 0.5
    c.kp, c.kd, c.torque_limit = 80.0, 2.0, 80.0
    c.decimation, c.extra_sim_steps, c.dt = 4, 1, cfg["dt"]
    for i, v in enumerate(cfg["rew_scales"]):
        c.rew[i] = float(v)
    c.base_height_target = cfg["base_height_target"]
    c.clip_obs, c.clip_actions = 3.0e38, 3.0e38
    for i in range(2):
        c.cmd_x[i], c.cmd_y[i], c.cmd_yaw[i] = cfg["cmd_x"][i], cfg["cmd_y"][i], cfg["cmd_yaw"][i]
    for i, v in enumerate(cfg["default_dof_pos"]):
        c.default_dof_pos[i] = float(v)
    for i, v in enumerate(cfg["init_root"]):
        c.init_root[i] = float(v)
    nv = cfg.get("noise_scale_vec")
    c.add_noise = 0 if nv is None else 1
    if nv is not None:
        c.noise_lin_vel, c.noise_ang_vel, c.noise_gravity = float(nv[0]), float(nv[3]), float(nv[6])
        c.noise_dof_pos, c.noise_dof_vel, c.noise_height = float(nv[12]), float(nv[24]), float(nv[36])
    c.base_body = int(cfg["base_body"])
    c.n_knee, c.n_feet = len(cfg["knee"]), len(cfg["feet"])
    for i, k in enumerate(cfg["knee"]):
        c.knee_bodies[i] = int(k)
    for i, k in enumerate(cfg["feet"]):
        c.feet_bodies[i] = int(k)
    c.hound_termination = 1 if cfg.get("hound") else 0
    extra = list(cfg.get("base_indices", []))
    c.n_term_extra = len(extra)
    for i, k in enumerate(extra):
        c.term_extra_bodies[i] = int(k)
    c.allow_knee_contacts = 1 if cfg["allow_knee"] else 0
    for i, d in enumerate([0, 3, 6, 9]):
        c.hip_dofs[i] = d
    c.max_episode_length = int(cfg["max_len"])
    c.push_interval = int(cfg.get("push_interval", 750))
    c.max_episode_length_s = float(cfg["max_episode_length_s"])
    c.custom_origins, c.curriculum = int(cfg["custom_origins"]), int(cfg["curriculum"])
    c.n_hx, c.n_hy = len(tm.HEIGHT_X), len(tm.HEIGHT_Y)
    for i, v in enumerate(tm.HEIGHT_X):
        c.hx[i] = float(v)
    for i, v in enumerate(tm.HEIGHT_Y):
        c.hy[i] = float(v)
    t = cfg.get("terrain")
    if t:
        c.hs_rows, c.hs_cols = t["height_samples"].shape
        c.border_size, c.hscale, c.vscale, c.env_length = t["border_size"], t["hscale"], t["vscale"], t["env_length"]
        c.env_rows, c.env_cols = t["terrain_origins"].shape[0], t["terrain_origins"].shape[1]
    c.seed = seed
    c.arm_chain = -1
    return c


def check_terrain_golden(make_backend, name):
    """post_physics_step of the terrain tasks through the two kernels vs the reference's own outputs (golden vectors made by
    executing the reference's eager methods): 1e-5 relative, masks / counters / curriculum levels bit-exact."""
    from tests.test_oracle_terrain import check_terrain_outputs, terrain_case

    g, st, cfg, draws = terrain_case(name)
    robot = "hound" if "hound" in name else "anymal_minimal"
    art = load_robot(robot)
    n = st["root"].shape[0]
    cfg["push_interval"] = int(g["push_interval"])
    c = terrain_cfg_struct(cfg)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    be = make_backend(art, terrain_params(), props, n)
    try:
        t = cfg.get("terrain")
        be.terrain_create(c, t["height_samples"] if t else None, t["terrain_origins"] if t else None)
        be.set_step(int(g["common_step_counter"]) + 1, 1)
        be.set_state(st["root"], np.stack([st["dof_pos"], st["dof_vel"]], axis=2))
        kw = dict(commands=st["commands"], progress=st["progress"], timeout=st["timeout_prev"].astype(np.int64), torques=st["torques"],
                  last_actions=st["last_actions"], last_dof_vel=st["last_dof_vel"], feet_air_time=st["feet_air_time"], episode_sums=st["episode_sums"],
                  contact=st["contact"], reset=np.zeros(n, np.int64))
        if cfg["custom_origins"]:
            kw.update(env_origins=st["env_origins"], terrain_levels=st["terrain_levels"], terrain_types=st["terrain_types"])
        be.set_task(**kw)
        be.task_step(st["actions"], draws, post_only=True)
        out = be.get_task()
        root, dof = be.get_state()
    finally:
        be.close()
    res = dict(root=root, dof_pos=dof[:, :, 0], dof_vel=dof[:, :, 1], commands=out["commands"], last_actions=out["last_actions"],
               last_dof_vel=out["last_dof_vel"], feet_air_time=out["feet_air_time"], episode_sums=out["episode_sums"], progress=out["progress"])
    if cfg["custom_origins"]:
        res.update(terrain_levels=out["terrain_levels"], env_origins=out["env_origins"])
    extras = out["extras"][:14] if "extras" in out else None
    if extras is not None and not cfg["custom_origins"]:
        extras[13] = g["o_extras"][13]
    check_terrain_outputs(g, res, out["obs"], out["rew"], out["reset"], out["timeout"], out["measured"], extras)


def check_terrain_step(make_backend, robot="anymal_minimal", n=8, steps=12, seed=11, heightfield=True):
    """Whole step of the terrain task (4 PD sim steps + 1 stale, post_physics_step) vs the oracle composition: float32 dynamics
    oracle + numpy task math in the reference's order, same injected draws, on a generated heightfield with curriculum."""
    from isaacgymenv_b200.terrain import Terrain

    art = load_robot(robot)
    nd, nb = art.num_dofs, art.num_bodies
    rng = np.random.default_rng(seed)
    sp = terrain_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    m = _abi.pack_model(art)
    hound = robot == "hound"
    knee = [i for i, b in enumerate(art.body_names) if ("thigh" if hound else "THIGH") in b]
    feet = [i for i, b in enumerate(art.body_names) if ("foot" if hound else "SHANK") in b]
    terrain = hf = hf_t = None
    if heightfield:
        tcfg = dict(terrainType="trimesh", curriculum=True, mapLength=8.0, mapWidth=8.0, numLevels=3, numTerrains=4,
                    terrainProportions=[0.1, 0.1, 0.35, 0.25, 0.2], slopeTreshold=0.5)
        tr = Terrain(tcfg, n, seed=3)
        terrain = dict(height_samples=tr.heightsamples, border_size=float(tr.border_size), hscale=tr.horizontal_scale, vscale=tr.vertical_scale,
                       env_length=tr.env_length, env_rows=tr.env_rows, terrain_origins=tr.env_origins.astype(np.float32))
        hf_t = _abi.Heightfield(rows=tr.tot_rows, cols=tr.tot_cols, horizontal_scale=tr.horizontal_scale, vertical_scale=tr.vertical_scale,
                                origin_x=-tr.border_size, origin_y=-tr.border_size, friction=1.0, restitution=0.0)
        hf = (hf_t, tr.heightsamples)
    dt = 0.02
    raw = np.array([-1.0, 1.0, -4.0, 0.5, -0.05, -0.2, -0.00002, -0.0005, -0.5, 1.0, -0.25, -0.1, -0.01, -0.05], np.float32) * np.float32(dt)
    nv = np.zeros(12 + 2 * nd + 140 + nd, np.float32)
    nv[:3], nv[3:6], nv[6:9], nv[12:24], nv[24:36], nv[36:176] = 0.2, 0.05, 0.05, 0.01, 0.075, 0.3
    cfg = dict(rew_scales=raw, knee=np.array(knee), feet=np.array(feet), base_indices=np.array([i for i, b in enumerate(art.body_names) if "shoulder" in b]),
               base_body=0, allow_knee=not hound and False, hound=hound, base_height_target=0.48 if hound else 0.52, noise_scale_vec=nv, dt=dt, max_len=9,
               push=False, default_dof_pos=default_pose(art).astype(np.float32), init_root=np.array([0, 0, 0.62, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], np.float32),
               cmd_x=[-1, 1], cmd_y=[-1, 1], cmd_yaw=[-3.14, 3.14], custom_origins=heightfield, curriculum=True, terrain=terrain,
               max_episode_length_s=20.0, lin_vel_scale=2.0, ang_vel_scale=0.25, dof_pos_scale=1.0, dof_vel_scale=0.05, height_meas_scale=5.0,
               push_interval=5)
    c = terrain_cfg_struct(cfg)
    be = make_backend(art, sp, props, n)
    resets = 0
    try:
        be.terrain_create(c, terrain["height_samples"] if terrain else None, terrain["terrain_origins"] if terrain else None, heightfield=hf)
        # start on the terrain origins, standing
        root, dof = standing_state(art, n, rng, 0.6)
        st = dict(root=root, dof_pos=dof[:, :, 0].copy(), dof_vel=dof[:, :, 1].copy(), contact=np.zeros((n, nb, 3), np.float32), torques=np.zeros((n, nd), np.float32),
                  commands=np.zeros((n, 4), np.float32), actions=np.zeros((n, nd), np.float32), last_actions=np.zeros((n, nd), np.float32),
                  last_dof_vel=np.zeros((n, nd), np.float32), feet_air_time=np.zeros((n, 4), np.float32), progress=np.zeros(n, np.int64),
                  timeout_prev=np.zeros(n, bool), episode_sums=np.zeros((13, n), np.float32), terrain_levels=(np.arange(n) % 3).astype(np.int64),
                  terrain_types=(np.arange(n) % 4).astype(np.int64))
        st["commands"][:, 0], st["commands"][:, 3] = 0.7, 0.5
        if heightfield:
            st["env_origins"] = terrain["terrain_origins"][st["terrain_levels"], st["terrain_types"]].copy()
            st["root"][:, :3] += st["env_origins"]
            be.set_task(env_origins=st["env_origins"], terrain_levels=st["terrain_levels"], terrain_types=st["terrain_types"])
        else:
            st["env_origins"] = np.zeros((n, 3), np.float32)
            st["terrain_levels"][:] = 0
            st["terrain_types"][:] = 0
        be.set_state(st["root"], np.stack([st["dof_pos"], st["dof_vel"]], axis=2))
        be.set_task(commands=st["commands"], reset=np.zeros(n, np.int64))
        for k in range(steps):
            step = k + 1
            actions = rng.uniform(-1, 1, (n, nd)).astype(np.float32)
            draws = dict(reset=rng.uniform(0, 1, (n, 2 * nd + 5)).astype(np.float32), noise=rng.uniform(0, 1, (n, len(nv))).astype(np.float32),
                         push=rng.uniform(0, 1, (n, 2)).astype(np.float32))
            be.set_step(step, 1)
            be.task_step(actions, draws)
            # oracle: decimation loop + extra sim step
            st["actions"] = actions.copy()
            dofs = np.stack([st["dof_pos"], st["dof_vel"]], axis=2).astype(np.float32)
            for it in range(5):
                if it < 4:
                    tq = np.clip(np.float32(80.0) * (np.float32(0.5) * actions + cfg["default_dof_pos"][None] - dofs[:, :, 0]) - np.float32(2.0) * dofs[:, :, 1], -80, 80).astype(np.float32)
                f, cf = O.simulate(m, sp, props, st["root"], dofs, np.zeros((n, nd), np.float32), tq, heightfield=hf_t, hf_samples=hf[1] if hf else None)
            st["dof_pos"], st["dof_vel"], st["torques"], st["contact"] = dofs[:, :, 0].copy(), dofs[:, :, 1].copy(), tq, cf
            cfg["push"] = step % cfg["push_interval"] == 0
            obs, rew, reset, timeout, measured, extras = tm.terrain_post_physics(st, cfg, draws)
            st["timeout_prev"] = timeout.astype(bool)
            rk, dk = be.get_state()
            t = be.get_task()
            assert np.array_equal(t["reset"], reset), f"step {k}: reset decisions differ"
            assert np.array_equal(t["progress"], st["progress"]) and np.array_equal(t["timeout"], timeout)
            assert np.abs(rk - st["root"]).max() < 3e-3, f"step {k}: root deviates {np.abs(rk - st['root']).max():.2e}"
            assert np.abs(dk[:, :, 0] - st["dof_pos"]).max() < 5e-3
            np.testing.assert_allclose(t["obs"], obs, rtol=0, atol=3e-2)
            np.testing.assert_allclose(t["rew"], rew, rtol=0, atol=2e-3)
            np.testing.assert_allclose(t["measured"], measured, rtol=0, atol=0.051)     # a sample may fall in the neighbouring 0.1 m cell
            assert np.array_equal(t["terrain_levels"], st["terrain_levels"])
            np.testing.assert_allclose(t["commands"], st["commands"], rtol=0, atol=2e-3)
            np.testing.assert_allclose(t["feet_air_time"], st["feet_air_time"], rtol=0, atol=1e-6)
            # keep the oracle glued to the kernel state so that tolerances do not accumulate over steps
            st["root"], st["dof_pos"], st["dof_vel"] = rk.copy(), dk[:, :, 0].copy(), dk[:, :, 1].copy()
            st["commands"], st["last_dof_vel"], st["episode_sums"] = t["commands"].copy(), t["last_dof_vel"].copy(), t["episode_sums"].copy()
            resets += int(reset.sum())
    finally:
        be.close()
    assert resets > 0


# ------------------------------------------------------------------------------------------------
# hound + arm (UsefulHound)
# ------------------------------------------------------------------------------------------------
def useful_cfg_struct(cfg, art):
    c = terrain_cfg_struct(cfg, nd=18)
    c.n_ctrl_dof, c.arm_chain = 12, 4
    c.arm_kp, c.arm_kp_null, c.arm_action_scale, c.arm_dof_noise = 150.0, 10.0, 1.0, 0.25
    for i, v in enumerate([0.1, 0.1, 0.1, 0.5, 0.5, 0.5]):
        c.arm_cmd_limit[i] = v
    c.eef_body = art.body_names.index("end_link")
    c.jac_body = art.joint_dict["joint6"]
    c.refresh_eef = 0
    c.hound_termination = 1
    return c


def check_useful_golden(make_backend):
    """UsefulHound post_physics_step and OSC torque law through the kernels vs the reference's own outputs."""
    from tests.test_oracle_terrain import useful_case

    g, st, cfg, draws = useful_case()
    art = load_robot("useful_hound")
    n = st["root"].shape[0]
    cfg["push_interval"] = int(g["push_interval"])
    c = useful_cfg_struct(cfg, art)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    be = make_backend(art, terrain_params(), props, n)
    try:
        be.terrain_create(c)
        # (a) the OSC law on the golden inputs
        dof = np.zeros((n, 18, 2), np.float32)
        dof[:, 12:, 0], dof[:, 12:, 1] = g["osc_q"], g["osc_qd"]
        be.set_state(st["root"], dof)
        eef = np.zeros((n, 13), np.float32)
        eef[:, 7:] = g["osc_eef_vel"]
        be.set_task(arm_mm=g["osc_mm"], arm_jac=g["osc_j"], eef_state=eef)
        acts = np.zeros((n, 18), np.float32)
        acts[:, 12:] = g["osc_dpose"] / np.array([[0.1, 0.1, 0.1, 0.5, 0.5, 0.5]], np.float32)
        be.task_step(acts, None, post_only=2)
        u = be.get_task()["torques"][:, 12:]
        np.testing.assert_allclose(u, g["osc_u"], rtol=3e-4, atol=3e-4)
        # (b) post_physics_step
        be.set_step(int(g["common_step_counter"]) + 1, 1)
        be.set_state(st["root"], g["dof_state"])
        be.set_task(commands=st["commands"], progress=st["progress"], timeout=st["timeout_prev"].astype(np.int64), torques=st["torques"],
                    last_actions=st["last_actions"], last_dof_vel=np.concatenate([st["last_dof_vel"], np.zeros((n, 6), np.float32)], axis=1),
                    feet_air_time=st["feet_air_time"], episode_sums=st["episode_sums"], contact=st["contact"], reset=np.zeros(n, np.int64),
                    eef_state=st["eef_state"])
        be.task_step(st["actions"], draws, post_only=True)
        out = be.get_task()
        root, dof = be.get_state()
    finally:
        be.close()
    assert np.array_equal(out["reset"], g["o_reset"]) and np.array_equal(out["progress"], g["o_progress"]) and np.array_equal(out["timeout"], g["o_timeout"])
    np.testing.assert_allclose(out["rew"], g["o_rew"], rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(out["obs"], g["o_obs"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(dof, g["o_dof_state"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(root, g["o_root"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["commands"], g["o_commands"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["last_actions"], g["o_last_actions"], rtol=0, atol=0)
    np.testing.assert_allclose(out["last_dof_vel"][:, :12], g["o_last_dof_vel"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["episode_sums"], g["o_episode_sums"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["feet_air_time"], g["o_feet_air_time"], rtol=1e-5, atol=1e-7)


def check_useful_step(make_backend, n=6, steps=8, seed=21):
    """Whole UsefulHound step (PD legs + OSC arm x 4 sim steps + 1 stale, post_physics_step) vs the oracle composition.  The oracle's arm
    mass-matrix block comes from its own CRBA (float64), the Jacobian slice from the independent numpy kinematics."""
    from isaacgymenv_b200.model import urdf

    art = load_robot("useful_hound")
    nd, nb = art.num_dofs, art.num_bodies
    rng = np.random.default_rng(seed)
    sp = terrain_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    m = _abi.pack_model(art)
    knee = [i for i, b in enumerate(art.body_names) if "thigh" in b]
    feet = [i for i, b in enumerate(art.body_names) if "foot" in b]
    shoulders = [i for i, b in enumerate(art.body_names) if "shoulder" in b]
    dt = 0.02
    raw = np.array([-1.0, 1.0, -4.0, 0.5, -0.05, -1.0, -0.00002, -0.0005, -4.0, 1.0, -0.25, -0.0, -0.01, -0.0], np.float32) * np.float32(dt)
    q0 = default_pose(art).astype(np.float32)
    cfg = dict(rew_scales=raw, knee=np.array(knee), feet=np.array(feet), base_indices=np.array(shoulders), base_body=0, allow_knee=True, hound=True,
               base_height_target=0.52, noise_scale_vec=None, dt=dt, max_len=7, push=False, default_dof_pos=q0[:12],
               init_root=np.array([0, 0, 0.62, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], np.float32), cmd_x=[-2, 2], cmd_y=[-1, 1], cmd_yaw=[-1, 1],
               custom_origins=False, curriculum=True, terrain=None, max_episode_length_s=20.0, lin_vel_scale=2.0, ang_vel_scale=0.25, dof_pos_scale=1.0,
               dof_vel_scale=0.05, height_meas_scale=5.0, push_interval=1000,
               arm=dict(dof_noise=0.25, lower=np.full(6, -1.57, np.float32), upper=np.full(6, 1.57, np.float32)))
    c = useful_cfg_struct(cfg, art)
    for i, v in enumerate(q0):
        c.default_dof_pos[i] = float(v)
    jac_body = int(c.jac_body)

    def arm_mats(root, dofs):
        mm = np.zeros((n, 6, 6), np.float32)
        jj = np.zeros((n, 6, 6), np.float32)
        for e in range(n):
            H, _ = O.crba_rnea(m, sp, root[e].astype(np.float64), dofs[e].astype(np.float64))
            mm[e] = H[18:24, 18:24]
            bp, _ = urdf.body_poses(art, dofs[e, :, 0].astype(np.float64), root[e, :3], root[e, 3:7])
            r = bp[jac_body] - root[e, :3]
            jj[e] = np.eye(6)
            jj[e, 0, 4], jj[e, 0, 5], jj[e, 1, 3], jj[e, 1, 5], jj[e, 2, 3], jj[e, 2, 4] = r[2], -r[1], -r[2], r[0], r[1], -r[0]
        return mm, jj

    be = make_backend(art, sp, props, n)
    resets = 0
    try:
        be.terrain_create(c)
        root, dof = standing_state(art, n, rng, 0.55)
        dof[:, 12:, 0] = rng.uniform(-0.3, 0.3, (n, 6))
        st = dict(root=root, dof_pos=dof[:, :12, 0].copy(), dof_vel=dof[:, :12, 1].copy(), arm_q=dof[:, 12:, 0].copy(), arm_qd=dof[:, 12:, 1].copy(),
                  contact=np.zeros((n, nb, 3), np.float32), torques=np.zeros((n, nd), np.float32), commands=np.zeros((n, 4), np.float32),
                  actions=np.zeros((n, nd), np.float32), last_actions=np.zeros((n, nd), np.float32), last_dof_vel=np.zeros((n, 12), np.float32),
                  feet_air_time=np.zeros((n, 4), np.float32), progress=np.zeros(n, np.int64), timeout_prev=np.zeros(n, bool),
                  episode_sums=np.zeros((13, n), np.float32), terrain_levels=np.zeros(n, np.int64), eef_state=np.zeros((n, 13), np.float32),
                  arm_commands=np.zeros((n, 3), np.float32))
        st["commands"][:, 0] = 0.5
        mm, jj = arm_mats(root, dof)
        be.set_state(root, dof)
        be.set_task(commands=st["commands"], reset=np.zeros(n, np.int64), arm_mm=mm, arm_jac=jj)
        for k in range(steps):
            actions = rng.uniform(-1, 1, (n, nd)).astype(np.float32)
            draws = dict(reset=rng.uniform(0, 1, (n, 35)).astype(np.float32), noise=np.zeros((n, 204), np.float32), push=np.zeros((n, 2), np.float32))
            be.set_step(k + 1, 1)
            be.task_step(actions, draws)
            st["actions"] = actions.copy()
            dofs = np.concatenate([np.stack([st["dof_pos"], st["dof_vel"]], axis=2), np.stack([st["arm_q"], st["arm_qd"]], axis=2)], axis=1).astype(np.float32)
            dpose = actions[:, 12:] * np.array([[0.1, 0.1, 0.1, 0.5, 0.5, 0.5]], np.float32)
            for it in range(5):
                if it < 4:
                    ua = tm.osc_torques(mm, jj, dpose, st["eef_state"][:, 7:], dofs[:, 12:, 0], dofs[:, 12:, 1])
                    tl = np.clip(np.float32(80.0) * (np.float32(0.5) * actions[:, :12] + q0[None, :12] - dofs[:, :12, 0]) - np.float32(2.0) * dofs[:, :12, 1], -80, 80)
                    tq = np.concatenate([tl, ua], axis=1).astype(np.float32)
                f, cf = O.simulate(m, sp, props, st["root"], dofs, np.zeros((n, nd), np.float32), tq)
            st["dof_pos"], st["dof_vel"], st["arm_q"], st["arm_qd"] = dofs[:, :12, 0].copy(), dofs[:, :12, 1].copy(), dofs[:, 12:, 0].copy(), dofs[:, 12:, 1].copy()
            st["torques"], st["contact"] = tq, cf
            mm, jj = arm_mats(st["root"], dofs)          # refresh_jacobian / refresh_mass_matrix happen before reset_idx
            obs, rew, reset, timeout, measured, extras = tm.terrain_post_physics(st, cfg, draws)
            st["timeout_prev"] = timeout.astype(bool)
            rk, dk = be.get_state()
            t = be.get_task()
            assert np.array_equal(t["reset"], reset), f"step {k}: reset decisions differ"
            assert np.abs(rk - st["root"]).max() < 3e-3, f"step {k}: root deviates {np.abs(rk - st['root']).max():.2e}"
            assert np.abs(dk[:, :12, 0] - st["dof_pos"]).max() < 5e-3 and np.abs(dk[:, 12:, 0] - st["arm_q"]).max() < 5e-3
            np.testing.assert_allclose(t["torques"], tq, rtol=0, atol=0.2)
            np.testing.assert_allclose(t["arm_mm"], mm, rtol=2e-3, atol=2e-4)
            np.testing.assert_allclose(t["arm_jac"], jj, rtol=0, atol=2e-3)
            np.testing.assert_allclose(t["obs"], obs, rtol=0, atol=3e-2)
            np.testing.assert_allclose(t["rew"], rew, rtol=0, atol=2e-3)
            st["root"], st["dof_pos"], st["dof_vel"], st["arm_q"], st["arm_qd"] = rk.copy(), dk[:, :12, 0].copy(), dk[:, :12, 1].copy(), dk[:, 12:, 0].copy(), dk[:, 12:, 1].copy()
            st["commands"], st["last_dof_vel"], st["episode_sums"] = t["commands"].copy(), t["last_dof_vel"][:, :12].copy(), t["episode_sums"].copy()
            mm, jj = t["arm_mm"].copy(), t["arm_jac"].copy()
            resets += int(reset.sum())
    finally:
        be.close()
    assert resets > 0


# ------------------------------------------------------------------------------------------------
def check_jacobian_mass_matrix(make_backend, robot="useful_hound", n=4, seed=31):
    """acquire/refresh_jacobian_tensor and _mass_matrix_tensor: the Jacobian against finite differences of the independent numpy
    kinematics (body-frame origin velocity per unit generalized velocity), the mass matrix against the oracle's CRBA joint block."""
    from isaacgymenv_b200.model import urdf

    art = load_robot(robot)
    nd, nb = art.num_dofs, art.num_bodies
    rng = np.random.default_rng(seed)
    sp = flat_params(ground=False)
    props = _abi.default_dof_props(art)
    root, dof = random_flying_state(art, n, rng)
    if art.fixed_base:
        root[:, :3], root[:, 3:7] = [0, 0, 2.0], [0, 0, 0, 1]
    be = make_backend(art, sp, props, n)
    try:
        be.set_state(root, dof)
        jac, mm = be.jacobian_mass_matrix()
    finally:
        be.close()
    m = _abi.pack_model(art)
    fb = 0 if art.fixed_base else 6
    row0 = 1 if art.fixed_base else 0
    assert jac.shape == (n, nb - row0, 6, nd + fb) and mm.shape == (n, nd, nd)
    for e in range(n):
        H, _ = O.crba_rnea(m, sp, root[e].astype(np.float64), dof[e].astype(np.float64))
        np.testing.assert_allclose(mm[e], H[fb:, fb:], rtol=2e-4, atol=2e-5)
        q = dof[e, :, 0].astype(np.float64)
        p0, r0 = urdf.body_poses(art, q, root[e, :3], root[e, 3:7])
        eps = 1e-6
        for d in range(nd):
            qq = q.copy()
            qq[d] += eps
            p1, r1 = urdf.body_poses(art, qq, root[e, :3], root[e, 3:7])
            lin = (p1 - p0) / eps
            for b in range(row0, nb):
                W = ((r1[b] - r0[b]) / eps) @ r0[b].T
                ang = np.array([W[2, 1], W[0, 2], W[1, 0]])
                np.testing.assert_allclose(jac[e, b - row0, :3, fb + d], lin[b], rtol=0, atol=2e-4)
                np.testing.assert_allclose(jac[e, b - row0, 3:, fb + d], ang, rtol=0, atol=2e-4)
        if fb:
            for b in range(nb):
                r = p0[b] - root[e, :3]
                want = np.eye(6)
                want[:3, 3:] = -np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
                np.testing.assert_allclose(jac[e, b, :, :6], want, rtol=0, atol=2e-5)


# ------------------------------------------------------------------------------------------------
# Houndarm (tasks/hound_arm.py): fused step vs an oracle composition
# ------------------------------------------------------------------------------------------------
def houndarm_params():
    sp = _abi.SimParams(dt=0.01667, substeps=2, num_position_iterations=8, num_velocity_iterations=1, contact_offset=0.005, rest_offset=0.0,
                        bounce_threshold_velocity=0.2, max_depenetration_velocity=1000.0, plane_static_friction=1.0, plane_dynamic_friction=1.0,
                        plane_restitution=0.0, has_ground=1, joint_limit_stiffness=2000.0, joint_limit_damping=20.0)
    return sp          # gravity zero: asset_options.disable_gravity (tasks/hound_arm.py:212)


FRANKA_DEFAULT = (0.0, 0.1963, 0.0, -2.6180, 0.0, 2.9416, 0.7854)      # tasks/manipulator.py:153-155


def houndarm_cfg(art, seed=42):
    """Houndarm (6-DOF open_manipulator_p) or, for the 7-DOF Franka model, Manipulator (tasks/manipulator.py, cfg/task/Manipulator.yaml)."""
    franka = art.num_dofs == 7
    c = _abi.HoundarmCfg(clip_obs=5.0, clip_actions=1.0, action_scale=1.0, dof_noise=0.25, kp=150.0, kp_null=10.0, dist_scale=0.1, vel_scale=0.1,
                         eef_body=art.body_names.index("panda_link7" if franka else "end_link"),
                         jac_body=art.joint_dict["panda_joint7" if franka else "joint6"] + 1, max_episode_length=1000 if franka else 150, seed=seed,
                         n_reset_tail=2 if franka else 0)
    for i, v in enumerate([0.1, 0.1, 0.1, 0.5, 0.5, 0.5]):
        c.cmd_limit[i] = v
    for i, v in enumerate([-0.5, 0.5, -0.5, 0.5, 0.2, 0.6] if franka else [-0.3, 0.3, -0.3, 0.3, 0.1, 0.3]):
        c.cmd_range[i] = v
    if franka:
        for i, v in enumerate(FRANKA_DEFAULT):
            c.default_dof_pos[i] = v
    return c


def _mat_to_quat(R):
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    q[3] = (R[k, j] - R[j, k]) / s
    return q


def houndarm_eef(art, root, dof, body):
    """End-effector rigid-body row (pos, quat xyzw, linear velocity, angular velocity) from the model compiler's numpy kinematics
    and central differences along qd -- independent of the kernels."""
    from isaacgymenv_b200.model.urdf import body_poses

    n = root.shape[0]
    out = np.zeros((n, 13), np.float64)
    eps = 1e-6
    for e in range(n):
        q, qd = dof[e, :, 0].astype(np.float64), dof[e, :, 1].astype(np.float64)
        rp, rq = root[e, :3].astype(np.float64), root[e, 3:7].astype(np.float64)
        pos, rot = body_poses(art, q, rp, rq)
        out[e, :3], out[e, 3:7] = pos[body], _mat_to_quat(rot[body])
        p1, r1 = body_poses(art, q + eps * qd, rp, rq)
        p0, r0 = body_poses(art, q - eps * qd, rp, rq)
        out[e, 7:10] = (p1[body] - p0[body]) / (2 * eps)
        W = (r1[body] @ r0[body].T - np.eye(3)) / (2 * eps)
        out[e, 10:13] = [0.5 * (W[2, 1] - W[1, 2]), 0.5 * (W[0, 2] - W[2, 0]), 0.5 * (W[1, 0] - W[0, 1])]
    return out


def arm_reset_positions(default_q, noise, draws, lower, upper, tail):
    """Joint positions after reset_idx of the arm reach tasks (hound_arm.py:441-446 / manipulator.py:407-417), float32 like the reference:
    clamp(default + noise * 2 (u - 0.5), limits), then the last ``tail`` joints set back to their default (pinned to the reference's own
    reset_idx by tests/test_manipulator.py)."""
    q = np.clip(np.asarray(default_q, np.float32) + np.float32(noise) * np.float32(2.0) * (np.asarray(draws, np.float32) - np.float32(0.5)),
                np.asarray(lower, np.float32), np.asarray(upper, np.float32))
    if tail:
        q[:, -tail:] = np.asarray(default_q, np.float32)[-tail:]
    return q


def check_houndarm_step(make_backend, n=12, steps=40, seed=23, robot="houndarm"):
    """Fused Houndarm step vs a composition of independent pieces: OSC torques (numpy restatement of the reference law) from the
    kernels' own Jacobian / mass-matrix tensors (themselves pinned to the oracle elsewhere), the float64 dynamics oracle for the
    sub-steps, reset draws, and observations / reward from the model compiler's numpy kinematics."""
    art = load_robot(robot)      # "manipulator": the same task on the 7-DOF Franka (six actions, seven torques, posture + reset quirks)
    sp, c = houndarm_params(), houndarm_cfg(art)
    c.max_episode_length = 17
    default_q = np.array([c.default_dof_pos[i] for i in range(art.num_dofs)], np.float32)
    tail = int(c.n_reset_tail)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    m = _abi.pack_model(art)
    nd = art.num_dofs
    rng = np.random.default_rng(seed)
    be = make_backend(art, sp, props, n)
    jm = make_backend(art, sp, props, n)          # twin used only for Jacobian / mass-matrix tensors at the same state
    lower, upper = np.array(art.lower, np.float32), np.array(art.upper, np.float32)
    effort = np.array(art.effort, np.float32)
    resets = timeouts = checked_torques = 0
    torque_dev = []
    try:
        be.houndarm_create(c)
        root = np.zeros((n, 13), np.float32)
        root[:, 0], root[:, 6] = -0.45, 1.0
        dof = np.zeros((n, nd, 2), np.float32)
        dof[:, :, 0] = np.clip(default_q + rng.uniform(-0.3, 0.3, (n, nd)), lower, upper)
        be.set_state(root, dof)
        commands = rng.uniform(-0.2, 0.3, (n, 3)).astype(np.float32)
        commands[:, 2] = np.abs(commands[:, 2])
        be.set_task(commands=commands, reset=np.zeros(n, np.int64), progress=np.zeros(n, np.int64))
        r64, d64 = root.astype(np.float64), dof.astype(np.float64)
        progress, reset = np.zeros(n, np.int64), np.zeros(n, np.int64)
        for k in range(steps):
            # moderate commands (a tenth of the entries beyond the clip range): violent motion drives the arm into configurations where
            # float32 cannot evaluate the law at all (see the tolerance note below)
            actions = (rng.uniform(-0.5, 0.5, (n, 6)) * np.where(rng.uniform(0, 1, (n, 6)) < 0.1, 3.0, 1.0)).astype(np.float32)
            draws = rng.uniform(0, 1, (n, 3 + nd)).astype(np.float32)
            # --- oracle composition ---
            a = np.clip(actions, -1, 1)
            jm.set_state(r64.astype(np.float32), d64.astype(np.float32))
            J, MM = jm.jacobian_mass_matrix()
            eef = houndarm_eef(art, r64, d64, c.eef_body)
            dpose = a * np.array([0.1, 0.1, 0.1, 0.5, 0.5, 0.5], np.float32) / np.float32(1.0)
            u = tm.osc_torques(MM[:, :nd, :nd].astype(np.float32), J[:, c.jac_body - 1, :, :nd].astype(np.float32), dpose.astype(np.float32),
                               eef[:, 7:].astype(np.float32), d64[:, :nd, 0].astype(np.float32), d64[:, :nd, 1].astype(np.float32), 150.0, 10.0, effort[:nd], exact=True,
                               default_q=default_q)
            # --- kernel ---
            be.task_step(actions, draws)
            rk, dk = be.get_state()
            t = be.get_task()
            # the operational-space torques.  The law inverts J M^-1 J^T, whose condition number on this arm (last link 19 g: mass
            # matrix 2e4, Jacobian 70-10000) is 4e6 ... 3e10 along a rollout.  A float32 evaluation -- the reference's torch.inverse
            # included -- is off by eps x cond there (measured: percent-level to total), which is why the kernels evaluate the law
            # in float64 (b2g_threads.cuh osc_prepare) and why the comparison is against the float64 restatement: agreement to
            # ~1e-6 of the largest torque wherever the problem is not outright singular (< 1e10), asserted after the loop.  The
            # oracle is advanced with the kernel's torques so that the dynamics comparison below stays a dynamics comparison.
            uk = t["dof_force"]
            jj = J[:, c.jac_body - 1, :, :nd].astype(np.float64)
            cond = np.array([np.linalg.cond(jj[e] @ np.linalg.inv(MM[e, :nd, :nd].astype(np.float64)) @ jj[e].T) for e in range(n)])
            ok = cond < 1e10
            checked_torques += int(ok.sum())
            if ok.any():
                torque_dev.extend((np.abs(uk[ok] - u[ok]).max(axis=1) / np.maximum(1.0, np.abs(u[ok]).max(axis=1))).tolist())
            O.simulate(m, sp, props, r64, d64, np.zeros((n, nd)), uk.astype(np.float64))
            progress += 1
            ids = np.nonzero(reset)[0]
            for kk in range(3):
                commands[ids, kk] = (c.cmd_range[2 * kk + 1] - c.cmd_range[2 * kk]) * draws[ids, kk] + c.cmd_range[2 * kk]
            newq = arm_reset_positions(default_q, 0.25, draws[ids, 3:3 + nd], lower, upper, tail)
            d64[ids, :, 0], d64[ids, :, 1] = newq, 0.0
            progress[ids], reset[ids] = 0, 0
            eef = houndarm_eef(art, r64, d64, c.eef_body)
            dist = np.linalg.norm(eef[:, :3] - commands, axis=1)
            rew = np.maximum((1 - np.tanh(10 * dist)) * 0.1 + (1 - np.tanh(10 * np.linalg.norm(eef[:, 7:], axis=1))) * (dist < 0.02) * 0.1, 0.0)
            reset = np.where(progress >= c.max_episode_length - 1, 1, reset)
            assert np.abs(dk[:, :, 0] - d64[:, :, 0]).max() < 2e-3, f"step {k}: joint positions deviate {np.abs(dk[:, :, 0] - d64[:, :, 0]).max():.2e}"
            assert np.abs(dk[:, :, 1] - d64[:, :, 1]).max() < 5e-2 * max(1.0, np.abs(d64[:, :, 1]).max()), f"step {k}: joint velocities deviate"
            np.testing.assert_array_equal(dk[ids, :, 0], newq)                       # reset draws applied exactly
            np.testing.assert_array_equal(t["commands"], commands)
            np.testing.assert_allclose(t["obs"][:, :3], eef[:, :3], atol=3e-3)
            sgn = np.sign((t["obs"][:, 3:7] * eef[:, 3:7]).sum(1))[:, None]
            np.testing.assert_allclose(t["obs"][:, 3:7] * sgn, eef[:, 3:7], atol=3e-3)
            np.testing.assert_array_equal(t["obs"][:, 7:10], commands)
            np.testing.assert_array_equal(t["obs_clamped"], np.clip(t["obs"], -5, 5))
            np.testing.assert_allclose(t["rew"], rew, atol=3e-3)
            assert np.array_equal(t["progress"], progress) and np.array_equal(t["reset"], reset)
            assert np.array_equal(t["timeout"], ((progress >= c.max_episode_length - 1) & (reset != 0)).astype(np.int64))
            np.testing.assert_array_equal(t["actions"], a)
            # the task math at float32 precision on the kernel's OWN state
            eef_k = houndarm_eef(art, rk, dk, c.eef_body)
            np.testing.assert_allclose(t["obs"][:, :3], eef_k[:, :3], rtol=1e-5, atol=2e-6)
            dist_k = np.linalg.norm(eef_k[:, :3] - commands, axis=1)
            sure = np.abs(dist_k - 0.02) > 1e-4
            rew_k = np.maximum((1 - np.tanh(10 * dist_k)) * 0.1 + (1 - np.tanh(10 * np.linalg.norm(eef_k[:, 7:], axis=1))) * (dist_k < 0.02) * 0.1, 0.0)
            np.testing.assert_allclose(t["rew"][sure], rew_k[sure], rtol=1e-4, atol=2e-6)
            # keep the oracle on the kernel's trajectory (errors must not accumulate across steps of a chaotic closed loop)
            r64, d64 = rk.astype(np.float64), dk.astype(np.float64)
            resets += len(ids)
            timeouts += int(t["timeout"].sum())
    finally:
        be.close()
        jm.close()
    assert resets > 0 and timeouts > 0, "test must exercise resets and time-outs"
    assert checked_torques > steps * n * 3 // 4, "too few non-singular samples for the torque comparison"
    td = np.array(torque_dev)
    assert np.median(td) < 1e-5 and np.quantile(td, 0.9) < 1e-4 and td.max() < 1e-2, \
        f"OSC torque deviation (relative to the env's largest torque): median {np.median(td):.2e}, q90 {np.quantile(td, 0.9):.2e}, max {td.max():.2e}"


# ------------------------------------------------------------------------------------------------
# coarse heightfield bound (contact early-out on rough terrain)
# ------------------------------------------------------------------------------------------------
def hf_coarse_case(robot="useful_hound", n=16, seed=5, slope_scale=1.0):
    """Robots dropped from different heights / tilts onto a generated rough heightfield (steps, slopes, stairs)."""
    from isaacgymenv_b200.terrain import Terrain

    art = load_robot(robot)
    rng = np.random.default_rng(seed)
    tcfg = dict(terrainType="trimesh", curriculum=True, mapLength=8.0, mapWidth=8.0, numLevels=3, numTerrains=4,
                terrainProportions=[0.1, 0.1, 0.35, 0.25, 0.2], slopeTreshold=0.5)
    tr = Terrain(tcfg, n, seed=3)
    hf_t = _abi.Heightfield(rows=tr.tot_rows, cols=tr.tot_cols, horizontal_scale=tr.horizontal_scale, vertical_scale=tr.vertical_scale * slope_scale,
                            origin_x=-tr.border_size, origin_y=-tr.border_size, friction=1.0, restitution=0.0)
    root, dof = standing_state(art, n, rng, 0.45)
    origins = tr.env_origins.reshape(-1, 3).astype(np.float32)
    root[:, :3] += origins[np.arange(n) % len(origins)] * np.array([1, 1, slope_scale], np.float32)
    root[:, :2] += rng.uniform(-3, 3, (n, 2)).astype(np.float32)
    root[n // 2:, 3:7] = rng.normal(size=(n - n // 2, 4))           # half of them tumbling: links at arbitrary attitudes
    root[:, 3:7] /= np.linalg.norm(root[:, 3:7], axis=1, keepdims=True)
    root[-2:, :2] = [-40.0, 500.0]                                  # far outside the field: clamped to the border samples
    root[-2:, 2] = 0.3
    return art, hf_t, tr.heightsamples, root, dof


def check_hf_coarse_identical(simulate, robot="useful_hound", steps=25, n=16):
    """The coarse-bound early-out must never change a result: B2G_NO_HFC=1 (exhaustive candidate tests) vs default, bit for bit.
    `simulate(art, sp, props, hf_t, samples, root, dof, steps)` -> (root, dof, contact) after `steps` sim steps."""
    art, hf_t, samples, root, dof = hf_coarse_case(robot, n=n)
    sp = terrain_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 80.0, 2.0)
    outs = []
    old = os.environ.get("B2G_NO_HFC")
    try:
        for flag in ("1", "0"):
            os.environ["B2G_NO_HFC"] = flag
            outs.append(simulate(art, sp, props, hf_t, samples, root.copy(), dof.copy(), steps))
    finally:
        if old is None:
            os.environ.pop("B2G_NO_HFC", None)
        else:
            os.environ["B2G_NO_HFC"] = old
    for a, b in zip(*outs):
        assert np.array_equal(a, b), "coarse heightfield bound changed a result"
    assert np.abs(outs[0][2]).max() > 1.0, "the case must be in contact"
    return outs[1]
