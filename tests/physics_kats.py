"""Analytic known-answer tests and solver-convergence measurements for the restated dynamics (the
physics behind `gym.simulate`, reference call site tasks/base/vec_task.py:379-382).

Isaac Gym / PhysX is a closed binary, so the dynamics cannot be pinned to reference output.  What CAN
be pinned is physics with a closed-form answer, and the production solver against a converged solve of its own contact
model.  Every check takes a backend factory (tests/backends.py): the GPU tests run it on libb200gym.so through the C ABI,
the CPU tests on the host lane emulator -- the same kernel code either way.

  check_cartpole_closed_form     forward dynamics of the cart-pole against the textbook Lagrangian equations
  check_torque_free_precession   free axisymmetric body: precession rate (I3 - I1) / I1 * w3, momentum and energy conserved
  check_block_on_slope           Coulomb friction either side of tan(theta) = mu, on the plane and on a heightfield
  check_resting_force            a standing quadruped carries m g, split over the feet as statics says
  solver_deviation               production solver vs the converged reference over a random-action rollout
"""
from __future__ import annotations

import numpy as np

from isaacgymenv_b200 import _abi
from isaacgymenv_b200.model.urdf import Articulation
from oracle import dyn_oracle as O
from tests import kernel_checks as kc

G = 9.81


# ---------------------------------------------------------------------------------------------------------------------
# synthetic one-body robot: a box (root link, 8 corner contact spheres) carrying a light rotor on a revolute joint about z
# through the box centre (the kernels' model class is a root plus >= 1 chain)
# ---------------------------------------------------------------------------------------------------------------------
def box_robot(mass=2.0, half=(0.1, 0.1, 0.05), radius=0.01, inertia=None, rotor_mass=1e-3, rotor_inertia=(1e-6, 1e-6, 1e-6)):
    hx, hy, hz = half
    if inertia is None:
        inertia = (mass / 3.0 * (hy * hy + hz * hz), mass / 3.0 * (hx * hx + hz * hz), mass / 3.0 * (hx * hx + hy * hy))
    corners = np.array([[sx * hx, sy * hy, sz * hz] for sz in (-1, 1) for sx in (-1, 1) for sy in (-1, 1)], np.float64)
    inf = np.inf
    return Articulation(
        name="kat_box", fixed_base=False, body_names=["box", "rotor"], body_link=np.array([0, 1]), body_pos=np.zeros((2, 3)),
        body_quat=np.tile([0.0, 0, 0, 1], (2, 1)), dof_names=["spin"], link_names=["box", "rotor"], link_parent=np.array([-1, 0]),
        joint_type=np.array([_abi.JOINT_REVOLUTE if hasattr(_abi, "JOINT_REVOLUTE") else 0]), joint_pos=np.zeros((1, 3)),
        joint_quat=np.array([[0.0, 0, 0, 1]]), joint_axis=np.array([[0.0, 0, 1]]), mass=np.array([mass, rotor_mass]), com=np.zeros((2, 3)),
        inertia=np.stack([np.diag(inertia), np.diag(rotor_inertia)]), lower=np.array([-inf]), upper=np.array([inf]),
        has_limits=np.array([False]), effort=np.array([0.0]), velocity=np.array([0.0]), damping=np.array([0.0]), friction=np.array([0.0]),
        armature=np.array([0.0]), chain_start=np.array([0]), chain_len=np.array([1]), cp_link=np.zeros(8, np.int64),
        cp_body=np.zeros(8, np.int64), cp_pos=corners, cp_radius=np.full(8, radius))


def _params(dt, substeps=1, npos=4, nvel=1, gravity=(0, 0, -G), ground=True, mu=1.0):
    sp = kc.flat_params(dt=dt, substeps=substeps, npos=npos, nvel=nvel, ground=ground)
    sp.plane_static_friction = sp.plane_dynamic_friction = mu
    for i in range(3):
        sp.gravity[i] = gravity[i]
    return sp


def _quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# ---------------------------------------------------------------------------------------------------------------------
def check_cartpole_closed_form(make_backend, n=256, seed=0):
    """Cart (mass M) on a rail along y, pole (mass m, centre of mass l above the pivot, inertia I about it) hinged about x;
    generalized forces (F, tau).  Lagrange:  (M + m) y'' - m l cos(th) th'' + m l sin(th) th'^2 = F
                                             (I + m l^2) th'' - m l cos(th) y'' - m g l sin(th) = tau
    The kernel's forward-dynamics probe (articulated-body algorithm) must give the same accelerations."""
    art = kc.load_robot("cartpole")
    sp = kc.cartpole_params()
    props = kc.cartpole_props(art)
    M, m = float(art.mass[1]), float(art.mass[2])
    l = float(art.com[2][2])
    I = float(art.inertia[2][0][0])
    assert art.joint_axis[0].tolist() == [0.0, 1.0, 0.0] and art.joint_axis[1].tolist() == [1.0, 0.0, 0.0]
    rng = np.random.default_rng(seed)
    dof = np.zeros((n, 2, 2), np.float32)
    dof[:, 0, 0] = rng.uniform(-2, 2, n)
    dof[:, 1, 0] = rng.uniform(-np.pi, np.pi, n)
    dof[:, :, 1] = rng.uniform(-3, 3, (n, 2))
    tau = rng.uniform(-50, 50, (n, 2)).astype(np.float32)
    root = np.zeros((n, 13), np.float32)
    root[:, 6] = 1
    be = make_backend(art, sp, props, n)
    try:
        be.set_state(root, dof)
        qdd, _ = be.forward_dynamics(tau)
    finally:
        be.close()
    th, thd = dof[:, 1, 0].astype(np.float64), dof[:, 1, 1].astype(np.float64)
    g = -float(sp.gravity[2])
    a11, a12, a22 = M + m, -m * l * np.cos(th), I + m * l * l
    b1 = tau[:, 0] - m * l * np.sin(th) * thd ** 2
    b2 = tau[:, 1] + m * g * l * np.sin(th)
    det = a11 * a22 - a12 * a12
    ydd = (a22 * b1 - a12 * b2) / det
    thdd = (a11 * b2 - a12 * b1) / det
    np.testing.assert_allclose(qdd[:, 0], ydd, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(qdd[:, 1], thdd, rtol=1e-3, atol=1e-3)
    return float(np.abs(qdd[:, 0] - ydd).max()), float(np.abs(qdd[:, 1] - thdd).max())


# ---------------------------------------------------------------------------------------------------------------------
def check_torque_free_precession(make_backend, dt=0.001, t_end=2.0, n=4):
    """A free axisymmetric body (I1 = I2 != I3) in zero gravity: in the body frame w3 is constant and (w1, w2) rotate at
    Omega = (I3 - I1) / I1 * w3 (Euler's equations, closed form).  World angular momentum and kinetic energy are conserved.
    Environments: prolate, oblate, spherical (Omega = 0) bodies and different spin rates."""
    cases = [  # (I1, I3, w_transverse, w3)
        (0.02, 0.05, 1.0, 6.0), (0.05, 0.02, 1.5, 5.0), (0.03, 0.03, 1.0, 4.0), (0.02, 0.05, 0.3, -8.0)]
    results = []
    for I1, I3, wt, w3 in cases:
        ri = (1e-5, 1e-5, 2e-5)
        art = box_robot(mass=2.0, inertia=(I1, I1, I3), rotor_mass=1e-3, rotor_inertia=ri)
        sp = _params(dt, gravity=(0, 0, 0), ground=False)
        props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 200.0, 2.0)     # the rotor is held: one rigid body
        be = make_backend(art, sp, props, n)
        try:
            root = np.zeros((n, 13), np.float32)
            root[:, 2] = 1.0
            q0 = np.array([0.3, -0.2, 0.1, 0.9]); q0 /= np.linalg.norm(q0)
            root[:, 3:7] = q0
            R0 = _quat_to_mat(q0)
            wb = np.array([wt, 0.0, w3])
            root[:, 10:13] = R0 @ wb
            dof = np.zeros((n, 1, 2), np.float32)
            be.set_state(root, dof)
            steps = int(round(t_end / dt))
            J1, J3 = I1 + ri[0], I3 + ri[2]
            Ib = np.diag([J1, J1, J3])
            L0 = R0 @ Ib @ wb
            E0 = 0.5 * wb @ Ib @ wb
            tgt = np.zeros((n, 1), np.float32)
            for _ in range(steps):
                be.simulate(tgt, tgt)
            r, d = be.get_state()
        finally:
            be.close()
        R = _quat_to_mat(r[0, 3:7].astype(np.float64))
        w_body = R.T @ r[0, 10:13].astype(np.float64)
        Om = (J3 - J1) / J1 * w3
        ph = Om * t_end
        expect = np.array([wt * np.cos(ph), wt * np.sin(ph), w3])
        L = R @ Ib @ w_body
        E = 0.5 * w_body @ Ib @ w_body
        err = np.abs(w_body - expect).max()
        results.append(dict(I1=I1, I3=I3, w3=w3, omega_precession=Om, w_body=w_body.tolist(), expect=expect.tolist(), err=float(err),
                            dL=float(np.abs(L - L0).max() / np.linalg.norm(L0)), dE=float(abs(E - E0) / E0), rotor=float(np.abs(d[0, 0]).max())))
        # first-order integrator: phase error O(dt * w^2 * t); at dt = 1 ms the body-frame rate stays within 3 % of |w|
        # explicit first-order integrator: the transverse amplitude grows by (1 + (Omega dt)^2 / 2) per step; the precession PHASE is what is pinned
        assert err < 0.03 * np.linalg.norm(wb) * max(dt / 0.001, 1.0), (I1, I3, w3, w_body, expect)
        assert abs(w_body[2] - w3) < 5e-3 * abs(w3)
        assert results[-1]["dL"] < 1e-2 and results[-1]["dE"] < 2e-2
    return results


# ---------------------------------------------------------------------------------------------------------------------
def check_block_on_slope(make_backend, heightfield=False, t_end=1.0, dt=0.005):
    """A box on an incline of tan(theta) = 0.5 with friction coefficients either side of it: mu < tan(theta) slides with
    a = g (sin(theta) - mu cos(theta)), mu > tan(theta) stays put.  Plane: the incline is a tilted gravity vector; heightfield:
    a real sloped field under real gravity.  mu_effective = (mu_ground + mu_shape) / 2 (PhysX's default combine mode)."""
    theta = np.arctan(0.5)
    mus = np.array([0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.8, 1.0])
    mu_ground = 0.6
    n = len(mus)
    art = box_robot()
    half_z, rad = 0.05, 0.01
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 200.0, 2.0)
    if heightfield:
        sp = _params(dt, ground=False, mu=mu_ground)
        rows, cols, hs, vs = 200, 60, 0.1, 0.005
        samples = (np.arange(rows)[:, None] * (hs * np.tan(theta) / vs) * np.ones((1, cols))).round().astype(np.int16)
        hf_t = _abi.Heightfield(rows=rows, cols=cols, horizontal_scale=hs, vertical_scale=vs, origin_x=0.0, origin_y=0.0, friction=mu_ground,
                                restitution=0.0)
        normal = np.array([-np.sin(theta), 0.0, np.cos(theta)])
        down = np.array([-np.cos(theta), 0.0, -np.sin(theta)])          # downhill unit vector
        x0 = 15.0
        centre = np.array([x0, 3.0, x0 * np.tan(theta)]) + normal * (half_z + rad)
        quat = np.array([0.0, np.sin(-theta / 2), 0.0, np.cos(-theta / 2)])
    else:
        sp = _params(dt, gravity=(G * np.sin(theta), 0.0, -G * np.cos(theta)), mu=mu_ground)
        normal, down = np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0])
        centre = np.array([0.0, 0.0, half_z + rad])
        quat = np.array([0.0, 0.0, 0.0, 1.0])
    be = make_backend(art, sp, props, n)
    try:
        if heightfield:
            be.add_heightfield(hf_t, samples)
        be.set_friction(2.0 * mus - mu_ground)
        root = np.zeros((n, 13), np.float32)
        root[:, :3] = centre
        root[:, 1] += 0.5 * np.arange(n) if heightfield else 0.0
        root[:, 3:7] = quat
        dof = np.zeros((n, 1, 2), np.float32)
        be.set_state(root, dof)
        tgt = np.zeros((n, 1), np.float32)
        steps = int(round(t_end / dt))
        fsum = np.zeros(n)
        for _ in range(steps):
            _, c = be.simulate(tgt, tgt)
            fsum += c[:, 0, :] @ normal
        r, _ = be.get_state()
    finally:
        be.close()
    v_down = r[:, 7:10].astype(np.float64) @ down
    travelled = (r[:, :3].astype(np.float64) - root[:, :3]) @ down
    a_expect = np.where(mus < np.tan(theta), G * (np.sin(theta) - mus * np.cos(theta)), 0.0)
    slide = mus < np.tan(theta)
    out = dict(mu=mus.tolist(), v=v_down.tolist(), v_expect=(a_expect * t_end).tolist(), travelled=travelled.tolist(),
               normal_force_over_mg_cos=(fsum / steps / (art.total_mass * G * np.cos(theta))).tolist())
    np.testing.assert_allclose(v_down[slide], a_expect[slide] * t_end, rtol=5e-3, err_msg=str(out))
    np.testing.assert_allclose(travelled[slide], 0.5 * a_expect[slide] * t_end * (t_end + dt), rtol=1e-2, err_msg=str(out))   # semi-implicit Euler: x_k = a dt^2 k (k + 1) / 2
    assert np.all(np.abs(v_down[~slide]) < 2e-3), out              # sticking: no creep
    assert np.all(np.abs(travelled[~slide]) < 2e-3), out
    np.testing.assert_allclose(fsum / steps, art.total_mass * G * np.cos(theta), rtol=0.02, err_msg=str(out))
    height = (r[:, :3].astype(np.float64) - root[:, :3]) @ normal
    assert np.all(np.abs(height) < 5e-3), height                     # neither sinking nor popping out
    return out


# ---------------------------------------------------------------------------------------------------------------------
def check_resting_force(make_backend, robot="anymal", n=8, settle_s=1.5, slots=0):
    """A quadruped standing on its default pose: the feet carry m g in total, the left/right split is symmetric and the
    front/rear split follows the centre of mass (moment balance about the y axis)."""
    art = kc.load_robot(robot)
    sp = kc.flat_params(slots=slots)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    rng = np.random.default_rng(0)
    root, dof = kc.standing_state(art, n, rng, 0.62 if "anymal" in robot else 0.5)
    root[:, 3:7] = [0, 0, 0, 1]
    root[:, 7:13] = 0
    dof[:, :, 0] = kc.default_pose(art)
    dof[:, :, 1] = 0
    tgt = np.tile(kc.default_pose(art), (n, 1)).astype(np.float32)
    be = make_backend(art, sp, props, n)
    try:
        be.set_state(root, dof)
        steps = int(settle_s / sp.dt)
        for _ in range(steps):
            _, c = be.simulate(tgt, np.zeros_like(tgt))
        r, d = be.get_state()
    finally:
        be.close()
    mg = art.total_mass * G
    fz = c[:, :, 2].astype(np.float64)
    total = fz.sum(1)
    loaded = [i for i in range(art.num_bodies) if fz[0, i] > 0.02 * mg]
    out = dict(robot=robot, residual_speed=float(np.abs(r[:, 7:13]).max()), residual_joint_speed=float(np.abs(d[:, :, 1]).max()),
               total_over_mg=float(total.mean() / mg), total_over_mg_range=[float(total.min() / mg), float(total.max() / mg)],
               loaded_bodies=[art.body_names[i] for i in loaded], base_height=float(r[:, 2].mean()))
    assert np.all(np.abs(r[:, 7:13]) < 2e-2), out       # at rest
    np.testing.assert_allclose(total, mg, rtol=0.01, err_msg=str(out))
    # the load sits on the legs (the 44 kg Hound sags onto its rear thighs under the reference's own Kp = 85: cfg/task/Hound.yaml:19), the
    # four legs share it
    assert loaded and all(any(k in art.body_names[i] for k in ("SHANK", "FOOT", "foot", "calf", "thigh")) for i in loaded), out
    legs = sorted({art.body_names[i][:2] for i in loaded})
    assert len(legs) == 4, legs
    share = np.stack([fz[:, [i for i in loaded if art.body_names[i][:2] == leg]].sum(1) for leg in legs], axis=1) / mg
    assert np.all(share > 0.12) and np.all(share < 0.40), share[0]
    # no horizontal net force at rest
    assert np.all(np.abs(c[:, :, :2].sum(1)) < 0.01 * mg)
    out["share"] = share[0].tolist()
    return out


# ---------------------------------------------------------------------------------------------------------------------
def solver_deviation(make_backend, robot="anymal", n=4096, steps=200, sample_every=20, n_ref=1024, seed=0, hard_limits=False, max_iter=500, slots=0):
    """Random-action rollout of the flat task's physics (implicit PD position targets 0.5 a + q0, tasks/anymal.py:226-229) on the
    backend; at every `sample_every`-th step the pre-step states of `n_ref` environments are also advanced ONE policy step by the
    converged reference solver (oracle: every contact, sequential Gauss-Seidel to convergence, float64) and by the production
    algorithm's oracle.  Returns the error table (production on the device vs converged reference) and the contact-cap statistics."""
    art = kc.load_robot(robot)
    sp = kc.flat_params(slots=slots)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    model = _abi.pack_model(art)
    nd = art.num_dofs
    rng = np.random.default_rng(seed)
    q0 = kc.default_pose(art).astype(np.float32)
    root, dof = kc.standing_state(art, n, rng, 0.62 if "anymal" in robot else 0.5)
    root[:, 3:7] = [0, 0, 0, 1]
    dof[:, :, 0] = q0
    be = make_backend(art, sp, props, n)
    rows = []
    mg = art.total_mass * G
    try:
        be.set_state(root, dof)
        be.contact_stats(reset=True)
        zero = np.zeros((n, nd), np.float32)
        for t in range(steps):
            a = (2 * rng.random((n, nd), dtype=np.float32) - 1)
            tgt = 0.5 * a + q0[None]
            sample = (t % sample_every) == sample_every - 1
            if sample:
                r_pre, d_pre = be.get_state()
            f_dev, c_dev = be.simulate(tgt, zero)
            # envs that fall over are put back on their feet so the rollout keeps visiting walking, stumbling and lying states
            if sample:
                r_dev, d_dev = be.get_state()
                idx = rng.choice(n, size=min(n_ref, n), replace=False)
                rr, dd = r_pre[idx].astype(np.float64), d_pre[idx].astype(np.float64)
                f_ref, c_ref, info = O.simulate_ref(model, sp, props, rr, dd, tgt[idx].astype(np.float64), zero[idx].astype(np.float64),
                                                    hard_limits=hard_limits, max_iter=max_iter, tol=1e-10)
                rp, dp = r_pre[idx].astype(np.float64), d_pre[idx].astype(np.float64)
                f_orc, c_orc = O.simulate(model, sp, props, rp, dp, tgt[idx].astype(np.float64), zero[idx].astype(np.float64))
                rows.append(dict(
                    dv=np.linalg.norm(r_dev[idx, 7:10] - rr[:, 7:10], axis=1), dw=np.linalg.norm(r_dev[idx, 10:13] - rr[:, 10:13], axis=1),
                    dx=np.linalg.norm(r_dev[idx, :3] - rr[:, :3], axis=1), dqd=np.abs(d_dev[idx, :, 1] - dd[:, :, 1]).max(1),
                    dq=np.abs(d_dev[idx, :, 0] - dd[:, :, 0]).max(1), dF=np.linalg.norm(c_dev[idx].sum(1) - c_ref.sum(1), axis=1) / mg,
                    F=np.linalg.norm(c_ref.sum(1), axis=1) / mg, ncon=info[:, 0], it_pos=info[:, 1], it_vel=info[:, 2], capped=info[:, 3],
                    term_dev=(np.linalg.norm(c_dev[idx], axis=2) > 1.0), term_ref=(np.linalg.norm(c_ref, axis=2) > 1.0),
                    port_dv=np.linalg.norm(r_dev[idx, 7:10] - rp[:, 7:10], axis=1), port_dqd=np.abs(d_dev[idx, :, 1] - dp[:, :, 1]).max(1),
                    z=r_pre[idx, 2]))
            if t % 50 == 49:
                r_now, d_now = be.get_state()
                fallen = r_now[:, 2] < 0.25
                k = int(fallen.sum())
                if k:
                    r_new, d_new = kc.standing_state(art, k, rng, 0.62 if "anymal" in robot else 0.5)
                    r_new[:, :2] = r_now[fallen, :2]
                    r_now[fallen], d_now[fallen] = r_new, d_new
                    d_now[fallen, :, 0] = q0
                    be.set_state(r_now, d_now)
        stats = be.contact_stats()
    finally:
        be.close()
    cat = {k: np.concatenate([r[k] for r in rows]) for k in rows[0]}
    # a small fraction of states has no fixed point the sweeps can reach (stick / slip two-cycles at mu = 1 with strong normal-tangential
    # coupling): the reference flags them (hit max_iter) and they are left out of the error statistics, their share is reported
    conv = cat["capped"] == 0
    not_converged = float(1.0 - conv.mean())
    cat = {k: v[conv] for k, v in cat.items()}

    def q(x):
        return dict(median=float(np.median(x)), p90=float(np.percentile(x, 90)), p99=float(np.percentile(x, 99)), max=float(np.max(x)))

    table = dict(
        robot=robot, envs=n, steps=steps, samples=int(len(cat["dv"])), hard_limits=bool(hard_limits), contact_slots=int(slots or 4),
        root_lin_vel_err_m_s=q(cat["dv"]), root_ang_vel_err_rad_s=q(cat["dw"]), root_pos_err_m=q(cat["dx"]), joint_vel_err_rad_s=q(cat["dqd"]),
        joint_pos_err_rad=q(cat["dq"]), net_contact_force_err_over_mg=q(cat["dF"]), net_contact_force_over_mg=q(cat["F"]),
        contact_flag_agreement=float((cat["term_dev"] == cat["term_ref"]).mean()),
        reference=dict(contacts_mean=float(cat["ncon"].mean()), contacts_max=int(cat["ncon"].max()), sweeps_pos_median=float(np.median(cat["it_pos"])),
                       sweeps_pos_max=int(cat["it_pos"].max()), sweeps_vel_median=float(np.median(cat["it_vel"])), not_converged_fraction=not_converged),
        device_vs_production_oracle=dict(root_lin_vel_err_m_s=q(cat["port_dv"]), joint_vel_err_rad_s=q(cat["port_dqd"])),
        base_height_m=q(cat["z"]),
        contact_cap=dict(active_contacts=stats[0], dropped_candidates=stats[1], env_substeps_with_drop=stats[2], env_substeps=stats[3],
                         dropped_fraction=stats[1] / max(stats[0] + stats[1], 1), env_substeps_with_drop_fraction=stats[2] / max(stats[3], 1)))
    return table


def production_scheme_convergence(robot="anymal", n=256, seed=0, budgets=((4, 1), (8, 2), (16, 4), (64, 16), (256, 64))):
    """Oracle only: the production algorithm (contact slots capped per chain, Gauss-Seidel along a chain / Jacobi across chains, block
    solve for sticking + scalar proximal step for sliding contacts) run with growing sweep budgets against the converged reference
    (one contact list, sequential Gauss-Seidel, exact per-contact Coulomb solve).  Different iterations, same fixed point."""
    art = kc.load_robot(robot)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    model = _abi.pack_model(art)
    sp = kc.flat_params()
    nd = art.num_dofs
    rng = np.random.default_rng(seed)
    root, dof = kc.standing_state(art, n, rng, 0.62 if "anymal" in robot else 0.5)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    q0 = kc.default_pose(art)
    zero = np.zeros((n, nd))
    for _ in range(60):
        O.simulate(model, sp, props, root, dof, 0.5 * (2 * rng.random((n, nd)) - 1) + q0, zero)
    tgt = 0.5 * (2 * rng.random((n, nd)) - 1) + q0
    r2, d2 = root.copy(), dof.copy()
    _, c2, info = O.simulate_ref(model, sp, props, r2, d2, tgt, zero, max_iter=2000, tol=1e-9)
    ok = info[:, 3] == 0
    out = []
    for npos, nvel in budgets:
        spx = kc.flat_params(npos=npos, nvel=nvel)
        r1, d1 = root.copy(), dof.copy()
        _, c1 = O.simulate(model, spx, props, r1, d1, tgt, zero)
        out.append(dict(npos=npos, nvel=nvel, root_lin_vel_err_median=float(np.median(np.linalg.norm(r1[ok, 7:10] - r2[ok, 7:10], axis=1))),
                        joint_vel_err_median=float(np.median(np.abs(d1[ok, :, 1] - d2[ok, :, 1]).max(1))),
                        net_force_err_over_mg_median=float(np.median(np.linalg.norm(c1[ok].sum(1) - c2[ok].sum(1), axis=1)) / (art.total_mass * G))))
    return dict(converged_fraction=float(ok.mean()), contacts_mean=float(info[:, 0].mean()), rows=out)


def check_hard_joint_limits_reference(robot="anymal", n=32, seed=3):
    """Oracle only: with flags & 1 the reference treats joint limits as hard rows -- a joint driven against its stop ends the step
    on the stop (not beyond it), where the production path's implicit spring lets it sink in by tau / k."""
    art = kc.load_robot(robot)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    model = _abi.pack_model(art)
    sp = kc.flat_params(ground=False)
    nd = art.num_dofs
    rng = np.random.default_rng(seed)
    root, dof = kc.random_flying_state(art, n, rng, scale_qd=0.0)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    lim = np.array([d for d in range(nd) if art.has_limits[d] and art.upper[d] - art.lower[d] < 3.0])
    assert len(lim) > 0
    dof[:, :, 0] = kc.default_pose(art)
    tgt = np.tile(kc.default_pose(art), (n, 1)).astype(np.float64)
    tgt[:, lim] = art.upper[lim] + 1.0                      # ask for a pose one radian beyond the upper stop
    zero = np.zeros((n, nd))
    rs, ds = root.copy(), dof.copy()
    for _ in range(40):
        O.simulate_ref(model, sp, props, root, dof, tgt, zero, hard_limits=True, max_iter=500, tol=1e-10)
        O.simulate(model, sp, props, rs, ds, tgt, zero)
    over_hard = (dof[:, lim, 0] - art.upper[lim]).max()
    over_soft = (ds[:, lim, 0] - art.upper[lim]).max()
    assert over_hard < 1e-6, over_hard
    assert np.abs(dof[:, lim, 0] - art.upper[lim]).max() < 1e-3          # resting ON the stop
    assert 1e-3 < over_soft < 0.1, over_soft                             # the spring sinks in by about Kp * 1 rad / k_limit
    return float(over_hard), float(over_soft)
