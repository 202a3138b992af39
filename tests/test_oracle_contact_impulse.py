"""One contact, derived independently.  A robot comes down on ONE foot; what a velocity-level contact solver must do is then a small closed
form in operational space,

    W = J M^-1 J^T,    v_free = v + h a(q, v, tau),    lambda = W^-1 (target - J v_free)   (kept if lambda_n >= 0 and inside the cone),

with every ingredient taken from somewhere else than the simulator: M and a from the Euler-Lagrange model of tests/test_oracle_lagrange.py
(AD of a kinematics-only Lagrangian), the contact-point Jacobian J by forward-mode AD of that model's forward kinematics.  The oracle reaches
the same numbers through articulated-body inertias and impulse propagation down the chains (oracle/dyn/oracle_dyn_impl.h); this test pins
its contact Jacobian, its Delassus operator, the two-stage scheme it takes from PhysX (position iterations with the penetration bias give the
velocity the positions are integrated with, velocity iterations without the push-out bias give the velocity that is kept) and the force it
reports, on the frictionless normal row (production solver: exact after one sweep) and on a sticking contact with friction (converged
reference solver).  Sliding friction is pinned by the block-on-a-slope known answer (tests/physics_kats.py)."""
import numpy as np
import pytest
import torch

from isaacgymenv_b200 import _abi
from oracle import dyn_oracle as O
from tests.kernel_checks import default_pose, flat_params, load_robot
from tests.test_oracle_lagrange import EXACT_FRAMES, Lagrange, _quat_to_mat


def _one_foot_case(robot, gap0, vel, seed, h, spin=0.3, slope=0.0, origin_x=0.0):
    """State with exactly one contact candidate (the lowest contact sphere at gap `gap0`, the next one far outside the contact offset),
    the Euler-Lagrange model, and the operational-space quantities of that contact."""
    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(dt=h, substeps=1, ground=True)
    rng = np.random.default_rng(seed)
    nd = art.num_dofs
    root = np.zeros(13)
    root[3:7] = [0.12, -0.08, 0.05, 1.0]
    root[3:7] /= np.linalg.norm(root[3:7])
    root[7:10] = vel
    root[10:13] = rng.normal(size=3) * spin
    dof = np.zeros((nd, 2))
    dof[:, 0] = default_pose(art) + rng.uniform(-0.1, 0.1, nd)
    dof[:, 1] = rng.normal(size=nd) * spin
    tau = rng.normal(size=nd) * np.where(np.arange(nd) < 12, 2.0, 0.02)      # legs / arm: well inside the effort and joint-velocity limits
    L = Lagrange(art, m, [sp.gravity[0], sp.gravity[1], sp.gravity[2]])
    ncp = m.n_cpts
    cp_link = [int(m.cp_link[i]) for i in range(ncp)]
    cp_pos = torch.tensor([list(m.cp_pos[i]) for i in range(ncp)], dtype=torch.float64)
    cp_rad = np.array([m.cp_radius[i] for i in range(ncp)], dtype=np.float64)
    rr = _quat_to_mat(torch.tensor(root[3:7]))
    x0 = torch.cat([torch.zeros(6, dtype=torch.float64), torch.tensor(dof[:, 0])])

    def frames(x, rp):
        com, rot = L.fk(x, rp, rr)
        return com - torch.einsum("lij,lj->li", rot, L.com), rot

    # ground: z = slope * (x - origin_x) (slope 0: the plane); gap of a sphere as the simulator defines it, (z - ground(x, y)) n_z - radius
    n = np.array([-slope, 0.0, 1.0]) / np.sqrt(1.0 + slope * slope)
    org, rot = frames(x0, torch.tensor(root[:3]))
    ctr = np.array([(org[cp_link[i]] + rot[cp_link[i]] @ cp_pos[i]).numpy() for i in range(ncp)])
    gaps = (ctr[:, 2] - slope * (ctr[:, 0] - origin_x)) * n[2] - cp_rad
    order = np.argsort(gaps)
    root[2] += (gap0 - gaps[order[0]]) / n[2]
    assert gaps[order[1]] - gaps[order[0]] + gap0 > sp.contact_offset + 0.01      # the second-lowest sphere is out of range
    ic = int(order[0])
    rp = torch.tensor(root[:3])
    off = -cp_rad[ic] * n                                  # contact point on the sphere's surface, as a point FIXED on the link
    Jc = torch.func.jacfwd(lambda x: (lambda o, r: o[cp_link[ic]] + r[cp_link[ic]] @ cp_pos[ic])(*frames(x, rp)))(x0).numpy()
    dR = torch.func.jacfwd(lambda x: frames(x, rp)[1][cp_link[ic]])(x0).numpy()
    Rl = frames(x0, rp)[1][cp_link[ic]].numpy()
    J = np.zeros((3, 6 + nd))
    for k in range(6 + nd):
        Wk = dR[:, :, k] @ Rl.T
        J[:, k] = Jc[:, k] + np.cross([Wk[2, 1], Wk[0, 2], Wk[1, 0]], off)
    xd = np.concatenate([root[7:10], root[10:13], dof[:, 1]])
    xdd, _, M = L.accelerations(root, dof, tau)
    return dict(art=art, m=m, sp=sp, root=root, dof=dof, tau=tau, n=n, J=J, Minv=np.linalg.inv(M), v_free=xd + h * xdd, body=int(m.cp_body[ic]), nd=nd,
                tgt=min(-gap0 / h, float(sp.max_depenetration_velocity)))


def _compare(c, r, d, contact, v_pos, v_fin, lam_world, h, tol):
    got_v = np.concatenate([r[0][7:10], r[0][10:13], d[0][:, 1]])
    scale = max(1.0, np.abs(v_fin).max())
    assert np.abs(got_v - v_fin).max() < tol * scale, np.abs(got_v - v_fin).max()
    assert np.abs(d[0][:, 0] - (c["dof"][:, 0] + h * v_pos[6:])).max() < tol        # positions integrate with the position-stage velocity
    assert np.abs(r[0][:3] - (c["root"][:3] + h * v_pos[:3])).max() < tol
    f = contact[0][c["body"]]
    assert np.abs(f - lam_world / h).max() < tol * max(1.0, np.abs(lam_world / h).max()), (f, lam_world / h)
    assert np.abs(np.delete(contact[0], c["body"], axis=0)).max() == 0.0             # one contact, one body


@pytest.mark.parametrize("robot", ["hound", "anymal", "useful_hound"])
@pytest.mark.parametrize("gap0,vz", [(-0.003, -0.5), (0.004, -1.5), (0.004, 0.5)])
def test_frictionless_contact_equals_the_operational_space_solution(robot, gap0, vz):
    """Penetrating (push-out bias in the position stage, removed in the velocity stage), approaching inside the contact offset (speculative
    contact: the impulse only takes away the part of the approach that would penetrate) and separating (no impulse)."""
    h = 0.005
    c = _one_foot_case(robot, gap0, [0.3, -0.2, vz], seed=1, h=h)
    sp, J, Minv, v_free, tgt = c["sp"], c["J"], c["Minv"], c["v_free"], c["tgt"]
    sp.plane_dynamic_friction = sp.plane_static_friction = 0.0
    Jn = J[2]
    Wn = Jn @ Minv @ Jn
    lam_p = max(0.0, (tgt - Jn @ v_free) / Wn)
    v_pos = v_free + Minv @ Jn * lam_p
    lam_v = max(0.0, lam_p - (Jn @ v_pos - min(tgt, 0.0)) / Wn)
    v_fin = v_pos + Minv @ Jn * (lam_v - lam_p)
    assert (lam_p > 0) == (vz < 0)
    props = _abi.default_dof_props(c["art"], _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    r, d = c["root"][None].copy(), c["dof"][None].copy()
    _, contact = O.simulate(c["m"], sp, props, r, d, np.zeros((1, c["nd"])), c["tau"][None], friction=np.zeros(1, np.float32))
    _compare(c, r, d, contact, v_pos, v_fin, np.array([0.0, 0.0, lam_v]), h, 5e-8 if robot in EXACT_FRAMES else 2e-6)


@pytest.mark.parametrize("robot", ["hound", "anymal"])
def test_sticking_contact_equals_the_operational_space_solution(robot):
    """Friction 1, a slow sideways drift: the contact sticks, so the converged solver's impulse is the 3 x 3 block solve in both stages."""
    h = 0.005
    c = _one_foot_case(robot, -0.002, [0.02, -0.015, -0.8], seed=4, h=h, spin=0.03)
    sp, J, Minv, v_free, tgt = c["sp"], c["J"], c["Minv"], c["v_free"], c["tgt"]
    W = J @ Minv @ J.T
    lam_p = np.linalg.solve(W, np.array([0.0, 0.0, tgt]) - J @ v_free)
    v_pos = v_free + Minv @ J.T @ lam_p
    lam_v = lam_p + np.linalg.solve(W, np.array([0.0, 0.0, min(tgt, 0.0)]) - J @ v_pos)
    v_fin = v_pos + Minv @ J.T @ (lam_v - lam_p)
    for lam in (lam_p, lam_v):
        assert lam[2] > 0 and np.hypot(lam[0], lam[1]) < 0.9 * lam[2]          # inside the friction cone (mu = 1): sticking
    props = _abi.default_dof_props(c["art"], _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    r, d = c["root"][None].copy(), c["dof"][None].copy()
    _, contact, info = O.simulate_ref(c["m"], sp, props, r, d, np.zeros((1, c["nd"])), c["tau"][None], friction=np.ones(1, np.float32), tol=1e-13)
    assert info[0][0] == 1 and info[0][3] == 0                                   # one contact, converged
    _compare(c, r, d, contact, v_pos, v_fin, lam_v, h, 5e-7 if robot in EXACT_FRAMES else 5e-6)


@pytest.mark.parametrize("robot", ["hound", "anymal", "useful_hound"])
@pytest.mark.parametrize("gap0,vz", [(-0.003, -0.5), (0.004, -1.5)])
def test_kernel_code_frictionless_contact_equals_the_operational_space_solution(robot, gap0, vz):
    """The step kernels' own contact code (float32, host lane emulator: contact candidates, per-slot Delassus terms, lane-Jacobi sweeps with
    the root sum across chains, impulse propagation) on the same one-foot cases, straight against the operational-space solution."""
    from tests.backends import EmuBackend

    h = 0.005
    c = _one_foot_case(robot, gap0, [0.3, -0.2, vz], seed=1, h=h)
    sp, J, Minv, v_free, tgt = c["sp"], c["J"], c["Minv"], c["v_free"], c["tgt"]
    sp.plane_dynamic_friction = sp.plane_static_friction = 0.0
    Jn = J[2]
    Wn = Jn @ Minv @ Jn
    lam_p = max(0.0, (tgt - Jn @ v_free) / Wn)
    v_pos = v_free + Minv @ Jn * lam_p
    lam_v = max(0.0, lam_p - (Jn @ v_pos - min(tgt, 0.0)) / Wn)
    v_fin = v_pos + Minv @ Jn * (lam_v - lam_p)
    assert lam_p > 0
    props = _abi.default_dof_props(c["art"], _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    be = EmuBackend(c["art"], sp, props, 1)
    try:
        be.set_state(c["root"][None].astype(np.float32), c["dof"][None].astype(np.float32))
        be.set_friction(np.zeros(1, np.float32))
        _, contact = be.simulate(np.zeros((1, c["nd"]), np.float32), c["tau"][None].astype(np.float32))
        r, d = be.get_state()
    finally:
        be.close()
    got_v = np.concatenate([r[0][7:10], r[0][10:13], d[0][:, 1]]).astype(np.float64)
    scale = max(1.0, np.abs(v_fin).max())
    assert np.abs(got_v - v_fin).max() < 2e-4 * scale, np.abs(got_v - v_fin).max()
    fz = float(contact[0][c["body"]][2])
    assert abs(fz - lam_v / h) < 1e-3 * max(1.0, lam_v / h), (fz, lam_v / h)


@pytest.mark.parametrize("robot", ["hound", "anymal"])
def test_frictionless_contact_on_a_sloped_heightfield(robot):
    """The rough-terrain path (int16 height samples, triangle normal, gap along the vertical scaled by n_z): a field of constant slope 0.25,
    one foot penetrating it -- the impulse acts along the field's normal (-0.25, 0, 1) / |.|, through the same operational-space solution."""
    h, slope, ox = 0.005, 0.25, -5.0
    c = _one_foot_case(robot, -0.003, [0.3, -0.2, -0.5], seed=1, h=h, slope=slope, origin_x=ox)
    sp, J, Minv, v_free, tgt, n = c["sp"], c["J"], c["Minv"], c["v_free"], c["tgt"], c["n"]
    hs, vs, rows, cols = 0.1, 0.005, 101, 101
    samples = np.repeat((np.arange(rows) * round(slope * hs / vs)).astype(np.int16)[:, None], cols, axis=1)      # rows run along x
    hf = _abi.Heightfield(rows=rows, cols=cols, horizontal_scale=hs, vertical_scale=vs, origin_x=ox, origin_y=-5.0, friction=0.0, restitution=0.0)
    Jn = n @ J
    Wn = Jn @ Minv @ Jn
    lam_p = max(0.0, (tgt - Jn @ v_free) / Wn)
    v_pos = v_free + Minv @ Jn * lam_p
    lam_v = max(0.0, lam_p - (Jn @ v_pos - min(tgt, 0.0)) / Wn)
    v_fin = v_pos + Minv @ Jn * (lam_v - lam_p)
    assert lam_p > 0 and lam_v > 0
    props = _abi.default_dof_props(c["art"], _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    r, d = c["root"][None].copy(), c["dof"][None].copy()
    _, contact = O.simulate(c["m"], sp, props, r, d, np.zeros((1, c["nd"])), c["tau"][None], heightfield=hf, hf_samples=samples, friction=np.zeros(1, np.float32))
    _compare(c, r, d, contact, v_pos, v_fin, n * lam_v, h, 5e-7 if robot in EXACT_FRAMES else 2e-6)


@pytest.mark.parametrize("robot", ["hound", "anymal"])
def test_two_frictionless_contacts_couple_through_the_base(robot):
    """Two feet of different legs touch at once: their normal rows couple through the floating base (W is a full 2 x 2 matrix), and the
    converged solver must land on the solution of the two-row complementarity problem.  The production scheme updates the two chains from the
    same velocities (Jacobi across chains); the coupling is weak -- W_01 / W_00 ~ 5e-5: the base is heavy, the legs are light, which is what
    makes the lane-Jacobi split cheap -- so 4 + 1 sweeps already sit within 5e-5 of the exact solution (measured and asserted < 1e-3)."""
    h = 0.005
    art = load_robot(robot)
    m = _abi.pack_model(art)
    sp = flat_params(dt=h, substeps=1, ground=True)
    sp.plane_dynamic_friction = sp.plane_static_friction = 0.0
    rng = np.random.default_rng(9)
    nd = art.num_dofs
    roll = 0.2
    root = np.zeros(13)
    root[3:7] = [np.sin(roll / 2), 0.0, 0.0, np.cos(roll / 2)]
    root[7:10] = [0.1, 0.0, -0.6]
    root[10:13] = rng.normal(size=3) * 0.1
    dof = np.zeros((nd, 2))
    dof[:, 0] = default_pose(art)
    dof[:, 1] = rng.normal(size=nd) * 0.1
    tau = rng.normal(size=nd) * 1.0
    L = Lagrange(art, m, [sp.gravity[0], sp.gravity[1], sp.gravity[2]])
    ncp = m.n_cpts
    cp_link = [int(m.cp_link[i]) for i in range(ncp)]
    cp_pos = torch.tensor([list(m.cp_pos[i]) for i in range(ncp)], dtype=torch.float64)
    cp_rad = np.array([m.cp_radius[i] for i in range(ncp)], dtype=np.float64)
    rr = _quat_to_mat(torch.tensor(root[3:7]))
    x0 = torch.cat([torch.zeros(6, dtype=torch.float64), torch.tensor(dof[:, 0])])

    def frames(x, rp):
        com, rot = L.fk(x, rp, rr)
        return com - torch.einsum("lij,lj->li", rot, L.com), rot

    org, rot = frames(x0, torch.tensor(root[:3]))
    gaps = np.array([float((org[cp_link[i]] + rot[cp_link[i]] @ cp_pos[i])[2]) for i in range(ncp)]) - cp_rad
    order = np.argsort(gaps)
    root[2] += -0.002 - gaps[order[0]]
    gaps += -0.002 - gaps[order[0]]
    ics = [int(i) for i in order[:2]]
    assert gaps[ics[1]] < 0.002 and gaps[order[2]] > sp.contact_offset + 0.01 and int(m.cp_chain[ics[0]]) != int(m.cp_chain[ics[1]])
    rp = torch.tensor(root[:3])
    n = np.array([0.0, 0.0, 1.0])
    rows = []
    for ic in ics:
        Jc = torch.func.jacfwd(lambda x, ic=ic: (lambda o, r: o[cp_link[ic]] + r[cp_link[ic]] @ cp_pos[ic])(*frames(x, rp)))(x0).numpy()
        dR = torch.func.jacfwd(lambda x, ic=ic: frames(x, rp)[1][cp_link[ic]])(x0).numpy()
        Rl = frames(x0, rp)[1][cp_link[ic]].numpy()
        row = np.zeros(6 + nd)
        for k in range(6 + nd):
            Wk = dR[:, :, k] @ Rl.T
            row[k] = (Jc[:, k] + np.cross([Wk[2, 1], Wk[0, 2], Wk[1, 0]], -cp_rad[ic] * n))[2]
        rows.append(row)
    Jn = np.array(rows)
    xd = np.concatenate([root[7:10], root[10:13], dof[:, 1]])
    xdd, _, M = L.accelerations(root, dof, tau)
    Minv = np.linalg.inv(M)
    v_free = xd + h * xdd
    W = Jn @ Minv @ Jn.T
    assert W[0, 1] != 0.0                                                         # the rows couple, weakly: the base is heavy, the legs are light
    tgt = np.minimum(-gaps[ics] / h, float(sp.max_depenetration_velocity))
    lam_p = np.linalg.solve(W, tgt - Jn @ v_free)
    v_pos = v_free + Minv @ Jn.T @ lam_p
    lam_v = lam_p + np.linalg.solve(W, np.minimum(tgt, 0.0) - Jn @ v_pos)
    v_fin = v_pos + Minv @ Jn.T @ (lam_v - lam_p)
    assert (lam_p > 0).all() and (lam_v > 0).all()                                # both contacts active in both stages: the LCP is this linear solve
    props = _abi.default_dof_props(art, _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    r, d = root[None].copy(), dof[None].copy()
    _, contact, info = O.simulate_ref(m, sp, props, r, d, np.zeros((1, nd)), tau[None], friction=np.zeros(1, np.float32), tol=1e-13)
    assert info[0][0] == 2 and info[0][3] == 0
    got = np.concatenate([r[0][7:10], r[0][10:13], d[0][:, 1]])
    tol = 5e-7 if robot in EXACT_FRAMES else 5e-6
    assert np.abs(got - v_fin).max() < tol * max(1.0, np.abs(v_fin).max()), np.abs(got - v_fin).max()
    for k, ic in enumerate(ics):
        assert abs(contact[0][int(m.cp_body[ic])][2] - lam_v[k] / h) < tol * max(1.0, lam_v[k] / h)
    # production scheme, 4 + 1 sweeps
    r2, d2 = root[None].copy(), dof[None].copy()
    _, contact2 = O.simulate(m, sp, props, r2, d2, np.zeros((1, nd)), tau[None], friction=np.zeros(1, np.float32))
    got2 = np.concatenate([r2[0][7:10], r2[0][10:13], d2[0][:, 1]])
    dev = np.abs(got2 - v_fin).max() / max(1.0, np.abs(v_fin).max())
    fdev = max(abs(contact2[0][int(m.cp_body[ic])][2] - lam_v[k] / h) / (lam_v[k] / h) for k, ic in enumerate(ics))
    print(f"{robot}: production 4+1 sweeps vs the two-contact solution: velocity {dev:.2e}, force {fdev:.2e}")
    assert dev < 1e-3 and fdev < 1e-3


def _coulomb_sliding(W, vc, tgt_n, mu):
    """Impulse lam = (t1, t2, n) with  v+ = vc + W lam,  v+_n = tgt_n,  lam_t = -mu lam_n v+_t / |v+_t|  (maximum dissipation), by Newton
    iteration on the three equations (scipy)."""
    from scipy.optimize import fsolve

    def res(lam):
        v = vc + W @ lam
        vt = v[:2]
        s = np.linalg.norm(vt)
        return np.array([lam[0] + mu * lam[2] * vt[0] / s, lam[1] + mu * lam[2] * vt[1] / s, v[2] - tgt_n])

    ln0 = (tgt_n - vc[2]) / W[2, 2]
    lam0 = np.array([-mu * ln0 * vc[0] / np.linalg.norm(vc[:2]), -mu * ln0 * vc[1] / np.linalg.norm(vc[:2]), ln0])
    lam, _, ok, msg = fsolve(res, lam0, xtol=1e-14, full_output=True)
    assert ok == 1, msg
    return lam


@pytest.mark.parametrize("robot", ["hound", "anymal"])
def test_sliding_contact_obeys_coulomb_with_maximum_dissipation(robot):
    """Friction 0.3, a fast sideways drift: the contact slides, and the converged solver's impulse must be the root of Coulomb's law in
    operational space -- normal velocity on target, friction of magnitude mu lambda_n opposite to the POST-impulse tangential velocity -- in
    both stages."""
    h, mu = 0.005, 0.3
    c = _one_foot_case(robot, -0.002, [0.9, -0.6, -0.8], seed=6, h=h, spin=0.03)
    sp, J, Minv, v_free, tgt = c["sp"], c["J"], c["Minv"], c["v_free"], c["tgt"]
    sp.plane_dynamic_friction = sp.plane_static_friction = mu
    W = J @ Minv @ J.T
    lam_p = _coulomb_sliding(W, J @ v_free, tgt, mu)
    v_pos = v_free + Minv @ J.T @ lam_p
    lam_v = _coulomb_sliding(W, J @ v_pos - W @ lam_p, min(tgt, 0.0), mu)      # total impulse of the velocity stage, started from lam_p
    v_fin = v_pos + Minv @ J.T @ (lam_v - lam_p)
    for lam, v in ((lam_p, v_pos), (lam_v, v_fin)):
        vt = (J @ v)[:2]
        assert lam[2] > 0 and np.linalg.norm(vt) > 0.2                            # sliding, in contact
        assert abs(np.hypot(lam[0], lam[1]) - mu * lam[2]) < 1e-10 and lam[:2] @ vt < 0
    props = _abi.default_dof_props(c["art"], _abi.DOF_MODE_EFFORT, 0.0, 0.0)
    r, d = c["root"][None].copy(), c["dof"][None].copy()
    _, contact, info = O.simulate_ref(c["m"], sp, props, r, d, np.zeros((1, c["nd"])), c["tau"][None], friction=np.full(1, mu, np.float32), tol=1e-13)
    assert info[0][0] == 1 and info[0][3] == 0
    _compare(c, r, d, contact, v_pos, v_fin, lam_v, h, 2e-6 if robot in EXACT_FRAMES else 1e-5)
