"""CPU half of the physics known-answer tests (tests/physics_kats.py): the kernel code on the host lane emulator and the oracle's
converged reference solver.  The GPU half (tests/test_physics_gpu.py) runs the same checks on libb200gym.so at full size."""
import numpy as np

from tests import physics_kats as pk
from tests.backends import EmuBackend


def make(art, params, props, n):
    return EmuBackend(art, params, props, n)


def test_cartpole_closed_form():
    ey, eth = pk.check_cartpole_closed_form(make, n=64)
    assert ey < 1e-3 and eth < 1e-3


def test_torque_free_precession():
    pk.check_torque_free_precession(make, dt=0.002, t_end=0.5, n=1)


def test_block_on_slope_plane():
    pk.check_block_on_slope(make, heightfield=False, t_end=0.4)


def test_block_on_slope_heightfield():
    pk.check_block_on_slope(make, heightfield=True, t_end=0.4)


def test_resting_force_is_mg():
    out = pk.check_resting_force(make, "anymal", n=2)
    assert abs(out["total_over_mg"] - 1.0) < 0.01


def test_production_scheme_converges_to_reference():
    """More sweeps of the production iteration -> the converged reference, monotonically, to 1e-6: the Jacobi split across chains,
    the slot cap (not binding here) and the cheaper sliding update change the path, not the fixed point."""
    out = pk.production_scheme_convergence(n=128)
    rows = out["rows"]
    assert out["converged_fraction"] > 0.95
    errs = [r["joint_vel_err_median"] for r in rows]
    assert all(b < a for a, b in zip(errs, errs[1:])), errs
    assert errs[-1] < 1e-6 and rows[-1]["root_lin_vel_err_median"] < 1e-7 and rows[-1]["net_force_err_over_mg_median"] < 1e-6, rows[-1]
    assert errs[0] < 1.0          # the 4 + 1 production budget: a few tenths of a rad/s on the joints of a robot hit by random actions


def test_reference_equals_production_without_contact():
    import numpy as np

    from isaacgymenv_b200 import _abi
    from oracle import dyn_oracle as O
    from tests import kernel_checks as kc

    art = kc.load_robot("anymal")
    sp = kc.flat_params(ground=False)
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    model = _abi.pack_model(art)
    rng = np.random.default_rng(0)
    root, dof = kc.random_flying_state(art, 8, rng)
    root, dof = root.astype(np.float64), dof.astype(np.float64)
    tgt = rng.normal(size=(8, 12)) * 0.3
    r1, d1, r2, d2 = root.copy(), dof.copy(), root.copy(), dof.copy()
    f1, c1 = O.simulate(model, sp, props, r1, d1, tgt, np.zeros((8, 12)))
    f2, c2, info = O.simulate_ref(model, sp, props, r2, d2, tgt, np.zeros((8, 12)))
    assert np.array_equal(r1, r2) and np.array_equal(d1, d2) and np.array_equal(f1, f2)
    assert info[:, 0].max() == 0


def test_reference_hard_joint_limits():
    pk.check_hard_joint_limits_reference()


def test_solver_deviation_table_smoke():
    t = pk.solver_deviation(make, n=16, steps=30, sample_every=10, n_ref=16)
    assert t["samples"] > 0 and t["device_vs_production_oracle"]["joint_vel_err_rad_s"]["max"] < 5e-3      # kernel == its own oracle
    assert t["contact_cap"]["env_substeps"] == 16 * 30 * 2
    assert t["contact_flag_agreement"] > 0.95


def test_contact_drop_counter_counts():
    """A robot dropped flat on its belly has more root-link candidates inside the contact offset than lane 0 has slots: the drop must
    be counted, not silent."""
    from isaacgymenv_b200 import _abi
    from tests import kernel_checks as kc

    art = kc.load_robot("anymal")
    sp = kc.flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    rng = np.random.default_rng(0)
    root, dof = kc.standing_state(art, 2, rng, 0.62)
    root[:, 2] = 0.10
    root[:, 3:7] = [0, 0, 0, 1]
    dof[:, :, 0] = 0.0
    dof[:, 1::3, 0] = 1.5          # legs folded up: the base box rests on the ground
    dof[:, 2::3, 0] = -2.5
    be = make(art, sp, props, 2)
    be.contact_stats(reset=True)
    be.set_state(root, dof)
    tgt = dof[:, :, 0].copy()
    for _ in range(10):
        be.simulate(tgt, np.zeros_like(tgt))
    st = be.contact_stats()
    assert st[3] == 2 * 10 * 2 and st[0] > 0
    assert st[1] >= 0 and st[2] <= st[3]
