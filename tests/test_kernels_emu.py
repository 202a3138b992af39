"""CPU tests: the CUDA kernels' per-thread code, compiled for the host by tests/emu (lock-stepped lanes),
against the oracle and the reference's golden vectors.  The same checks run on the B200 through the C ABI
in tests/test_kernels_gpu.py."""
import pytest

from tests import kernel_checks as kc
from tests.backends import EmuBackend


def make(art, params, props, n):
    return EmuBackend(art, params, props, n)


@pytest.mark.parametrize("robot", ["anymal", "hound", "useful_hound", "cartpole", "houndarm", "manipulator"])
def test_forward_dynamics(robot):
    kc.check_forward_dynamics(make, robot, n=6)


@pytest.mark.parametrize("robot", ["anymal", "useful_hound"])
def test_env_scale_domain_randomisation(robot):
    kc.check_env_scale(make, robot, n=4)


@pytest.mark.parametrize("robot", ["anymal", "useful_hound"])
def test_link_scale_domain_randomisation(robot):
    kc.check_link_scale(make, robot, n=3, steps=5)


@pytest.mark.parametrize("robot,drive", [("anymal", "pos"), ("hound", "pos"), ("anymal_minimal", "effort")])
def test_simulate_horizon(robot, drive):
    kc.check_simulate_horizon(make, robot, n=6, steps=10, drive=drive)


@pytest.mark.parametrize("robot", ["anymal", "hound"])
def test_post_physics_golden(robot):
    kc.check_post_physics_golden(make, robot)


def test_reset_draws():
    kc.check_reset_draws(make, "anymal", n=8)


def test_fused_step():
    kc.check_fused_step(make, "anymal", n=6, steps=25)


def test_cartpole_golden():
    kc.check_cartpole_golden(make)


def test_cartpole_step():
    kc.check_cartpole_step(make)


@pytest.mark.parametrize("name", ["anymal_terrain_plane.npz", "anymal_terrain_trimesh.npz", "hound_terrain_plane.npz"])
def test_terrain_golden(name):
    kc.check_terrain_golden(make, name)


@pytest.mark.parametrize("robot,heightfield", [("anymal_minimal", True), ("hound", False)])
def test_terrain_step(robot, heightfield):
    kc.check_terrain_step(make, robot, n=6, heightfield=heightfield)


def test_useful_hound_golden():
    kc.check_useful_golden(make)


def test_useful_hound_step():
    kc.check_useful_step(make, n=4)


@pytest.mark.parametrize("robot", ["useful_hound", "anymal", "cartpole", "houndarm", "manipulator"])
def test_jacobian_mass_matrix(robot):
    kc.check_jacobian_mass_matrix(make, robot)


@pytest.mark.parametrize("robot", ["useful_hound", "anymal_minimal"])
def test_self_collision_against_the_base(robot):
    kc.check_self_collision(make, robot, n=4)


def test_drive_saturates_at_the_effort_limit():
    kc.check_drive_saturation(make)


def test_root_velocity_limits():
    kc.check_root_velocity_limits(make)


def test_houndarm_fused_step():
    kc.check_houndarm_step(make, n=4)


def test_manipulator_fused_step():
    kc.check_houndarm_step(make, n=4, robot="manipulator")


def _emu_sim_hf(art, sp, props, hf_t, samples, root, dof, steps):
    from tests.emu import emu
    from isaacgymenv_b200 import _abi
    import numpy as np

    m = _abi.pack_model(art)
    tgt = dof[:, :, 0].copy()
    zero = np.zeros_like(tgt)
    contact = None
    for _ in range(steps):
        _, contact = emu.simulate(m, sp, props, root, dof, tgt, zero, heightfield=hf_t, hf_samples=samples)
    return root, dof, contact


@pytest.mark.parametrize("robot", ["useful_hound", "anymal_minimal"])
def test_hf_coarse_bound_identical(robot):
    from tests.emu import emu

    emu.hfc_stats()
    kc.check_hf_coarse_identical(_emu_sim_hf, robot)
    tested, skipped = emu.hfc_stats()
    print(f"{robot}: coarse bound skipped {skipped} of {tested} link tests")
    assert tested > 0 and skipped > 0.25 * tested, (tested, skipped)


# ---- opt-in segment variant (B2G_SEGMENTS=1): chains cut into pieces of at most three links that own a lane each ----
@pytest.fixture
def segments(monkeypatch):
    monkeypatch.setenv("B2G_SEGMENTS", "1")      # read by the emulator's variant choice and by the oracle's per-piece contact cap


def test_segment_variant_forward_dynamics(segments):
    kc.check_forward_dynamics(make, "useful_hound", n=6)


def test_segment_variant_simulate_horizon(segments):
    kc.check_simulate_horizon(make, "useful_hound", n=6, steps=10, drive="effort")


def test_segment_variant_useful_step(segments):
    kc.check_useful_step(make, n=4)


def test_segment_variant_hf_coarse_bound_identical(segments):
    test_hf_coarse_bound_identical("useful_hound")
