"""GPU parity tests: libb200gym.so through its C ABI on cuda:0 against the oracle and the reference's
golden vectors (same checks as the CPU emulator tests, larger batches)."""
import pytest

from tests import kernel_checks as kc

pytestmark = pytest.mark.gpu


def make(art, params, props, n):
    from tests.backends import CudaBackend

    return CudaBackend(art, params, props, n)


@pytest.mark.parametrize("robot", ["anymal", "hound", "useful_hound", "cartpole", "houndarm", "manipulator"])
def test_forward_dynamics(robot):
    kc.check_forward_dynamics(make, robot, n=256)


@pytest.mark.parametrize("robot", ["anymal", "useful_hound"])
def test_env_scale_domain_randomisation(robot):
    kc.check_env_scale(make, robot, n=64)


@pytest.mark.parametrize("robot", ["anymal", "useful_hound"])
def test_link_scale_domain_randomisation(robot):
    kc.check_link_scale(make, robot, n=48)


@pytest.mark.parametrize("robot,drive", [("anymal", "pos"), ("hound", "pos"), ("anymal_minimal", "effort"), ("useful_hound", "effort")])
def test_simulate_horizon(robot, drive):
    kc.check_simulate_horizon(make, robot, n=128, steps=10, drive=drive)


@pytest.mark.parametrize("robot", ["anymal", "hound"])
def test_post_physics_golden(robot):
    kc.check_post_physics_golden(make, robot)


@pytest.mark.parametrize("robot", ["anymal", "hound"])
def test_reset_draws(robot):
    kc.check_reset_draws(make, robot, n=1000)


@pytest.mark.parametrize("robot", ["anymal", "hound"])
def test_fused_step(robot):
    kc.check_fused_step(make, robot, n=64, steps=40)


def test_cartpole_golden():
    kc.check_cartpole_golden(make)


def test_cartpole_step():
    kc.check_cartpole_step(make)


@pytest.mark.parametrize("name", ["anymal_terrain_plane.npz", "anymal_terrain_trimesh.npz", "hound_terrain_plane.npz"])
def test_terrain_golden(name):
    kc.check_terrain_golden(make, name)


@pytest.mark.parametrize("robot,heightfield", [("anymal_minimal", True), ("hound", False)])
def test_terrain_step(robot, heightfield):
    kc.check_terrain_step(make, robot, n=64, heightfield=heightfield)


def test_useful_hound_golden():
    kc.check_useful_golden(make)


def test_useful_hound_step():
    kc.check_useful_step(make, n=64)


@pytest.mark.parametrize("robot", ["useful_hound", "anymal", "cartpole", "houndarm", "manipulator"])
def test_jacobian_mass_matrix(robot):
    kc.check_jacobian_mass_matrix(make, robot)


@pytest.mark.parametrize("robot", ["useful_hound", "anymal_minimal"])
def test_self_collision_against_the_base(robot):
    kc.check_self_collision(make, robot, n=64)


def test_drive_saturates_at_the_effort_limit():
    kc.check_drive_saturation(make)


def test_root_velocity_limits():
    kc.check_root_velocity_limits(make)


def test_houndarm_fused_step():
    kc.check_houndarm_step(make, n=48)


def test_manipulator_fused_step():
    kc.check_houndarm_step(make, n=48, robot="manipulator")


def _gpu_sim_hf(art, sp, props, hf_t, samples, root, dof, steps):
    import ctypes as C

    import numpy as np

    be = make(art, sp, props, root.shape[0])
    try:
        smp = np.ascontiguousarray(samples, np.int16)
        be._lib.check(be.lib.b2g_sim_add_heightfield(be.sim, C.byref(hf_t), smp.ctypes.data_as(C.c_void_p)), "add_heightfield")
        be.set_state(root, dof)
        tgt = dof[:, :, 0].copy()
        contact = None
        for _ in range(steps):
            _, contact = be.simulate(tgt, np.zeros_like(tgt))
        r, d = be.get_state()
    finally:
        be.close()
    return r, d, contact


@pytest.mark.parametrize("robot", ["useful_hound", "hound", "anymal_minimal"])
def test_hf_coarse_bound_identical(robot):
    """The coarse heightfield bound (per-link contact early-out on rough terrain) never changes a result: library built with and
    without it (B2G_NO_HFC=1), bit for bit, robots standing, tumbling and outside the field."""
    kc.check_hf_coarse_identical(_gpu_sim_hf, robot, steps=30, n=256)


# ---- opt-in segment variant (B2G_SEGMENTS=1, read when the sim is created): <8,3> kernels ----
@pytest.fixture
def segments(monkeypatch):
    monkeypatch.setenv("B2G_SEGMENTS", "1")


def test_segment_variant_forward_dynamics(segments):
    kc.check_forward_dynamics(make, "useful_hound", n=256)


def test_segment_variant_simulate_horizon(segments):
    kc.check_simulate_horizon(make, "useful_hound", n=128, steps=10, drive="effort")


def test_segment_variant_useful_step(segments):
    kc.check_useful_step(make, n=64)


def test_segment_variant_useful_golden(segments):
    kc.check_useful_golden(make)
