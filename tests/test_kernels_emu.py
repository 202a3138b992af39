"""CPU tests: the CUDA kernels' per-thread code, compiled for the host by tests/emu (lock-stepped lanes),
against the oracle and the reference's golden vectors.  The same checks run on the B200 through the C ABI
in tests/test_kernels_gpu.py."""
import pytest

from tests import kernel_checks as kc
from tests.backends import EmuBackend


def make(art, params, props, n):
    return EmuBackend(art, params, props, n)


@pytest.mark.parametrize("robot", ["anymal", "hound", "useful_hound", "cartpole"])
def test_forward_dynamics(robot):
    kc.check_forward_dynamics(make, robot, n=6)


@pytest.mark.parametrize("robot", ["anymal", "useful_hound"])
def test_env_scale_domain_randomisation(robot):
    kc.check_env_scale(make, robot, n=4)


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


@pytest.mark.parametrize("robot", ["useful_hound", "anymal", "cartpole"])
def test_jacobian_mass_matrix(robot):
    kc.check_jacobian_mass_matrix(make, robot)


def test_houndarm_fused_step():
    kc.check_houndarm_step(make, n=4)
