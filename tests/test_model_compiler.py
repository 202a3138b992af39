"""Model compiler (URDF -> flat articulation): known answers from SURVEY.md appendix A, ordering conventions the
tasks rely on, and consistency of the committed compiled models with a fresh compile when the reference assets
are available (build container only)."""
import os

import numpy as np
import pytest

from isaacgymenv_b200 import _abi
from isaacgymenv_b200.model import urdf
from isaacgymenv_b200.model.store import COMPILED_DIR, find_compiled, load_articulation
from tests.kernel_checks import default_pose, load_robot

REF_ASSETS = "/root/reference/assets"


def test_anymal_collapsed_masses_and_order():
    a = load_robot("anymal")
    assert a.body_names == ["base", "LF_HIP", "LF_THIGH", "LF_SHANK", "LH_HIP", "LH_THIGH", "LH_SHANK", "RF_HIP", "RF_THIGH", "RF_SHANK",
                            "RH_HIP", "RH_THIGH", "RH_SHANK"]
    assert a.dof_names == ["LF_HAA", "LF_HFE", "LF_KFE", "LH_HAA", "LH_HFE", "LH_KFE", "RF_HAA", "RF_HFE", "RF_KFE", "RH_HAA", "RH_HFE", "RH_KFE"]
    np.testing.assert_allclose(a.mass[:4], [26.37317, 2.78100, 3.07100, 0.58842], atol=5e-5)
    assert abs(a.total_mass - 52.13485) < 1e-4
    assert list(a.chain_start) == [0, 3, 6, 9] and list(a.chain_len) == [3, 3, 3, 3]
    np.testing.assert_allclose(a.lower[0], -0.72)
    np.testing.assert_allclose(a.upper[0], 0.49)
    assert (a.effort == 80).all() and (a.velocity == 20).all()


def test_anymal_minimal_masses():
    a = load_robot("anymal_minimal")
    np.testing.assert_allclose(a.mass[:4], [27.80286, 2.51203, 3.27327, 0.55505], atol=5e-5)
    assert abs(a.total_mass - 53.16426) < 1e-4


def test_hound_models():
    h = load_robot("hound")
    assert h.num_bodies == 17 and h.num_dofs == 12 and abs(h.total_mass - 43.77) < 0.01
    assert h.body_names[:5] == ["trunk", "FL_shoulder", "FL_thigh", "FL_calf", "FL_foot"]
    # foot bodies ride on the calf link (fixed joint kept as an API body, folded dynamically)
    assert h.body_link[4] == h.body_link[3] == 3
    u = load_robot("useful_hound")
    assert u.num_bodies == 24 and u.num_dofs == 18 and abs(u.total_mass - 49.17) < 0.01
    assert list(u.chain_len) == [3, 3, 3, 3, 6]
    assert u.dof_names[12:] == ["joint1", "joint2", "joint3", "joint4", "joint5", "joint6"]
    assert u.body_link[u.body_names.index("link1")] == 0      # arm base is fixed to the trunk


def test_foot_positions_known_answers():
    """SURVEY.md appendix A forward-kinematics check values at the default joint angles."""
    a = load_robot("anymal")
    lp, lr = urdf.forward_kinematics(a, default_pose(a))
    h = load_robot("hound")
    bp, _ = urdf.body_poses(h, default_pose(h))
    for name, want in (("FL_foot", (0.36251, 0.2135, -0.47723)), ("FR_foot", (0.36251, -0.2135, -0.47723)),
                       ("RL_foot", (-0.33549, 0.2135, -0.47723)), ("RR_foot", (-0.33549, -0.2135, -0.47723))):
        np.testing.assert_allclose(bp[h.body_names.index(name)], want, atol=2e-5)
    u = load_robot("useful_hound")
    bp, _ = urdf.body_poses(u, np.zeros(18))
    np.testing.assert_allclose(bp[u.body_names.index("end_link")], (0.561, 0.0, 0.503), atol=1e-5)
    # the ANYmal foot sphere (r 0.03) is the lowest contact candidate of each shank
    for leg in range(4):
        link = 3 * leg + 3
        pts = [(lp[link] + lr[link] @ a.cp_pos[i])[2] - a.cp_radius[i] for i in range(len(a.cp_link)) if a.cp_link[i] == link]
        assert min(pts) < -0.5


def test_cartpole_model():
    c = load_robot("cartpole")
    assert c.fixed_base and c.dof_names == ["slider_to_cart", "cart_to_pole"]
    assert list(c.joint_type) == [urdf.JOINT_PRISMATIC, urdf.JOINT_REVOLUTE]
    np.testing.assert_allclose(c.mass[1:], [1.0, 1.0])
    # pole inertia derived from its collision box (0.04 x 0.06 x 1.0, mass 1)
    np.testing.assert_allclose(np.diag(c.inertia[2]), [(0.06 ** 2 + 1) / 12, (0.04 ** 2 + 1) / 12, (0.04 ** 2 + 0.06 ** 2) / 12], rtol=1e-6)


def test_pack_model_groups_contact_candidates():
    for name in ("anymal", "hound", "useful_hound"):
        art = load_robot(name)
        m = _abi.pack_model(art)
        links = list(m.cp_link)[:m.n_cpts]
        owners = list(m.cp_chain)[:m.n_cpts]
        seen = set()
        prev = None
        for l, o in zip(links, owners):
            key = ("root", o) if l == 0 else ("link", l)
            if key != prev:
                assert key not in seen, "contact candidates must be contiguous per link / per root owner"
                seen.add(key)
                prev = key
        first_root = links.index(0) if 0 in links else len(links)
        assert all(l == 0 for l in links[first_root:])


def test_lenient_float():
    assert urdf._lenient_float("0.0.0000001") == 0.0      # Hound URDF quirk
    assert urdf._lenient_float(" 1e-3 ") == 1e-3


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets only exist in the build container")
def test_committed_models_match_fresh_compile():
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import compile_assets

    for rel, opts in compile_assets.JOBS:
        fresh = urdf.compile_urdf(os.path.join(REF_ASSETS, rel), opts)
        stored = load_articulation(find_compiled(rel, opts))
        assert fresh.body_names == stored.body_names and fresh.dof_names == stored.dof_names
        np.testing.assert_allclose(fresh.mass, stored.mass, rtol=1e-12)
        np.testing.assert_allclose(fresh.inertia, stored.inertia, rtol=1e-9, atol=1e-15)
        np.testing.assert_allclose(fresh.cp_pos, stored.cp_pos, atol=1e-12)
    assert os.path.isdir(COMPILED_DIR)
