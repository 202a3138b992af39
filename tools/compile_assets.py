#!/usr/bin/env python
"""Compile the hot-path robots from the reference's URDF assets into in-tree model files.

Run in the build container (needs /root/reference); the outputs under
``isaacgymenv_b200/assets/compiled`` are committed.  Asset options mirror the reference call sites:
  anymal.urdf          tasks/anymal.py:168-181          (collapse, capsules, density 0.001)
  anymal_minimal.urdf  tasks/anymal_terrain.py:213-229  (collapse, capsules)
  Hound.urdf           tasks/hound.py:168-181, cfg/task/Hound.yaml:53 (no collapse, cylinders kept)
  UsefulHound Hound    tasks/useful_hound.py:316-327    (no collapse)
  cartpole.urdf        tasks/cartpole.py:86-88          (fixed base)
  open_manipulator_p   tasks/hound_arm.py:203-216       (fixed base, no collapse, gravity disabled)
  franka_panda_manipulator  tasks/manipulator.py:199-214  (fixed base, no collapse, gravity disabled; no <inertial>: mass properties
                                                          from the convex hulls of the collision meshes at the default density)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from isaacgymenv_b200.model.store import COMPILED_DIR, options_key, save_articulation  # noqa: E402
from isaacgymenv_b200.model.urdf import AssetOptions, compile_urdf  # noqa: E402

ROOT = os.environ.get("B2G_REFERENCE_ROOT", "/root/reference") + "/assets"
JOBS = [
    ("urdf/anymal_c/urdf/anymal.urdf", AssetOptions(collapse_fixed_joints=True, replace_cylinder_with_capsule=True, density=0.001, thickness=0.01)),
    ("urdf/anymal_c/urdf/anymal_minimal.urdf", AssetOptions(collapse_fixed_joints=True, replace_cylinder_with_capsule=True, density=0.001, thickness=0.01)),
    ("urdf/Hound_new/Hound.urdf", AssetOptions(collapse_fixed_joints=False, replace_cylinder_with_capsule=False, density=0.001, thickness=0.01)),
    ("urdf/Hound_new/Hound.urdf", AssetOptions(collapse_fixed_joints=True, replace_cylinder_with_capsule=True, density=0.001, thickness=0.01)),
    ("urdf/UsefulHound/urdf/Hound.urdf", AssetOptions(collapse_fixed_joints=False, replace_cylinder_with_capsule=False, density=0.001, thickness=0.01)),
    ("urdf/UsefulHound/urdf/Hound.urdf", AssetOptions(collapse_fixed_joints=True, replace_cylinder_with_capsule=True, density=0.001, thickness=0.01)),
    ("urdf/cartpole.urdf", AssetOptions(fix_base_link=True)),
    ("urdf/open_manipulator_p_gazebo/urdf/open_manipulator_p.urdf", AssetOptions(fix_base_link=True, collapse_fixed_joints=False,
                                                                                replace_cylinder_with_capsule=False, disable_gravity=True, thickness=0.001)),
    ("urdf/franka_description/robots/franka_panda_manipulator.urdf", AssetOptions(fix_base_link=True, collapse_fixed_joints=False, disable_gravity=True,
                                                                                  thickness=0.001)),
]

if __name__ == "__main__":
    for rel, opts in JOBS:
        art = compile_urdf(os.path.join(ROOT, rel), opts)
        out = os.path.join(COMPILED_DIR, options_key(rel, opts) + ".json")
        save_articulation(art, out)
        print(f"{rel}: {art.num_bodies} bodies, {art.num_dofs} dofs, {len(art.cp_link)} contact pts, {art.total_mass:.5f} kg -> {os.path.relpath(out)}")
