"""Flat-terrain task for the fork's Hound quadruped, drop-in for the reference's ``tasks/hound.py`` -- a clone of
its Anymal task with the Hound URDF (``:168``), ``collapseFixedJoints`` taken from the config (``:172``), cylinders
kept (``:173``) and body names foot / thigh / trunk (``:192-224``)."""
from __future__ import annotations

from .anymal import Anymal


class Hound(Anymal):
    ASSET_FILE = "urdf/Hound_new/Hound.urdf"
    ACTOR_NAME = "hound"
    BASE_NAME = "trunk"
    KNEE_KEY = "thigh"

    def _extremity_key(self, collapse):
        return "calf" if collapse else "foot"

    def _asset_options(self):
        o = super()._asset_options()
        o.collapse_fixed_joints = self.cfg["env"]["urdfAsset"]["collapseFixedJoints"]
        o.replace_cylinder_with_capsule = False
        return o
