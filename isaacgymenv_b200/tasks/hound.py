"""Flat-terrain task for the fork's Hound quadruped, drop-in for the reference's ``tasks/hound.py`` -- a clone of
its Anymal task with the Hound URDF (``:168``), ``collapseFixedJoints`` taken from the config (``:172``), cylinders
kept (``:173``) and body names foot / thigh / trunk (``:192-224``)."""
from __future__ import annotations

from .anymal import Anymal


class Hound(Anymal):
    # the calf box (4 corners) and the foot sphere rest on the ground together: 4 slots dropped 16 % of the candidates (counted by
    # b2g_sim_contact_stats), 6 drop < 0.1 %
    CONTACT_SLOTS = 6
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
