"""Rough-terrain task for the fork's Hound quadruped, drop-in for the reference's ``tasks/Hound_terrain.py``: the
AnymalTerrain task with the Hound URDF and names, ``collapseFixedJoints`` from the config (:220), cylinders kept (:221),
termination that always includes knee ("thigh") and ``baseName`` ("shoulder") contacts (:304-311) and a base-height target
of 0.48 m (:347)."""
from __future__ import annotations

from .anymal_terrain import AnymalTerrain


class HoundTerrain(AnymalTerrain):
    # the calf box (4 corners) and the foot sphere rest on the ground together: 4 slots dropped 16 % of the candidates (counted by
    # b2g_sim_contact_stats), 6 drop < 0.1 %
    CONTACT_SLOTS = 6
    ACTOR_NAME = "houndterrain"
    BASE_NAME = "trunk"
    HOUND_TERMINATION = True
    BASE_HEIGHT_TARGET = 0.48

    def _asset_options(self):
        o = super()._asset_options()
        o.collapse_fixed_joints = self.cfg["env"]["urdfAsset"]["collapseFixedJoints"]
        o.replace_cylinder_with_capsule = False
        o.flip_visual_attachments = False
        return o

    def _extra_termination_names(self, body_names):
        base_name = self.cfg["env"]["urdfAsset"]["baseName"]
        return [s for s in body_names if base_name in s]
