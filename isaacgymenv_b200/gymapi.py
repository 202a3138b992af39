"""``gymapi`` shim: the slice of Isaac Gym's ``isaacgym.gymapi`` that the hot-path tasks call, served by
``libb200gym.so`` (hand-written sm_100a kernels) instead of the closed Isaac Gym binary + PhysX.

Method names, argument order and return conventions follow the reference call sites
(``isaacgymenvs/tasks/base/vec_task.py:247-262,337,382-386``, ``tasks/anymal.py:110-126,159-229,258-297``,
``tasks/anymal_terrain.py:196-209,236,281,439-448``, ``tasks/cartpole.py:86-114,163``) so task code written
against Isaac Gym runs unchanged.  Differences that matter are stated in DESIGN.md ("gym API shim").

There is no CPU simulation path: ``create_sim`` needs a CUDA device and fails loudly otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np

from . import _abi, _lib
from .model import store as _store
from .model import urdf as _urdf

# ---- constants (values as in Isaac Gym Preview 4) ----
SIM_PHYSX = 0
SIM_FLEX = 1
UP_AXIS_Y = 0
UP_AXIS_Z = 1
DOF_MODE_NONE = _abi.DOF_MODE_NONE
DOF_MODE_POS = _abi.DOF_MODE_POS
DOF_MODE_VEL = _abi.DOF_MODE_VEL
DOF_MODE_EFFORT = _abi.DOF_MODE_EFFORT
STATE_NONE, STATE_POS, STATE_VEL, STATE_ALL = 0, 1, 2, 3
DOMAIN_SIM, DOMAIN_ENV, DOMAIN_ACTOR = 0, 1, 2
ENV_SPACE, LOCAL_SPACE, GLOBAL_SPACE = 0, 1, 2
CC_NEVER, CC_LAST_SUBSTEP, CC_ALL_SUBSTEPS = 0, 1, 2
KEY_ESCAPE, KEY_V, KEY_R = 256, 86, 82
MESH_VISUAL, MESH_COLLISION, MESH_VISUAL_AND_COLLISION = 1, 2, 3


def ContactCollection(v):
    return int(v)


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return f"Vec3({self.x}, {self.y}, {self.z})"


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)

    @staticmethod
    def from_axis_angle(axis, angle):
        a = np.array([axis.x, axis.y, axis.z], dtype=np.float64)
        a /= np.linalg.norm(a)
        s = np.sin(angle / 2)
        return Quat(a[0] * s, a[1] * s, a[2] * s, np.cos(angle / 2))

    def __repr__(self):
        return f"Quat({self.x}, {self.y}, {self.z}, {self.w})"


class Transform:
    def __init__(self, p: Optional[Vec3] = None, r: Optional[Quat] = None):
        self.p = p or Vec3()
        self.r = r or Quat()


class PlaneParams:
    def __init__(self):
        self.normal = Vec3(0.0, 0.0, 1.0)
        self.distance = 0.0
        self.static_friction = 1.0
        self.dynamic_friction = 1.0
        self.restitution = 0.0
        self.segmentation_id = 0


class TriangleMeshParams:
    def __init__(self):
        self.nb_vertices = 0
        self.nb_triangles = 0
        self.transform = Transform()
        self.static_friction = 1.0
        self.dynamic_friction = 1.0
        self.restitution = 0.0
        self.segmentation_id = 0


class HeightFieldParams:
    """Isaac Gym's gymapi.HeightFieldParams (used by ``add_heightfield``)."""

    def __init__(self):
        self.column_scale = 1.0
        self.row_scale = 1.0
        self.vertical_scale = 1.0
        self.nbRows = 0
        self.nbColumns = 0
        self.transform = Transform()
        self.static_friction = 1.0
        self.dynamic_friction = 1.0
        self.restitution = 0.0


class AssetOptions(_urdf.AssetOptions):
    pass


class _PhysXParams:
    def __init__(self):
        self.solver_type = 1
        self.num_threads = 0
        self.use_gpu = True
        self.num_position_iterations = 4
        self.num_velocity_iterations = 1
        self.contact_offset = 0.02
        self.rest_offset = 0.001
        self.bounce_threshold_velocity = 0.2
        self.max_depenetration_velocity = 100.0
        self.default_buffer_size_multiplier = 2.0
        self.max_gpu_contact_pairs = 1024 * 1024
        self.num_subscenes = 0
        self.contact_collection = CC_LAST_SUBSTEP
        self.friction_offset_threshold = 0.04
        self.friction_correlation_distance = 0.025
        self.always_use_articulations = False
        # not Isaac Gym parameters: joint limits are one-sided implicit spring-dampers in this engine
        self.joint_limit_stiffness = 2000.0
        self.joint_limit_damping = 20.0
        # not an Isaac Gym parameter either: contact slots per solver lane (0 = the library's default of 4), b2g_sim_params
        self.max_contacts_per_chain = 0


class _FlexParams:
    pass


class SimParams:
    """gymapi.SimParams (reference: vec_task.py:514-562)."""

    def __init__(self):
        self.dt = 1.0 / 60.0
        self.substeps = 2
        self.up_axis = UP_AXIS_Y
        self.gravity = Vec3(0.0, -9.8, 0.0)
        self.use_gpu_pipeline = False
        self.num_client_threads = 0
        self.enable_actor_creation_warning = True
        self.physx = _PhysXParams()
        self.flex = _FlexParams()


class RigidShapeProperties:
    def __init__(self):
        self.friction = 1.0
        self.rolling_friction = 0.0
        self.torsion_friction = 0.0
        self.restitution = 0.0
        self.compliance = 0.0
        self.thickness = 0.0
        self.contact_offset = 0.02
        self.rest_offset = 0.0
        self.filter = 0


class RigidBodyProperties:
    def __init__(self, mass=0.0, com=None, inertia=None):
        self.mass = float(mass)
        self.com = com or Vec3()
        self.inertia = inertia
        self.invMass = 1.0 / mass if mass > 0 else 0.0
        self.flags = 0


DOF_PROPS_DTYPE = np.dtype([("hasLimits", np.bool_), ("lower", np.float32), ("upper", np.float32), ("driveMode", np.int32),
                            ("velocity", np.float32), ("effort", np.float32), ("stiffness", np.float32), ("damping", np.float32),
                            ("friction", np.float32), ("armature", np.float32)])


class Asset:
    """Result of ``gym.load_asset``: a compiled articulation plus mutable shape properties."""

    def __init__(self, art: _urdf.Articulation, options: AssetOptions, source: str):
        self.art = art
        self.options = options
        self.source = source
        self.shape_friction = 1.0
        self.n_shapes = max(len(art.cp_link), 1)


class Env:
    def __init__(self, sim: "Sim", index: int):
        self.sim = sim
        self.index = index
        self.actors: List[str] = []


class GymTensor:
    """What ``gym.acquire_*_tensor`` returns and ``gymtorch.wrap_tensor`` consumes."""

    def __init__(self, desc: Optional[_abi.TensorDesc] = None, tensor=None):
        self.desc = desc
        self.tensor = tensor    # set by gymtorch.unwrap_tensor (keeps the torch tensor alive)

    @property
    def data_ptr(self) -> int:
        if self.tensor is not None:
            return self.tensor.data_ptr()
        return int(self.desc.data)


class Sim:
    def __init__(self, compute_device: int, graphics_device: int, sim_type: int, params: SimParams):
        self.compute_device = compute_device
        self.graphics_device = graphics_device
        self.sim_type = sim_type
        self.params = params
        self.handle = C.c_void_p()
        self.envs: List[Env] = []
        self.asset: Optional[Asset] = None
        self.actor_name = None
        self.actor_pose = None
        self.dof_props = None           # numpy structured array
        self.env_friction: List[float] = []
        self.actor_poses: List[tuple] = []
        self.env_spacing = 0.0
        self.num_per_row = 1
        self.prepared = False
        self.frame_count = 0
        self.model_struct = None
        self._keep = []

    # -- helpers --
    def c_params(self) -> _abi.SimParams:
        p, px = self.params, self.params.physx
        g = p.gravity
        c = _abi.SimParams(dt=p.dt, substeps=int(p.substeps), num_position_iterations=int(px.num_position_iterations),
                           num_velocity_iterations=int(px.num_velocity_iterations), contact_offset=px.contact_offset,
                           rest_offset=px.rest_offset, bounce_threshold_velocity=px.bounce_threshold_velocity,
                           max_depenetration_velocity=px.max_depenetration_velocity, plane_static_friction=1.0,
                           plane_dynamic_friction=1.0, plane_restitution=0.0, has_ground=0,
                           joint_limit_stiffness=float(getattr(px, "joint_limit_stiffness", 2000.0)),
                           joint_limit_damping=float(getattr(px, "joint_limit_damping", 20.0)),
                           max_contacts_per_chain=int(getattr(px, "max_contacts_per_chain", 0) or 0))
        c.gravity[0], c.gravity[1], c.gravity[2] = g.x, g.y, g.z
        asset = getattr(self, "asset", None)
        if asset is not None and getattr(asset.options, "disable_gravity", False):
            # AssetOptions.disable_gravity (tasks/hound_arm.py:212): the sim holds one articulation type, so its bodies are all of them
            c.gravity[0] = c.gravity[1] = c.gravity[2] = 0.0
        # AssetOptions.max_linear_velocity / max_angular_velocity (Isaac Gym defaults 1000 m/s, 64 rad/s; no hot-path task changes them)
        opts = getattr(asset, "options", None)
        c.max_linear_velocity = float(getattr(opts, "max_linear_velocity", 1000.0) or 0.0)
        c.max_angular_velocity = float(getattr(opts, "max_angular_velocity", 64.0) or 0.0)
        sc = getattr(px, "self_collision", None)
        c.self_collision = int(bool(getattr(self, "self_collision", False) if sc is None else sc))
        return c

    def stream(self):
        import torch

        return C.c_void_p(torch.cuda.current_stream(self.compute_device).cuda_stream)


def _props_to_struct(art, props) -> _abi.DofProps:
    p = _abi.default_dof_props(art)
    if props is not None:
        big = 3.0e38
        for d in range(art.num_dofs):
            p.drive_mode[d] = int(props["driveMode"][d])
            p.stiffness[d] = float(props["stiffness"][d])
            p.damping[d] = float(props["damping"][d])
            p.effort[d] = float(props["effort"][d])
            p.velocity[d] = float(props["velocity"][d])
            if bool(props["hasLimits"][d]):
                p.lower[d], p.upper[d] = float(props["lower"][d]), float(props["upper"][d])
            else:
                p.lower[d], p.upper[d] = -big, big
    return p


class Gym:
    """The object ``gymapi.acquire_gym()`` returns."""

    # ---------------- lifecycle ----------------
    def create_sim(self, compute_device=0, graphics_device=-1, type=SIM_PHYSX, params: Optional[SimParams] = None):
        """vec_task.py:63,337.  Returns None on failure (the caller prints and quits, vec_task.py:338-340)."""
        params = params or SimParams()
        if type != SIM_PHYSX:
            raise ValueError("only SIM_PHYSX-style rigid-body simulation is provided")
        if params.up_axis != UP_AXIS_Z:
            raise _lib.B2GError("the B200 engine is z-up: set sim.up_axis to 'z' (every hot-path task config does)")
        if not (params.use_gpu_pipeline and getattr(params.physx, "use_gpu", True)):
            raise _lib.B2GError("sim_device=cpu / pipeline=cpu requested: this build has no CPU simulation path "
                                "(hand-written CUDA only). Use sim_device=cuda:N pipeline=gpu.")
        sim = Sim(compute_device, graphics_device, type, params)
        lib = _lib.load()
        cp = sim.c_params()
        _lib.check(lib.b2g_sim_create(int(compute_device), C.byref(cp), C.byref(sim.handle)), "create_sim")
        return sim

    def destroy_sim(self, sim: Sim):
        if sim.handle:
            _lib.load().b2g_sim_destroy(sim.handle)
            sim.handle = C.c_void_p()

    def prepare_sim(self, sim: Sim):
        """vec_task.py:262."""
        if sim.asset is None:
            raise _lib.B2GError("prepare_sim called before any actor was created")
        if sim.prepared:        # idempotent: tasks that acquire tensors inside _create_envs (tasks/hound_arm.py:288) prepare early
            return True
        lib = _lib.load()
        art = sim.asset.art
        sim.model_struct = _abi.pack_model(art)
        props = _props_to_struct(art, sim.dof_props)
        pose = sim.actor_pose or Transform()
        pose7 = (C.c_float * 7)(pose.p.x, pose.p.y, pose.p.z, pose.r.x, pose.r.y, pose.r.z, pose.r.w)
        n = len(sim.envs)
        _lib.check(lib.b2g_sim_add_articulation(sim.handle, C.byref(sim.model_struct), C.byref(props), n, pose7,
                                                float(sim.env_spacing), int(sim.num_per_row)), "create_actor")
        # what the actor brings to the sim parameters: AssetOptions.disable_gravity and velocity limits, create_actor's collision filter
        cp = sim.c_params()
        _lib.check(lib.b2g_sim_set_params(sim.handle, C.byref(cp)), "set_sim_params")
        _lib.check(lib.b2g_sim_prepare(sim.handle), "prepare_sim")
        sim.prepared = True
        import torch

        if len(set(sim.actor_poses)) > 1:       # per-env start poses (terrain tasks place robots on their tiles)
            root = self._tensor(sim, _abi.T_ROOT_STATE)
            root[:, :7] = torch.tensor(sim.actor_poses, dtype=torch.float32, device=root.device)
        if any(abs(f - 1.0) > 0 for f in sim.env_friction):

            fr = self._tensor(sim, _abi.T_FRICTION)
            fr.copy_(torch.tensor(sim.env_friction, dtype=torch.float32, device=fr.device))
        return True

    def simulate(self, sim: Sim):
        """vec_task.py:382, anymal_terrain.py:448."""
        _lib.check(_lib.load().b2g_sim_simulate(sim.handle, sim.stream()), "simulate")
        sim.frame_count += 1

    def fetch_results(self, sim: Sim, wait_for_latest_sim_step=True):
        return None

    def step_graphics(self, sim):
        return None

    def get_frame_count(self, sim: Sim):
        return sim.frame_count

    def get_sim_params(self, sim: Sim):
        return sim.params

    def set_sim_params(self, sim: Sim, params: SimParams):
        sim.params = params
        cp = sim.c_params()
        _lib.check(_lib.load().b2g_sim_set_params(sim.handle, C.byref(cp)), "set_sim_params")

    def get_sim_time(self, sim: Sim):
        return sim.frame_count * sim.params.dt

    # ---------------- scene ----------------
    def add_ground(self, sim: Sim, plane: PlaneParams):
        """tasks/anymal.py:159-164."""
        n = plane.normal
        if abs(n.z - 1.0) > 1e-6 or abs(n.x) > 1e-6 or abs(n.y) > 1e-6:
            raise _lib.B2GError("ground plane normal must be +z")
        _lib.check(_lib.load().b2g_sim_add_ground(sim.handle, plane.static_friction, plane.dynamic_friction, plane.restitution), "add_ground")

    def add_heightfield(self, sim: Sim, heights, params: HeightFieldParams):
        """Isaac Gym ``add_heightfield``: int16 samples (row-major, rows along x)."""
        h = np.ascontiguousarray(np.asarray(heights, dtype=np.int16).reshape(params.nbRows, params.nbColumns))
        if abs(params.row_scale - params.column_scale) > 1e-9:
            raise _lib.B2GError("heightfield needs square cells")
        hf = _abi.Heightfield(rows=params.nbRows, cols=params.nbColumns, horizontal_scale=params.row_scale,
                              vertical_scale=params.vertical_scale, origin_x=params.transform.p.x, origin_y=params.transform.p.y,
                              friction=params.dynamic_friction, restitution=params.restitution)
        _lib.check(_lib.load().b2g_sim_add_heightfield(sim.handle, C.byref(hf), h.ctypes.data_as(C.c_void_p)), "add_heightfield")
        sim.heightfield = (hf, h)

    def add_triangle_mesh(self, sim: Sim, vertices, triangles, params: TriangleMeshParams):
        """tasks/anymal_terrain.py:196-209.  The terrain meshes the reference builds come from a regular
        height grid (``convert_heightfield_to_trimesh``); the grid is recovered from the vertices and stored as a
        heightfield (vertical walls added by the slope threshold collapse to steep cells -- stated deviation)."""
        v = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)
        xs, ys = np.unique(np.round(v[:, 0], 5)), np.unique(np.round(v[:, 1], 5))
        if len(xs) * len(ys) != len(v) or len(xs) < 2 or len(ys) < 2:
            raise _lib.B2GError("add_triangle_mesh supports gridded terrain meshes only (vertices on a regular x-y grid)")
        hs = float(xs[1] - xs[0])
        if abs((ys[1] - ys[0]) - hs) > 1e-4:
            raise _lib.B2GError("terrain grid must have square cells")
        ix = np.rint((v[:, 0] - xs[0]) / hs).astype(np.int64)
        iy = np.rint((v[:, 1] - ys[0]) / hs).astype(np.int64)
        z = np.zeros((len(xs), len(ys)), dtype=np.float32)
        z[ix, iy] = v[:, 2]
        vs = 0.0005
        hp = HeightFieldParams()
        hp.nbRows, hp.nbColumns, hp.row_scale, hp.column_scale, hp.vertical_scale = len(xs), len(ys), hs, hs, vs
        hp.transform.p.x = float(xs[0]) + params.transform.p.x
        hp.transform.p.y = float(ys[0]) + params.transform.p.y
        hp.static_friction, hp.dynamic_friction, hp.restitution = params.static_friction, params.dynamic_friction, params.restitution
        raw = np.clip(np.rint((z + params.transform.p.z) / vs), -32768, 32767).astype(np.int16)
        self.add_heightfield(sim, raw, hp)

    def load_asset(self, sim: Sim, rootpath: str, filename: str, options: Optional[AssetOptions] = None):
        """tasks/anymal.py:183.  Compiles the URDF if it is on disk, else falls back to the in-tree compiled model."""
        options = options or AssetOptions()
        path = os.path.join(rootpath, filename)
        if os.path.isfile(path):
            art = _urdf.compile_urdf(path, options)
        else:
            comp = _store.find_compiled(filename, options, rootpath)
            if comp is None:
                raise FileNotFoundError(f"asset {path} not found and no compiled model for it under {_store.COMPILED_DIR}")
            art = _store.load_articulation(comp)
            art.fixed_base = bool(options.fix_base_link)
            art.armature = np.full(art.num_dofs, float(options.armature))
        return Asset(art, options, path)

    def get_asset_dof_count(self, asset: Asset):
        return asset.art.num_dofs

    def get_asset_rigid_body_count(self, asset: Asset):
        return asset.art.num_bodies

    def get_asset_rigid_body_names(self, asset: Asset):
        return list(asset.art.body_names)

    def get_asset_dof_names(self, asset: Asset):
        return list(asset.art.dof_names)

    def get_asset_rigid_body_dict(self, asset: Asset):
        return {n: i for i, n in enumerate(asset.art.body_names)}

    def get_asset_dof_dict(self, asset: Asset):
        return {n: i for i, n in enumerate(asset.art.dof_names)}

    def get_asset_rigid_shape_count(self, asset: Asset):
        return asset.n_shapes

    def find_asset_rigid_body_index(self, asset: Asset, name: str):
        return asset.art.body_names.index(name) if name in asset.art.body_names else -1

    def find_asset_dof_index(self, asset: Asset, name: str):
        return asset.art.dof_names.index(name) if name in asset.art.dof_names else -1

    def get_asset_dof_properties(self, asset: Asset):
        """tasks/anymal.py:198: numpy structured array, one row per DOF."""
        art = asset.art
        p = np.zeros(art.num_dofs, dtype=DOF_PROPS_DTYPE)
        p["hasLimits"] = art.has_limits
        p["lower"] = np.where(art.has_limits, art.lower, 0.0)
        p["upper"] = np.where(art.has_limits, art.upper, 0.0)
        p["driveMode"] = int(asset.options.default_dof_drive_mode)
        p["velocity"] = art.velocity
        p["effort"] = art.effort
        p["stiffness"] = 0.0
        p["damping"] = art.damping
        p["friction"] = art.friction
        p["armature"] = art.armature
        return p

    def get_asset_rigid_shape_properties(self, asset: Asset):
        """tasks/anymal_terrain.py:236."""
        out = []
        for _ in range(asset.n_shapes):
            r = RigidShapeProperties()
            r.friction = asset.shape_friction
            out.append(r)
        return out

    def set_asset_rigid_shape_properties(self, asset: Asset, props):
        """tasks/anymal_terrain.py:281: per-env friction buckets (the value in force at create_actor time sticks)."""
        if len(props) > 0:
            asset.shape_friction = float(props[0].friction)

    def create_env(self, sim: Sim, lower: Vec3, upper: Vec3, num_per_row: int):
        """tasks/anymal.py:212."""
        if sim.prepared:
            raise _lib.B2GError("create_env after prepare_sim")
        env = Env(sim, len(sim.envs))
        sim.envs.append(env)
        sim.env_spacing = float(upper.x - lower.x) if upper is not None and lower is not None else 0.0
        sim.num_per_row = int(num_per_row)
        return env

    def create_actor(self, env: Env, asset: Asset, pose: Transform, name: str = "", group: int = -1, filter: int = -1, segmentationId: int = 0):
        """tasks/anymal.py:213.  One actor (articulation type) per env.  `filter` is Isaac Gym's collision-filter mask: bodies whose masks share
        a bit do not collide, so 0 (the rough-terrain tasks, tasks/anymal_terrain.py:282) switches self-collision ON and 1 (the flat tasks)
        off.  On -> b2g_sim_params.self_collision: the links that do not hang off the base directly collide with the base's box
        (link-link pairs are not modelled, DESIGN.md deviations); `sim.physx.self_collision` in the yaml overrides it."""
        sim = env.sim
        sim.self_collision = int(filter) == 0
        if sim.asset is None:
            sim.asset, sim.actor_name, sim.actor_pose = asset, name, pose
        elif sim.asset is not asset:
            raise _lib.B2GError("one articulation type per sim is supported")
        if env.actors:
            raise _lib.B2GError("one actor per env is supported")
        env.actors.append(name)
        sim.env_friction.append(float(asset.shape_friction))
        sim.actor_poses.append((pose.p.x, pose.p.y, pose.p.z, pose.r.x, pose.r.y, pose.r.z, pose.r.w))
        return 0

    # ---- actor bookkeeping used by the reference's domain-randomisation code (tasks/base/vec_task.py:583, utils/dr_utils.py:233-237)
    def get_actor_count(self, env: Env):
        return len(env.actors)

    def get_actor_handle(self, env: Env, index: int):
        if not 0 <= int(index) < len(env.actors):
            raise _lib.B2GError(f"actor index {index} out of range")
        return int(index)

    def get_actor_name(self, env: Env, actor: int):
        return env.actors[int(actor)]

    def find_actor_handle(self, env: Env, name: str):
        return env.actors.index(name) if name in env.actors else -1

    def get_actor_rigid_shape_count(self, env: Env, actor: int):
        return env.sim.asset.n_shapes

    def get_actor_tendon_properties(self, env: Env, actor: int):
        return []           # no tendons in the articulation model of this engine

    def set_actor_tendon_properties(self, env: Env, actor: int, props):
        return True

    def set_rigid_body_color(self, env: Env, actor: int, body: int, mesh_type=None, color=None):
        return None         # rendering only (headless engine)

    def set_actor_scale(self, env: Env, actor: int, scale: float):
        """Isaac Gym returns False when the actor could not be rescaled; geometric scaling of a compiled articulation is not supported
        here, so a request for anything but 1.0 reports failure the same way (vec_task.py:786 ignores the result)."""
        return abs(float(scale) - 1.0) < 1e-12

    def get_actor_dof_properties(self, env: Env, actor: int):
        sim = env.sim
        return sim.dof_props.copy() if sim.dof_props is not None else self.get_asset_dof_properties(sim.asset)

    def set_actor_dof_properties(self, env: Env, actor: int, props):
        """tasks/anymal.py:214.  Drive properties are shared by all envs (the reference passes the same array)."""
        sim = env.sim
        sim.dof_props = np.array(props, dtype=DOF_PROPS_DTYPE, copy=True)
        if sim.prepared:
            st = _props_to_struct(sim.asset.art, sim.dof_props)
            _lib.check(_lib.load().b2g_sim_set_dof_props(sim.handle, C.byref(st)), "set_actor_dof_properties")
        return True

    def enable_actor_dof_force_sensors(self, env: Env, actor: int):
        return True

    def get_actor_rigid_body_properties(self, env: Env, actor: int):
        art = env.sim.asset.art
        out = []
        for b in range(art.num_bodies):
            l = int(art.body_link[b])
            own = [bb for bb in range(art.num_bodies) if art.body_link[bb] == l]
            out.append(RigidBodyProperties(mass=float(art.mass[l]) if own[0] == b else 0.0, com=Vec3(*art.com[l])))
        return out

    def set_actor_rigid_body_properties(self, env, actor, props, recomputeInertia=False):
        return True

    def get_actor_rigid_shape_properties(self, env: Env, actor: int):
        return self.get_asset_rigid_shape_properties(env.sim.asset)

    def set_actor_rigid_shape_properties(self, env, actor, props):
        return True

    def find_actor_rigid_body_handle(self, env: Env, actor: int, name: str):
        """tasks/anymal.py:220-224: body index inside the actor."""
        names = env.sim.asset.art.body_names
        return names.index(name) if name in names else -1

    def find_actor_dof_handle(self, env: Env, actor: int, name: str):
        names = env.sim.asset.art.dof_names
        return names.index(name) if name in names else -1

    def get_actor_rigid_body_names(self, env: Env, actor: int):
        return list(env.sim.asset.art.body_names)

    def get_actor_dof_names(self, env: Env, actor: int):
        return list(env.sim.asset.art.dof_names)

    def get_actor_joint_dict(self, env: Env, actor: int):
        """tasks/useful_hound.py:450."""
        return dict(env.sim.asset.art.joint_dict)

    def get_actor_rigid_body_count(self, env: Env, actor: int):
        return env.sim.asset.art.num_bodies

    def get_actor_dof_count(self, env: Env, actor: int):
        return env.sim.asset.art.num_dofs

    def get_sim_dof_count(self, sim: Sim):
        return sim.asset.art.num_dofs * len(sim.envs)

    def get_sim_actor_count(self, sim: Sim):
        return len(sim.envs)

    def get_sim_rigid_body_count(self, sim: Sim):
        return sim.asset.art.num_bodies * len(sim.envs)

    def get_env_count(self, sim: Sim):
        return len(sim.envs)

    def get_env(self, sim: Sim, i: int):
        return sim.envs[i]

    # ---------------- tensors ----------------
    def _desc(self, sim: Sim, kind: int) -> _abi.TensorDesc:
        if not sim.prepared:
            # Isaac Gym hands out tensor descriptors before prepare_sim (the reference's arm tasks acquire theirs inside _create_envs,
            # tasks/manipulator.py:300-323, and VecTask.__init__ prepares afterwards): every actor exists by then, so prepare now --
            # prepare_sim is idempotent
            if sim.asset is None or not sim.envs:
                raise _lib.B2GError("acquire_*_tensor before any actor was created")
            self.prepare_sim(sim)
        d = _abi.TensorDesc()
        _lib.check(_lib.load().b2g_sim_tensor(sim.handle, kind, C.byref(d)), "acquire tensor")
        return d

    def _tensor(self, sim: Sim, kind: int):
        return _lib.desc_to_torch(self._desc(sim, kind))

    def acquire_actor_root_state_tensor(self, sim):
        return GymTensor(self._desc(sim, _abi.T_ROOT_STATE))

    def acquire_dof_state_tensor(self, sim):
        return GymTensor(self._desc(sim, _abi.T_DOF_STATE))

    def acquire_net_contact_force_tensor(self, sim):
        return GymTensor(self._desc(sim, _abi.T_NET_CONTACT))

    def acquire_dof_force_tensor(self, sim):
        return GymTensor(self._desc(sim, _abi.T_DOF_FORCE))

    def acquire_rigid_body_state_tensor(self, sim):
        return GymTensor(self._desc(sim, _abi.T_RIGID_BODY_STATE))

    def acquire_jacobian_tensor(self, sim, actor_name):
        return GymTensor(self._desc(sim, _abi.T_JACOBIAN))

    def acquire_mass_matrix_tensor(self, sim, actor_name):
        return GymTensor(self._desc(sim, _abi.T_MASS_MATRIX))

    def _refresh(self, sim, kind):
        _lib.check(_lib.load().b2g_sim_refresh(sim.handle, kind, sim.stream()), "refresh tensor")
        return True

    def refresh_actor_root_state_tensor(self, sim):
        return self._refresh(sim, _abi.T_ROOT_STATE)

    def refresh_dof_state_tensor(self, sim):
        return self._refresh(sim, _abi.T_DOF_STATE)

    def refresh_net_contact_force_tensor(self, sim):
        return self._refresh(sim, _abi.T_NET_CONTACT)

    def refresh_dof_force_tensor(self, sim):
        return self._refresh(sim, _abi.T_DOF_FORCE)

    def refresh_rigid_body_state_tensor(self, sim):
        return self._refresh(sim, _abi.T_RIGID_BODY_STATE)

    def refresh_jacobian_tensors(self, sim):
        return self._refresh(sim, _abi.T_JACOBIAN)

    def refresh_mass_matrix_tensors(self, sim):
        return self._refresh(sim, _abi.T_MASS_MATRIX)

    def _set_full(self, sim, kind, t: GymTensor):
        _lib.check(_lib.load().b2g_sim_set_tensor(sim.handle, kind, C.c_void_p(t.data_ptr), sim.stream()), "set tensor")
        return True

    def _set_indexed(self, sim, kind, t: GymTensor, idx: GymTensor, n: int):
        _lib.check(_lib.load().b2g_sim_set_indexed(sim.handle, kind, C.c_void_p(t.data_ptr), C.c_void_p(idx.data_ptr), int(n), sim.stream()),
                   "set tensor indexed")
        return True

    def set_dof_position_target_tensor(self, sim, t):
        return self._set_full(sim, _abi.T_DOF_TARGET, t)

    def set_dof_velocity_target_tensor(self, sim, t):
        return self._set_full(sim, _abi.T_DOF_TARGET, t)

    def set_dof_actuation_force_tensor(self, sim, t):
        return self._set_full(sim, _abi.T_DOF_ACTUATION, t)

    def set_actor_root_state_tensor(self, sim, t):
        return self._set_full(sim, _abi.T_ROOT_STATE, t)

    def set_dof_state_tensor(self, sim, t):
        return self._set_full(sim, _abi.T_DOF_STATE, t)

    def set_actor_root_state_tensor_indexed(self, sim, t, idx, n):
        return self._set_indexed(sim, _abi.T_ROOT_STATE, t, idx, n)

    def set_dof_state_tensor_indexed(self, sim, t, idx, n):
        return self._set_indexed(sim, _abi.T_DOF_STATE, t, idx, n)

    def set_dof_position_target_tensor_indexed(self, sim, t, idx, n):
        return self._set_indexed(sim, _abi.T_DOF_TARGET, t, idx, n)

    def set_dof_actuation_force_tensor_indexed(self, sim, t, idx, n):
        return self._set_indexed(sim, _abi.T_DOF_ACTUATION, t, idx, n)

    # ---------------- viewer (headless only) ----------------
    def create_viewer(self, sim, props=None):
        return None

    def subscribe_viewer_keyboard_event(self, viewer, key, name):
        return None

    def query_viewer_has_closed(self, viewer):
        return False

    def query_viewer_action_events(self, viewer):
        return []

    def viewer_camera_look_at(self, viewer, env, pos, target):
        return None

    def draw_viewer(self, viewer, sim, render_collision=True):
        return None

    def poll_viewer_events(self, viewer):
        return None

    def sync_frame_time(self, sim):
        return None

    def clear_lines(self, viewer):
        return None

    def write_viewer_image_to_file(self, viewer, path):
        return None


class CameraProperties:
    pass


_GYM = None


def acquire_gym() -> Gym:
    """vec_task.py:247."""
    global _GYM
    if _GYM is None:
        _GYM = Gym()
    return _GYM
