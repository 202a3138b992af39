"""URDF -> flat articulation description (the "model compiler").

Replaces what Isaac Gym's ``gym.load_asset`` does for the hot-path tasks (call sites in the
reference: ``tasks/anymal.py:170-203``, ``tasks/anymal_terrain.py:218-236``,
``tasks/hound.py:170-199``, ``tasks/useful_hound.py:316-388``, ``tasks/cartpole.py:86-88``):

* parse links / joints / inertials / collision primitives (standard URDF semantics: child frame =
  parent * T(xyz) * Rz(yaw) Ry(pitch) Rx(roll), then the joint motion about/along ``axis``);
* order bodies and DOFs depth-first with children visited in ASCII-sorted child-link-name order,
  which reproduces the orderings the reference tasks assume (LF,LH,RF,RH for
  ANYmal; FL,FR,RL,RR,link1.. for Hound);
* ``collapse_fixed_joints``: fold every link that hangs on a fixed joint into its parent body (mass,
  centre of mass, inertia with rotation + parallel-axis shift, collision shapes);
* without collapsing, fixed children remain *API bodies* (they get rows in the rigid-body / contact
  force tensors, e.g. Hound's 4 foot bodies) but ride rigidly on their parent's *dynamic link*;
* collision primitives become contact spheres (point + radius): sphere -> 1, capsule/cylinder -> 2,
  box -> 8 corners, mesh -> 8 corners of its bounding box (stated deviation).

The dynamics engine supports "star" topologies only: one root body plus K serial chains.  Every robot
on the hot path (ANYmal-C, Hound, Hound+arm, Cartpole) has that shape; anything else raises.
"""
from __future__ import annotations

import os
import struct
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

JOINT_REVOLUTE = 0
JOINT_PRISMATIC = 1


# ----------------------------------------------------------------------------------------------
# small rotation helpers (float64 numpy)
# ----------------------------------------------------------------------------------------------
def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def mat_to_quat(m):
    """3x3 rotation -> quaternion (x, y, z, w), w >= 0."""
    m = np.asarray(m, dtype=np.float64)
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s])
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def axis_angle_mat(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * k + (1 - np.cos(angle)) * (k @ k)


def _lenient_float(s: str) -> float:
    """``float`` that tolerates the malformed ``izz="0.0.0000001"`` in the Hound URDFs (strtod-like:
    parse the longest valid prefix)."""
    s = s.strip()
    try:
        return float(s)
    except ValueError:
        for end in range(len(s) - 1, 0, -1):
            try:
                return float(s[:end])
            except ValueError:
                continue
        raise


def _vec(s: Optional[str], default):
    if s is None:
        return np.array(default, dtype=np.float64)
    return np.array([_lenient_float(t) for t in s.split()], dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# raw URDF
# ----------------------------------------------------------------------------------------------
@dataclass
class Shape:
    kind: str                      # "box" | "sphere" | "cylinder" | "capsule" | "mesh"
    pos: np.ndarray                # in the owning link frame
    rot: np.ndarray                # 3x3
    size: np.ndarray               # box: full extents; sphere: [r]; cylinder/capsule: [r, length]; mesh: aabb lo+hi (6)
    body: str = ""                 # API body (URDF link) the shape belongs to
    hull: object = None            # mesh: (volume, centroid, unit-density inertia about it) of the convex hull, or None


@dataclass
class RawLink:
    name: str
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))   # about the COM, link axes
    has_inertia: bool = False
    shapes: List[Shape] = field(default_factory=list)


@dataclass
class RawJoint:
    name: str
    kind: str
    parent: str
    child: str
    pos: np.ndarray
    rot: np.ndarray
    axis: np.ndarray
    lower: float = 0.0
    upper: float = 0.0
    effort: float = 0.0
    velocity: float = 0.0
    damping: float = 0.0
    friction: float = 0.0
    has_limits: bool = False


def _mesh_points(path, scale):
    """Vertices of a (binary or ASCII) STL or a Wavefront OBJ, scaled."""
    with open(path, "rb") as f:
        data = f.read()
    if path.lower().endswith(".obj"):
        vals = []
        for line in data.decode("ascii", "ignore").splitlines():
            t = line.split()
            if len(t) >= 4 and t[0] == "v":
                vals.append([float(t[1]), float(t[2]), float(t[3])])
        return np.array(vals, dtype=np.float64).reshape(-1, 3) * scale
    pts = None
    if len(data) >= 84:
        ntri = struct.unpack_from("<I", data, 80)[0]
        if 84 + ntri * 50 == len(data):
            rec = np.frombuffer(data, dtype=np.dtype([("n", "<3f4"), ("v", "<9f4"), ("a", "<u2")]), count=ntri, offset=84)
            pts = rec["v"].reshape(-1, 3).astype(np.float64)
    if pts is None:
        vals = []
        for line in data.decode("ascii", "ignore").splitlines():
            t = line.split()
            if len(t) == 4 and t[0] == "vertex":
                vals.append([float(t[1]), float(t[2]), float(t[3])])
        pts = np.array(vals, dtype=np.float64)
    return pts * scale


def _stl_aabb(path, scale):
    """Axis-aligned bounds of a mesh file, scaled."""
    pts = _mesh_points(path, scale)
    return pts.min(0), pts.max(0)


def _hull_mass_properties(pts):
    """(volume, centroid, inertia about the centroid for unit density) of the convex hull of ``pts`` -- PhysX cooks collision meshes
    into convex hulls and derives a link's mass properties from them when the URDF has no <inertial> (franka_panda_manipulator.urdf).
    None when the hull is degenerate or scipy is unavailable (the caller falls back to the bounding box)."""
    try:
        from scipy.spatial import ConvexHull

        hull = ConvexHull(pts)
    except Exception:
        return None
    vol, first, second = 0.0, np.zeros(3), np.zeros((3, 3))
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = pts[simplex[0]], pts[simplex[1]], pts[simplex[2]]
        if np.dot(np.cross(b - a, c - a), eq[:3]) < 0.0:
            b, c = c, b
        v = np.dot(a, np.cross(b, c)) / 6.0      # signed volume of the tetrahedron (origin, a, b, c)
        s = a + b + c
        vol += v
        first += v * s / 4.0
        second += v / 20.0 * (np.outer(s, s) + np.outer(a, a) + np.outer(b, b) + np.outer(c, c))
    if vol <= 1e-12:
        return None
    com = first / vol
    i0 = np.trace(second) * np.eye(3) - second
    return vol, com, i0 - vol * (np.dot(com, com) * np.eye(3) - np.outer(com, com))


def _parse_xml_lenient(path: str):
    """``ET.parse`` that, like the forgiving XML readers URDF importers use, stops at the end of the root element: the reference's
    ``franka_panda_manipulator.urdf`` closes ``</robot>`` after ``panda_joint7`` and then carries the hand and the fingers inside a
    comment that itself contains comments -- not XML, and outside the document anyway.  8 links and 7 revolute joints remain, which is
    what ``tasks/manipulator.py`` is written against (7-entry gain vectors indexed over ``num_franka_dofs``, :216-234)."""
    text = open(path, "r", encoding="utf-8", errors="replace").read()
    try:
        return ET.fromstring(text)
    except ET.ParseError:
        m = re.search(r"<\s*([A-Za-z_][\w:.-]*)", re.sub(r"<\?.*?\?>|<!--.*?-->", "", text, flags=re.S))
        end = text.find(f"</{m.group(1)}>") if m else -1
        if end < 0:
            raise
        return ET.fromstring(text[:end + len(m.group(1)) + 3])


def parse_urdf(path: str):
    root = _parse_xml_lenient(path)
    base_dir = os.path.dirname(os.path.abspath(path))
    links: Dict[str, RawLink] = {}
    for le in root.findall("link"):
        lk = RawLink(name=le.get("name"))
        ine = le.find("inertial")
        if ine is not None:
            o = ine.find("origin")
            r_i = np.eye(3)
            if o is not None:
                lk.com = _vec(o.get("xyz"), [0, 0, 0])
                r_i = rpy_to_mat(_vec(o.get("rpy"), [0, 0, 0]))
            m = ine.find("mass")
            lk.mass = _lenient_float(m.get("value")) if m is not None else 0.0
            it = ine.find("inertia")
            if it is not None:
                g = lambda k: _lenient_float(it.get(k, "0"))
                i_loc = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
                lk.inertia = r_i @ i_loc @ r_i.T
                lk.has_inertia = True
        for ce in le.findall("collision"):
            o = ce.find("origin")
            pos = _vec(o.get("xyz"), [0, 0, 0]) if o is not None else np.zeros(3)
            rot = rpy_to_mat(_vec(o.get("rpy"), [0, 0, 0])) if o is not None else np.eye(3)
            ge = ce.find("geometry")
            if ge is None or len(ge) == 0:
                continue
            g = ge[0]
            if g.tag == "box":
                sh = Shape("box", pos, rot, _vec(g.get("size"), [0, 0, 0]), lk.name)
            elif g.tag == "sphere":
                sh = Shape("sphere", pos, rot, np.array([_lenient_float(g.get("radius"))]), lk.name)
            elif g.tag == "cylinder":
                sh = Shape("cylinder", pos, rot, np.array([_lenient_float(g.get("radius")), _lenient_float(g.get("length"))]), lk.name)
            elif g.tag == "mesh":
                fn = g.get("filename", "")
                scale = _vec(g.get("scale"), [1, 1, 1])
                for prefix in ("package://", "file://"):
                    if fn.startswith(prefix):
                        fn = fn[len(prefix):]
                cand = [os.path.join(base_dir, fn), os.path.join(base_dir, os.path.basename(fn))]
                up = base_dir
                for _ in range(4):      # package://<pkg>/... : the package directory is an ancestor of the URDF's directory
                    up = os.path.dirname(up)
                    cand.append(os.path.join(up, fn))
                mp = next((c for c in cand if os.path.isfile(c)), None)
                if mp is None or not mp.lower().endswith((".stl", ".obj")):
                    continue  # visual-only formats (.dae) carry no collision here
                pts = _mesh_points(mp, scale)
                if len(pts) == 0:
                    continue
                sh = Shape("mesh", pos, rot, np.concatenate([pts.min(0), pts.max(0)]), lk.name)
                sh.hull = _hull_mass_properties(pts)
            else:
                continue
            lk.shapes.append(sh)
        links[lk.name] = lk
    joints: List[RawJoint] = []
    for je in root.findall("joint"):
        o = je.find("origin")
        pos = _vec(o.get("xyz"), [0, 0, 0]) if o is not None else np.zeros(3)
        rot = rpy_to_mat(_vec(o.get("rpy"), [0, 0, 0])) if o is not None else np.eye(3)
        ax = je.find("axis")
        axis = _vec(ax.get("xyz"), [1, 0, 0]) if ax is not None else np.array([1.0, 0, 0])
        j = RawJoint(je.get("name"), je.get("type"), je.find("parent").get("link"), je.find("child").get("link"), pos, rot, axis)
        lim = je.find("limit")
        if lim is not None:
            j.effort = _lenient_float(lim.get("effort", "0"))
            j.velocity = _lenient_float(lim.get("velocity", "0"))
            if lim.get("lower") is not None or lim.get("upper") is not None:
                j.lower = _lenient_float(lim.get("lower", "0"))
                j.upper = _lenient_float(lim.get("upper", "0"))
                j.has_limits = j.kind != "continuous"
        dyn = je.find("dynamics")
        if dyn is not None:
            j.damping = _lenient_float(dyn.get("damping", "0"))
            j.friction = _lenient_float(dyn.get("friction", "0"))
        joints.append(j)
    return root.get("name"), links, joints


# ----------------------------------------------------------------------------------------------
# compiled articulation
# ----------------------------------------------------------------------------------------------
@dataclass
class AssetOptions:
    """Subset of ``gymapi.AssetOptions`` the reference sets (``tasks/anymal.py:170-181``)."""
    default_dof_drive_mode: int = 0
    collapse_fixed_joints: bool = False
    replace_cylinder_with_capsule: bool = False
    flip_visual_attachments: bool = False
    fix_base_link: bool = False
    density: float = 1000.0
    angular_damping: float = 0.0
    linear_damping: float = 0.0
    max_angular_velocity: float = 64.0
    max_linear_velocity: float = 1000.0
    armature: float = 0.0
    thickness: float = 0.02
    disable_gravity: bool = False
    override_com: bool = False
    override_inertia: bool = False
    use_mesh_materials: bool = False
    vhacd_enabled: bool = False


@dataclass
class Articulation:
    name: str
    fixed_base: bool
    # API bodies (rows of the rigid-body / net-contact-force tensors), depth-first order
    body_names: List[str]
    body_link: np.ndarray          # (nb,) dynamic link index the body rides on (0 = root link)
    body_pos: np.ndarray           # (nb,3) body frame in its link frame
    body_quat: np.ndarray          # (nb,4) xyzw
    # dynamic links: link 0 = root, link 1+d = child link of DOF d
    dof_names: List[str]
    link_names: List[str]
    link_parent: np.ndarray        # (nl,) parent link (-1 for root)
    joint_type: np.ndarray         # (nd,)
    joint_pos: np.ndarray          # (nd,3) joint frame origin in the parent link frame
    joint_quat: np.ndarray         # (nd,4) joint frame rotation in the parent link frame (q = 0)
    joint_axis: np.ndarray         # (nd,3) unit axis in the child link frame
    mass: np.ndarray               # (nl,)
    com: np.ndarray                # (nl,3)
    inertia: np.ndarray            # (nl,3,3) about the COM, link axes
    lower: np.ndarray              # (nd,)
    upper: np.ndarray
    has_limits: np.ndarray         # (nd,) bool
    effort: np.ndarray
    velocity: np.ndarray
    damping: np.ndarray
    friction: np.ndarray
    armature: np.ndarray
    # star topology
    chain_start: np.ndarray        # (nc,) first DOF of each chain
    chain_len: np.ndarray          # (nc,)
    # contact spheres
    cp_link: np.ndarray            # (np,)
    cp_body: np.ndarray            # (np,)
    cp_pos: np.ndarray             # (np,3) in the link frame
    cp_radius: np.ndarray          # (np,)
    joint_dict: Dict[str, int] = field(default_factory=dict)   # joint name (incl. fixed) -> joint index

    @property
    def num_dofs(self):
        return len(self.dof_names)

    @property
    def num_bodies(self):
        return len(self.body_names)

    @property
    def num_links(self):
        return len(self.link_names)

    @property
    def total_mass(self):
        return float(self.mass.sum())


def _shape_points(sh: Shape, opts: AssetOptions):
    """Contact spheres (centre in the shape's own frame, radius) for one primitive."""
    if sh.kind == "sphere":
        return [(np.zeros(3), float(sh.size[0]))]
    if sh.kind == "cylinder":
        r, length = float(sh.size[0]), float(sh.size[1])
        if opts.replace_cylinder_with_capsule:
            half = 0.5 * length          # capsule: cylindrical part keeps the URDF length, caps added
        else:
            half = max(0.5 * length - r, 0.0)   # inscribed capsule stands in for the flat-ended cylinder
        if half <= 1e-9:
            return [(np.zeros(3), r)]
        return [(np.array([0, 0, half]), r), (np.array([0, 0, -half]), r)]
    if sh.kind == "box":
        hx, hy, hz = 0.5 * sh.size
        return [(np.array([sx * hx, sy * hy, sz * hz]), 0.0) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    if sh.kind == "mesh":
        lo, hi = sh.size[:3], sh.size[3:]
        return [(np.array([x, y, z]), 0.0) for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]
    return []


def _shape_inertia(sh: Shape, density: float):
    """(mass, com, inertia about com) of a primitive in the owning link frame, for links whose URDF
    gives no inertia (Isaac Gym derives it from the collision shapes; used by cartpole.urdf)."""
    if sh.kind == "box":
        x, y, z = sh.size
        m = density * x * y * z
        i = m / 12.0 * np.diag([y * y + z * z, x * x + z * z, x * x + y * y])
    elif sh.kind == "sphere":
        r = sh.size[0]
        m = density * 4.0 / 3.0 * np.pi * r ** 3
        i = 0.4 * m * r * r * np.eye(3)
    elif sh.kind in ("cylinder", "capsule"):
        r, l = sh.size
        m = density * np.pi * r * r * l
        i = np.diag([m * (3 * r * r + l * l) / 12.0, m * (3 * r * r + l * l) / 12.0, 0.5 * m * r * r])
    elif getattr(sh, "hull", None) is not None:      # mesh: convex hull
        vol, c, i = sh.hull
        return density * vol, sh.pos + sh.rot @ c, density * (sh.rot @ i @ sh.rot.T)
    else:
        lo, hi = sh.size[:3], sh.size[3:]
        x, y, z = hi - lo
        m = density * x * y * z
        i = m / 12.0 * np.diag([y * y + z * z, x * x + z * z, x * x + y * y])
        c = sh.pos + sh.rot @ (0.5 * (lo + hi))
        return m, c, sh.rot @ i @ sh.rot.T
    return m, sh.pos.copy(), sh.rot @ i @ sh.rot.T


def _combine(m1, c1, i1, m2, c2, i2):
    """Merge two rigid bodies given (mass, com, inertia about own com) in one common frame."""
    m = m1 + m2
    if m <= 0.0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    c = (m1 * c1 + m2 * c2) / m

    def shift(mm, cc, ii):
        d = cc - c
        return ii + mm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))

    return m, c, shift(m1, c1, i1) + shift(m2, c2, i2)


def _prune_points(points):
    """Drop contact spheres that can never be the lowest point of their link against any plane
    (strictly inside the support hull of the others): keeps the candidate list short."""
    if len(points) <= 1:
        return points
    rng = np.random.default_rng(0)
    dirs = rng.normal(size=(4096, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs = np.concatenate([dirs, np.eye(3), -np.eye(3)])
    c = np.array([p[2] for p in points])
    r = np.array([p[3] for p in points])
    support = dirs @ c.T + r[None, :]
    best = support.max(axis=1, keepdims=True)
    keep = (support >= best - 1e-9).any(axis=0)
    return [p for p, k in zip(points, keep) if k]


def compile_urdf(path: str, options: Optional[AssetOptions] = None, prune_contacts: bool = True) -> Articulation:
    opts = options or AssetOptions()
    name, links, joints = parse_urdf(path)
    children: Dict[str, List[RawJoint]] = {}
    child_names = set()
    for j in joints:
        children.setdefault(j.parent, []).append(j)
        child_names.add(j.child)
    roots = [n for n in links if n not in child_names]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")
    root_name = roots[0]

    # links without an inertia tensor: derive one from their collision shapes (mass given -> rescale)
    for lk in links.values():
        if lk.has_inertia or not lk.shapes:
            continue
        acc = (0.0, np.zeros(3), np.zeros((3, 3)))
        for sh in lk.shapes:
            acc = _combine(*acc, *_shape_inertia(sh, opts.density))
        if acc[0] > 0:
            scale = (lk.mass / acc[0]) if lk.mass > 0 else 1.0
            if lk.mass <= 0:
                lk.mass = acc[0]
                lk.com = acc[1]
            d = acc[1] - lk.com
            i_about = acc[2] + acc[0] * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
            lk.inertia = i_about * scale
            lk.has_inertia = True

    body_names: List[str] = []
    body_link: List[int] = []
    body_T: List[tuple] = []
    link_names: List[str] = [root_name]
    link_parent: List[int] = [-1]
    link_inert = [(links[root_name].mass, links[root_name].com.copy(), links[root_name].inertia.copy())]
    dofs: List[RawJoint] = []
    dof_frames: List[tuple] = []
    cpts: List[tuple] = []          # (link, body, pos, radius)
    joint_dict: Dict[str, int] = {}
    jcount = [0]

    def add_shapes(lname, link_idx, body_idx, pos, rot):
        for sh in links[lname].shapes:
            for (c, r) in _shape_points(sh, opts):
                p = pos + rot @ (sh.pos + sh.rot @ c)
                cpts.append((link_idx, body_idx, p, r))

    def visit(lname, link_idx, pos, rot, body_idx):
        """``lname`` rides on dynamic link ``link_idx`` at (pos, rot) in that link's frame."""
        add_shapes(lname, link_idx, body_idx, pos, rot)
        for j in sorted(children.get(lname, []), key=lambda jj: jj.child):
            joint_dict[j.name] = jcount[0]
            jcount[0] += 1
            cpos = pos + rot @ j.pos
            crot = rot @ j.rot
            ch = links[j.child]
            if j.kind == "fixed":
                m, c, i = link_inert[link_idx]
                link_inert[link_idx] = _combine(m, c, i, ch.mass, cpos + crot @ ch.com, crot @ ch.inertia @ crot.T)
                if opts.collapse_fixed_joints:
                    b = body_idx
                else:
                    b = len(body_names)
                    body_names.append(j.child)
                    body_link.append(link_idx)
                    body_T.append((cpos, crot))
                visit(j.child, link_idx, cpos, crot, b)
            elif j.kind in ("revolute", "continuous", "prismatic"):
                new_link = len(link_names)
                link_names.append(j.child)
                link_parent.append(link_idx)
                link_inert.append((ch.mass, ch.com.copy(), ch.inertia.copy()))
                dofs.append(j)
                dof_frames.append((cpos, crot))
                b = len(body_names)
                body_names.append(j.child)
                body_link.append(new_link)
                body_T.append((np.zeros(3), np.eye(3)))
                visit(j.child, new_link, np.zeros(3), np.eye(3), b)
            else:
                raise ValueError(f"unsupported joint type {j.kind} ({j.name})")

    body_names.append(root_name)
    body_link.append(0)
    body_T.append((np.zeros(3), np.eye(3)))
    visit(root_name, 0, np.zeros(3), np.eye(3), 0)

    nl, nd = len(link_names), len(dofs)
    # star topology check + chains
    chain_start, chain_len = [], []
    d = 0
    while d < nd:
        if link_parent[d + 1] != 0:
            raise ValueError("only root + serial chains are supported")
        n = 1
        while d + n < nd and link_parent[d + n + 1] == d + n:
            n += 1
        chain_start.append(d)
        chain_len.append(n)
        d += n
    for li in range(1, nl):
        kids = [k for k in range(1, nl) if link_parent[k] == li]
        if len(kids) > 1 or (kids and kids[0] != li + 1):
            raise ValueError("only root + serial chains are supported (branching below the root)")

    if prune_contacts:
        pruned = []
        for li in range(nl):
            pts = [p for p in cpts if p[0] == li]
            pruned.extend(_prune_points(pts))
        cpts = pruned
    # order contact candidates: per chain, distal links first (feet win the per-chain cap); root last
    cpts.sort(key=lambda p: (p[0] == 0, _chain_of(p[0], chain_start, chain_len), -p[0]))

    axis = np.array([j.axis / np.linalg.norm(j.axis) for j in dofs]).reshape(nd, 3)
    art = Articulation(
        name=name,
        fixed_base=bool(opts.fix_base_link),
        body_names=body_names,
        body_link=np.array(body_link, dtype=np.int32),
        body_pos=np.array([t[0] for t in body_T]).reshape(-1, 3),
        body_quat=np.array([mat_to_quat(t[1]) for t in body_T]).reshape(-1, 4),
        dof_names=[j.name for j in dofs],
        link_names=link_names,
        link_parent=np.array(link_parent, dtype=np.int32),
        joint_type=np.array([JOINT_PRISMATIC if j.kind == "prismatic" else JOINT_REVOLUTE for j in dofs], dtype=np.int32),
        joint_pos=np.array([f[0] for f in dof_frames]).reshape(nd, 3),
        joint_quat=np.array([mat_to_quat(f[1]) for f in dof_frames]).reshape(nd, 4),
        joint_axis=axis,
        mass=np.array([t[0] for t in link_inert]),
        com=np.array([t[1] for t in link_inert]).reshape(nl, 3),
        inertia=np.array([t[2] for t in link_inert]).reshape(nl, 3, 3),
        lower=np.array([j.lower if j.has_limits else -np.inf for j in dofs]),
        upper=np.array([j.upper if j.has_limits else np.inf for j in dofs]),
        has_limits=np.array([j.has_limits for j in dofs], dtype=bool),
        effort=np.array([j.effort for j in dofs]),
        velocity=np.array([j.velocity for j in dofs]),
        damping=np.array([j.damping for j in dofs]),
        friction=np.array([j.friction for j in dofs]),
        armature=np.full(nd, float(opts.armature)),
        chain_start=np.array(chain_start, dtype=np.int32),
        chain_len=np.array(chain_len, dtype=np.int32),
        cp_link=np.array([p[0] for p in cpts], dtype=np.int32),
        cp_body=np.array([p[1] for p in cpts], dtype=np.int32),
        cp_pos=np.array([p[2] for p in cpts]).reshape(-1, 3),
        cp_radius=np.array([p[3] for p in cpts]),
        joint_dict=joint_dict,
    )
    return art


def _chain_of(link, chain_start, chain_len):
    if link == 0:
        return -1
    d = link - 1
    for c, (s, n) in enumerate(zip(chain_start, chain_len)):
        if s <= d < s + n:
            return c
    return -1


# ----------------------------------------------------------------------------------------------
# float64 forward kinematics (host side; known-answer tests + initial body states)
# ----------------------------------------------------------------------------------------------
def forward_kinematics(art: Articulation, q, root_pos=(0, 0, 0), root_quat=(0, 0, 0, 1)):
    """World pose (pos (nl,3), rot (nl,3,3)) of every dynamic link frame."""
    q = np.asarray(q, dtype=np.float64)
    pos = np.zeros((art.num_links, 3))
    rot = np.zeros((art.num_links, 3, 3))
    pos[0] = np.asarray(root_pos, dtype=np.float64)
    rot[0] = quat_to_mat(np.asarray(root_quat, dtype=np.float64))
    for d in range(art.num_dofs):
        p = art.link_parent[d + 1]
        rj = rot[p] @ quat_to_mat(art.joint_quat[d])
        pj = pos[p] + rot[p] @ art.joint_pos[d]
        if art.joint_type[d] == JOINT_REVOLUTE:
            rot[d + 1] = rj @ axis_angle_mat(art.joint_axis[d], q[d])
            pos[d + 1] = pj
        else:
            rot[d + 1] = rj
            pos[d + 1] = pj + rj @ (art.joint_axis[d] * q[d])
    return pos, rot


def body_poses(art: Articulation, q, root_pos=(0, 0, 0), root_quat=(0, 0, 0, 1)):
    """World pose of every API body frame."""
    lp, lr = forward_kinematics(art, q, root_pos, root_quat)
    pos = np.zeros((art.num_bodies, 3))
    rot = np.zeros((art.num_bodies, 3, 3))
    for b in range(art.num_bodies):
        l = art.body_link[b]
        pos[b] = lp[l] + lr[l] @ art.body_pos[b]
        rot[b] = lr[l] @ quat_to_mat(art.body_quat[b])
    return pos, rot
