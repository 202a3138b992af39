"""ctypes mirror of ``include/b200gym.h`` (the C ABI of ``libb200gym.so``).

Only PODs and function prototypes live here; the structures must match the header field for field
(``tests/test_abi.py`` checks ``sizeof`` against the values the library reports).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

B2G_ABI_VERSION = 2
MAX_DOF = 24
MAX_LINKS = MAX_DOF + 1
MAX_BODIES = 32
MAX_CHAINS = 8
MAX_CHAIN_LEN = 6
MAX_FIXED_CHAIN_LEN = 7
MAX_CPTS = 128
MAX_CONTACTS_PER_CHAIN = 8
LINK_SCALE_COLS = 6

DOF_MODE_NONE, DOF_MODE_POS, DOF_MODE_VEL, DOF_MODE_EFFORT = 0, 1, 2, 4

(T_ROOT_STATE, T_DOF_STATE, T_NET_CONTACT, T_DOF_FORCE, T_RIGID_BODY_STATE, T_DOF_TARGET, T_DOF_ACTUATION,
 T_JACOBIAN, T_MASS_MATRIX, T_FRICTION, T_ENV_SCALE, T_LINK_SCALE) = range(12)

(TT_OBS, TT_OBS_CLAMPED, TT_REW, TT_RESET, TT_PROGRESS, TT_TIMEOUT, TT_COMMANDS, TT_ACTIONS, TT_RAND_OVERRIDE, TT_TORQUES, TT_LAST_ACTIONS,
 TT_LAST_DOF_VEL, TT_FEET_AIR_TIME, TT_EPISODE_SUMS, TT_ENV_ORIGINS, TT_TERRAIN_LEVELS, TT_TERRAIN_TYPES, TT_NOISE_OVERRIDE, TT_PUSH_OVERRIDE,
 TT_EXTRAS, TT_MEASURED_HEIGHTS, TT_ARM_MM, TT_ARM_JAC, TT_EEF_STATE, TT_ARM_COMMANDS) = range(25)
REW_TERMS = 14

f32, i32 = C.c_float, C.c_int32


class Model(C.Structure):
    _fields_ = [
        ("fixed_base", i32), ("n_dof", i32), ("n_bodies", i32), ("n_chains", i32), ("n_cpts", i32),
        ("chain_start", i32 * MAX_CHAINS), ("chain_len", i32 * MAX_CHAINS),
        ("link_mass", f32 * MAX_LINKS), ("link_com", (f32 * 3) * MAX_LINKS), ("link_inertia", (f32 * 6) * MAX_LINKS),
        ("joint_type", i32 * MAX_DOF), ("joint_pos", (f32 * 3) * MAX_DOF), ("joint_quat", (f32 * 4) * MAX_DOF),
        ("joint_axis", (f32 * 3) * MAX_DOF),
        ("lower", f32 * MAX_DOF), ("upper", f32 * MAX_DOF), ("effort", f32 * MAX_DOF), ("vel_limit", f32 * MAX_DOF),
        ("armature", f32 * MAX_DOF),
        ("body_link", i32 * MAX_BODIES), ("body_pos", (f32 * 3) * MAX_BODIES), ("body_quat", (f32 * 4) * MAX_BODIES),
        ("cp_link", i32 * MAX_CPTS), ("cp_body", i32 * MAX_CPTS), ("cp_chain", i32 * MAX_CPTS),
        ("cp_pos", (f32 * 3) * MAX_CPTS), ("cp_radius", f32 * MAX_CPTS),
    ]


class SimParams(C.Structure):
    _fields_ = [
        ("dt", f32), ("substeps", i32), ("gravity", f32 * 3),
        ("num_position_iterations", i32), ("num_velocity_iterations", i32),
        ("contact_offset", f32), ("rest_offset", f32), ("bounce_threshold_velocity", f32),
        ("max_depenetration_velocity", f32),
        ("plane_static_friction", f32), ("plane_dynamic_friction", f32), ("plane_restitution", f32),
        ("has_ground", i32), ("joint_limit_stiffness", f32), ("joint_limit_damping", f32), ("max_contacts_per_chain", i32),
        ("max_linear_velocity", f32), ("max_angular_velocity", f32), ("self_collision", i32),
    ]


class DofProps(C.Structure):
    _fields_ = [
        ("drive_mode", i32 * MAX_DOF), ("stiffness", f32 * MAX_DOF), ("damping", f32 * MAX_DOF),
        ("effort", f32 * MAX_DOF), ("lower", f32 * MAX_DOF), ("upper", f32 * MAX_DOF), ("velocity", f32 * MAX_DOF),
    ]


class Heightfield(C.Structure):
    _fields_ = [
        ("rows", i32), ("cols", i32), ("horizontal_scale", f32), ("vertical_scale", f32),
        ("origin_x", f32), ("origin_y", f32), ("friction", f32), ("restitution", f32),
    ]


class TensorDesc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", i32), ("ndim", i32), ("shape", C.c_int64 * 4), ("device_id", i32)]


class AnymalCfg(C.Structure):
    _fields_ = [
        ("lin_vel_scale", f32), ("ang_vel_scale", f32), ("dof_pos_scale", f32), ("dof_vel_scale", f32),
        ("action_scale", f32),
        ("rew_lin_vel_xy", f32), ("rew_ang_vel_z", f32), ("rew_torque", f32),
        ("clip_obs", f32), ("clip_actions", f32),
        ("cmd_x", f32 * 2), ("cmd_y", f32 * 2), ("cmd_yaw", f32 * 2),
        ("default_dof_pos", f32 * MAX_DOF), ("init_root", f32 * 13),
        ("base_body", i32), ("n_knee", i32), ("knee_bodies", i32 * 8),
        ("max_episode_length", C.c_int64), ("seed", C.c_uint64),
    ]


class CartpoleCfg(C.Structure):
    _fields_ = [("reset_dist", f32), ("max_push_effort", f32), ("clip_obs", f32), ("clip_actions", f32),
                ("max_episode_length", C.c_int64), ("seed", C.c_uint64)]


class HoundarmCfg(C.Structure):
    _fields_ = [("clip_obs", f32), ("clip_actions", f32), ("action_scale", f32), ("dof_noise", f32), ("cmd_limit", f32 * 6),
                ("kp", f32), ("kp_null", f32), ("cmd_range", f32 * 6), ("dist_scale", f32), ("vel_scale", f32),
                ("eef_body", i32), ("jac_body", i32), ("max_episode_length", C.c_int64), ("seed", C.c_uint64),
                ("default_dof_pos", f32 * 8), ("n_reset_tail", i32), ("pad_", i32)]


class TerrainCfg(C.Structure):
    _fields_ = [
        ("lin_vel_scale", f32), ("ang_vel_scale", f32), ("dof_pos_scale", f32), ("dof_vel_scale", f32), ("height_meas_scale", f32),
        ("action_scale", f32), ("kp", f32), ("kd", f32), ("torque_limit", f32),
        ("decimation", i32), ("extra_sim_steps", i32), ("dt", f32), ("rew", f32 * REW_TERMS), ("base_height_target", f32),
        ("clip_obs", f32), ("clip_actions", f32), ("cmd_x", f32 * 2), ("cmd_y", f32 * 2), ("cmd_yaw", f32 * 2),
        ("default_dof_pos", f32 * MAX_DOF), ("init_root", f32 * 13),
        ("add_noise", i32), ("noise_lin_vel", f32), ("noise_ang_vel", f32), ("noise_gravity", f32), ("noise_dof_pos", f32),
        ("noise_dof_vel", f32), ("noise_height", f32),
        ("base_body", i32), ("n_knee", i32), ("knee_bodies", i32 * 8), ("n_feet", i32), ("feet_bodies", i32 * 8),
        ("hound_termination", i32), ("n_term_extra", i32), ("term_extra_bodies", i32 * 8), ("allow_knee_contacts", i32),
        ("hip_dofs", i32 * 4), ("max_episode_length", C.c_int64), ("push_interval", i32), ("max_episode_length_s", f32),
        ("custom_origins", i32), ("curriculum", i32), ("n_hx", i32), ("n_hy", i32), ("hx", f32 * 16), ("hy", f32 * 16),
        ("hs_rows", i32), ("hs_cols", i32), ("border_size", f32), ("hscale", f32), ("vscale", f32), ("env_length", f32),
        ("env_rows", i32), ("env_cols", i32), ("seed", C.c_uint64),
        ("n_ctrl_dof", i32), ("arm_chain", i32), ("arm_kp", f32), ("arm_kp_null", f32), ("arm_action_scale", f32), ("arm_dof_noise", f32),
        ("arm_cmd_limit", f32 * 6), ("eef_body", i32), ("jac_body", i32), ("refresh_eef", i32),
    ]


def _fill(dst, src):
    a = np.asarray(src)
    if a.ndim == 1:
        for i, v in enumerate(a):
            dst[i] = v.item() if hasattr(v, "item") else v
    else:
        for i, row in enumerate(a):
            _fill(dst[i], row)


def pack_model(art) -> Model:
    """``model.urdf.Articulation`` -> C ``b2g_model`` (float32 parameters)."""
    nd, nb, ncp, nc = art.num_dofs, art.num_bodies, len(art.cp_link), len(art.chain_start)
    if nd > MAX_DOF or nb > MAX_BODIES or ncp > MAX_CPTS or nc > MAX_CHAINS or (nc and max(art.chain_len) > (MAX_FIXED_CHAIN_LEN if (art.fixed_base and nc == 1) else MAX_CHAIN_LEN)):
        raise ValueError(f"articulation too large for the ABI limits (dof {nd}, bodies {nb}, contact points {ncp}, chains {nc})")
    m = Model()
    m.fixed_base = int(art.fixed_base)
    m.n_dof, m.n_bodies, m.n_chains, m.n_cpts = nd, nb, nc, ncp
    _fill(m.chain_start, art.chain_start)
    _fill(m.chain_len, art.chain_len)
    _fill(m.link_mass, art.mass.astype(np.float32))
    _fill(m.link_com, art.com.astype(np.float32))
    inert = np.stack([art.inertia[:, 0, 0], art.inertia[:, 1, 1], art.inertia[:, 2, 2],
                      art.inertia[:, 0, 1], art.inertia[:, 0, 2], art.inertia[:, 1, 2]], axis=1)
    _fill(m.link_inertia, inert.astype(np.float32))
    if nd:
        _fill(m.joint_type, art.joint_type)
        _fill(m.joint_pos, art.joint_pos.astype(np.float32))
        _fill(m.joint_quat, art.joint_quat.astype(np.float32))
        _fill(m.joint_axis, art.joint_axis.astype(np.float32))
        big = np.float32(3.0e38)
        _fill(m.lower, np.where(art.has_limits, art.lower, -big).astype(np.float32))
        _fill(m.upper, np.where(art.has_limits, art.upper, big).astype(np.float32))
        _fill(m.effort, art.effort.astype(np.float32))
        _fill(m.vel_limit, art.velocity.astype(np.float32))
        _fill(m.armature, art.armature.astype(np.float32))
    _fill(m.body_link, art.body_link)
    _fill(m.body_pos, art.body_pos.astype(np.float32))
    _fill(m.body_quat, art.body_quat.astype(np.float32))
    if ncp:
        owner = contact_owner_chains(art)
        # kernels need candidates grouped: per link (chains in order, distal link first), then root-link
        # candidates grouped by owner lane
        order = sorted(range(ncp), key=lambda i: (art.cp_link[i] == 0, owner[i], -int(art.cp_link[i]), i))
        order = np.array(order, dtype=np.int64)
        _fill(m.cp_link, art.cp_link[order])
        _fill(m.cp_body, art.cp_body[order])
        _fill(m.cp_chain, owner[order])
        _fill(m.cp_pos, art.cp_pos[order].astype(np.float32))
        _fill(m.cp_radius, art.cp_radius[order].astype(np.float32))
    return m


def contact_owner_chains(art) -> np.ndarray:
    """Chain (solver lane) that owns each contact candidate: the chain of its link; root-link candidates all go to
    lane 0 (the solver is Gauss-Seidel inside a lane and Jacobi across lanes, which is only safe for weakly coupled
    contacts, i.e. contacts on different chains)."""
    owner = np.zeros(len(art.cp_link), dtype=np.int32)
    for i, l in enumerate(art.cp_link):
        if l == 0:
            owner[i] = 0     # all root-link candidates belong to lane 0: contacts on one body must not be Jacobi-split
        else:
            d = l - 1
            for c, (s, n) in enumerate(zip(art.chain_start, art.chain_len)):
                if s <= d < s + n:
                    owner[i] = c
    return owner


def default_dof_props(art, drive_mode=DOF_MODE_NONE, stiffness=0.0, damping=0.0) -> DofProps:
    """What ``gym.get_asset_dof_properties`` would return, as the C struct."""
    p = DofProps()
    big = np.float32(3.0e38)
    for d in range(art.num_dofs):
        p.drive_mode[d] = drive_mode
        p.stiffness[d] = stiffness
        p.damping[d] = damping
        p.effort[d] = float(art.effort[d])
        p.lower[d] = float(art.lower[d]) if art.has_limits[d] else -big
        p.upper[d] = float(art.upper[d]) if art.has_limits[d] else big
        p.velocity[d] = float(art.velocity[d])
    return p
