// Device-side (kernel-facing) model / parameter structures, built by the host API from the public
// b2g_model / b2g_dof_props / b2g_sim_params PODs of include/b200gym.h.
#pragma once

#include <stdint.h>

#include "b200gym.h"

namespace b2g {

struct DevDof {
    float jpos[3];      // joint frame origin in the parent link frame
    int type;           // b2g_joint_type
    float jrot[9];      // joint frame rotation in the parent frame (row-major)
    float axis[3];      // unit axis in the child frame
    float mass;
    float com[3];
    float armature;
    float inertia[6];   // xx, xy, xz, yy, yz, zz about the com, link axes
    int cp_start, cp_count;   // contact candidates riding on this link
    float cp_c[3], cp_h[3];   // bounding box (centre, half extents incl. radii) of those candidates, link frame
    float kp, kd, effort, vel_limit;
    float lower, upper;
    int drive_mode;
    int pad;
};

struct DevModel {
    int fixed_base, n_dof, n_bodies, n_chains;
    int chain_start[B2G_MAX_CHAINS];
    int chain_len[B2G_MAX_CHAINS];
    int root_cp_start[B2G_MAX_CHAINS];   // root-link candidates owned by each lane
    int root_cp_count[B2G_MAX_CHAINS];
    // Segment view of the chains (the <8,3> kernels): every chain is cut into pieces of at most three links that own a lane each.
    // Lane c < n_chains holds the proximal piece of chain c (so root candidates and task code that name a chain keep their lane);
    // the distal pieces of chains longer than three links follow in chain order.  n_seg = 0: the model does not fit (more than
    // B2G_MAX_CHAINS pieces) and the kernels that walk whole chains are used instead.
    int n_seg;
    int seg_start[B2G_MAX_CHAINS], seg_len[B2G_MAX_CHAINS];
    int seg_par[B2G_MAX_CHAINS];      // lane of the piece this one hangs off (its last link), -1 = the root
    int seg_child[B2G_MAX_CHAINS];    // lane of the piece hanging off this one, -1 = none
    int seg_store[B2G_MAX_CHAINS];    // index of this piece's slot in the shared-memory ancestor store (pieces with a child), -1 = none
    int n_seg_store;
    float root_mass;
    float root_com[3];
    float root_inertia[6];
    float root_cp_c[3], root_cp_h[3];    // bounding box of the root-link contact candidates
    float self_box_c[3], self_box_h[3];  // self-collision box: bounding box of the candidates of API body 0 (the base itself, without bodies fixed to it)
    DevDof dof[B2G_MAX_DOF];
    float cp[B2G_MAX_CPTS][4];   // link-frame position, radius
    int cp_body[B2G_MAX_CPTS];
    int body_link[B2G_MAX_BODIES];
    float body_pos[B2G_MAX_BODIES][3];
    float body_quat[B2G_MAX_BODIES][4];
};

struct DevParams {
    float h;             // sub-step length
    int substeps;
    float g[3];
    int npos, nvel;
    float contact_offset;
    float max_depen;
    float mu_ground;
    int has_ground;
    float limit_kp, limit_kd;   // joint-limit spring / damper
    int block_align;            // bit mask of the points of a sub-step at which the block's warps re-align (b2g_dynamics.cuh::block_align)
    float max_lin_vel, max_ang_vel;   // root velocity clamps (0 = none), b2g_sim_params::max_linear_velocity / max_angular_velocity
    int self_collide;           // b2g_sim_params::self_collision: link candidates also collide with the root's bounding box
    int max_contacts;           // contact slots per lane in use (<= B2G_MAX_CONTACTS_PER_CHAIN); sizes the shared-memory scratch
    // heightfield (null -> plane z = 0)
    const int16_t* hf;
    int hf_rows, hf_cols;
    float hf_hs, hf_vs, hf_ox, hf_oy;
    // coarse conservative bound of the heightfield (b2g_host_pack.h::build_hf_coarse; null -> no early-out): per block of
    // B2G_HFC_BLOCK x B2G_HFC_BLOCK samples, dilated by the largest link radius, {max height, min normal z}
    const float* hfc;
    int hfc_rows, hfc_cols;
    // contact statistics (b2g_sim_contact_stats; null -> not collected): [0] active contact points, [1] candidates inside the contact
    // offset that found no free slot (B2G_MAX_CONTACTS_PER_CHAIN per lane) and were DROPPED, [2] environment sub-steps in which at
    // least one candidate was dropped, [3] environment sub-steps simulated
    unsigned long long* stats;
};

constexpr int B2G_HFC_SHIFT = 3;
constexpr int B2G_HFC_BLOCK = 1 << B2G_HFC_SHIFT;

}  // namespace b2g
