// Host-side translation of the public PODs (include/b200gym.h) into the kernel-facing structures.
#pragma once

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "b2g_dev.h"

namespace b2g {

inline void quat_to_rot9(const float* q, float* m) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w); m[2] = 2 * (x * z + y * w);
    m[3] = 2 * (x * y + z * w); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - x * w);
    m[6] = 2 * (x * z - y * w); m[7] = 2 * (y * z + x * w); m[8] = 1 - 2 * (x * x + y * y);
}

constexpr int kSegLinks = 3;      // links per lane of the segment kernels

// Segment view of the chains (DevModel::seg_*): proximal pieces in lanes 0..n_chains-1, distal pieces behind them.
inline void build_segments(DevModel& d) {
    for (int c = 0; c < B2G_MAX_CHAINS; c++) { d.seg_start[c] = 0; d.seg_len[c] = 0; d.seg_par[c] = -1; d.seg_child[c] = -1; d.seg_store[c] = -1; }
    d.n_seg = 0; d.n_seg_store = 0;
    int n = d.n_chains;
    for (int c = 0; c < d.n_chains; c++) n += d.chain_len[c] > kSegLinks ? 1 : 0;
    if (n > B2G_MAX_CHAINS || d.n_chains == 0 || d.fixed_base) return;
    int next = d.n_chains;
    for (int c = 0; c < d.n_chains; c++) {
        const int len = d.chain_len[c];
        d.seg_start[c] = d.chain_start[c];
        d.seg_len[c] = len > kSegLinks ? kSegLinks : len;
        if (len > kSegLinks) {      // B2G_MAX_CHAIN_LEN = 2 kSegLinks: one distal piece at most
            d.seg_start[next] = d.chain_start[c] + kSegLinks;
            d.seg_len[next] = len - kSegLinks;
            d.seg_par[next] = c;
            d.seg_child[c] = next;
            d.seg_store[c] = d.n_seg_store++;
            next++;
        }
    }
    d.n_seg = next;
}

// Returns 0 on success, <0 when the model violates an assumption of the kernels (message in *why).
inline int pack_dev_model(const b2g_model& m, const b2g_dof_props& p, DevModel& d, const char** why) {
    memset(&d, 0, sizeof(d));
    *why = "";
    if (m.n_dof < 0 || m.n_dof > B2G_MAX_DOF || m.n_bodies < 1 || m.n_bodies > B2G_MAX_BODIES || m.n_chains < 0 ||
        m.n_chains > B2G_MAX_CHAINS || m.n_cpts < 0 || m.n_cpts > B2G_MAX_CPTS) {
        *why = "model dimensions out of range";
        return -1;
    }
    d.fixed_base = m.fixed_base; d.n_dof = m.n_dof; d.n_bodies = m.n_bodies; d.n_chains = m.n_chains;
    int expect = 0;
    for (int c = 0; c < m.n_chains; c++) {
        const int max_len = (m.fixed_base && m.n_chains == 1) ? B2G_MAX_FIXED_CHAIN_LEN : B2G_MAX_CHAIN_LEN;
        if (m.chain_start[c] != expect || m.chain_len[c] < 1 || m.chain_len[c] > max_len) {
            *why = "chains must be contiguous, non-empty and at most B2G_MAX_CHAIN_LEN (fixed-base single chain: B2G_MAX_FIXED_CHAIN_LEN) long";
            return -1;
        }
        d.chain_start[c] = m.chain_start[c]; d.chain_len[c] = m.chain_len[c];
        expect += m.chain_len[c];
    }
    if (expect != m.n_dof) { *why = "chain lengths do not add up to n_dof"; return -1; }
    build_segments(d);
    d.root_mass = m.link_mass[0];
    for (int i = 0; i < 3; i++) d.root_com[i] = m.link_com[0][i];
    // public order xx,yy,zz,xy,xz,yz -> kernel order xx,xy,xz,yy,yz,zz
    auto inertia6 = [](const float* s, float* o) { o[0] = s[0]; o[1] = s[3]; o[2] = s[4]; o[3] = s[1]; o[4] = s[5]; o[5] = s[2]; };
    inertia6(m.link_inertia[0], d.root_inertia);
    for (int k = 0; k < m.n_dof; k++) {
        DevDof& D = d.dof[k];
        for (int i = 0; i < 3; i++) { D.jpos[i] = m.joint_pos[k][i]; D.axis[i] = m.joint_axis[k][i]; D.com[i] = m.link_com[k + 1][i]; }
        D.type = m.joint_type[k];
        quat_to_rot9(m.joint_quat[k], D.jrot);
        D.mass = m.link_mass[k + 1];
        D.armature = m.armature[k];
        inertia6(m.link_inertia[k + 1], D.inertia);
        D.kp = p.stiffness[k]; D.kd = p.damping[k]; D.effort = p.effort[k]; D.vel_limit = p.velocity[k];
        D.lower = p.lower[k]; D.upper = p.upper[k]; D.drive_mode = p.drive_mode[k];
        D.cp_start = 0; D.cp_count = 0;
    }
    // contact candidates must be grouped: non-root candidates by link (any link order), then root candidates
    // grouped by owner chain.  Ranges are derived here.
    for (int c = 0; c < B2G_MAX_CHAINS; c++) { d.root_cp_start[c] = 0; d.root_cp_count[c] = 0; }
    for (int i = 0; i < m.n_cpts; i++) {
        for (int a = 0; a < 3; a++) d.cp[i][a] = m.cp_pos[i][a];
        d.cp[i][3] = m.cp_radius[i];
        d.cp_body[i] = m.cp_body[i];
        const int l = m.cp_link[i];
        if (l < 0 || l > m.n_dof || m.cp_body[i] < 0 || m.cp_body[i] >= m.n_bodies) { *why = "contact candidate refers to a bad link/body"; return -1; }
        if (l > 0) {
            DevDof& D = d.dof[l - 1];
            if (D.cp_count == 0) D.cp_start = i;
            else if (D.cp_start + D.cp_count != i) { *why = "contact candidates of a link must be contiguous"; return -1; }
            D.cp_count++;
        } else {
            const int c = m.cp_chain[i];
            if (c < 0 || c >= (m.n_chains > 0 ? m.n_chains : 1)) { *why = "root contact candidate has a bad owner chain"; return -1; }
            if (d.root_cp_count[c] == 0) d.root_cp_start[c] = i;
            else if (d.root_cp_start[c] + d.root_cp_count[c] != i) { *why = "root contact candidates of a chain must be contiguous"; return -1; }
            d.root_cp_count[c]++;
        }
    }
    // bounding boxes of each link's candidates (conservative early-out in the kernels)
    for (int l = 0; l <= m.n_dof; l++) {
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        int cnt = 0;
        for (int i = 0; i < m.n_cpts; i++) {
            if (m.cp_link[i] != l) continue;
            cnt++;
            for (int a = 0; a < 3; a++) {
                lo[a] = fminf(lo[a], m.cp_pos[i][a] - m.cp_radius[i]);
                hi[a] = fmaxf(hi[a], m.cp_pos[i][a] + m.cp_radius[i]);
            }
        }
        float* c = l == 0 ? d.root_cp_c : d.dof[l - 1].cp_c;
        float* h = l == 0 ? d.root_cp_h : d.dof[l - 1].cp_h;
        for (int a = 0; a < 3; a++) { c[a] = cnt ? 0.5f * (lo[a] + hi[a]) : 0.0f; h[a] = cnt ? 0.5f * (hi[a] - lo[a]) : 0.0f; }
    }
    {   // self-collision box (b2g_sim_params::self_collision): the base body's own candidates and their radii
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        int cnt = 0;
        for (int i = 0; i < m.n_cpts; i++) {
            if (m.cp_link[i] != 0 || m.cp_body[i] != 0) continue;
            cnt++;
            for (int a = 0; a < 3; a++) {
                lo[a] = fminf(lo[a], m.cp_pos[i][a] - m.cp_radius[i]);
                hi[a] = fmaxf(hi[a], m.cp_pos[i][a] + m.cp_radius[i]);
            }
        }
        for (int a = 0; a < 3; a++) { d.self_box_c[a] = cnt ? 0.5f * (lo[a] + hi[a]) : 0.0f; d.self_box_h[a] = cnt ? 0.5f * (hi[a] - lo[a]) : -1e30f; }
    }
    for (int b = 0; b < m.n_bodies; b++) {
        d.body_link[b] = m.body_link[b];
        for (int a = 0; a < 3; a++) d.body_pos[b][a] = m.body_pos[b][a];
        for (int a = 0; a < 4; a++) d.body_quat[b][a] = m.body_quat[b][a];
    }
    return 0;
}

// Largest distance between a link's candidate bounding-box centre and any of its candidate centres (box half diagonal).
inline float max_link_radius(const DevModel& d) {
    float r2 = d.root_cp_h[0] * d.root_cp_h[0] + d.root_cp_h[1] * d.root_cp_h[1] + d.root_cp_h[2] * d.root_cp_h[2];
    for (int l = 0; l < d.n_dof; l++) {
        const float* h = d.dof[l].cp_h;
        r2 = fmaxf(r2, h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
    }
    return sqrtf(r2);
}

// Conservative coarse bound of a heightfield for the per-link contact early-out (b2g_dynamics.cuh::may_touch).
// Block (I, J) covers the samples [I*B, I*B+B) x [J*B, J*B+B) dilated by ceil(rmax / hs) + 2 samples on every side, so the
// bound found from a link's bounding-box centre holds for every candidate of that link.  out[2*(I*ccols+J)] = max height
// (+ 1e-4 m), out[..+1] = min over the dilated block's triangles of the normal's z (x 0.9999): the margins absorb the
// kernels' approximate division / rsqrt, so skipping a link never changes a result.
inline void build_hf_coarse(const int16_t* smp, int rows, int cols, float hs, float vs, float rmax, std::vector<float>& out, int& crows, int& ccols) {
    const int B = B2G_HFC_BLOCK;
    const int dil = (int)ceilf(rmax / hs) + 2;
    crows = (rows + B - 1) / B; ccols = (cols + B - 1) / B;
    // per fine cell: min normal z of its two triangles (same expressions as ground_sample)
    std::vector<float> nz((size_t)(rows - 1) * (cols - 1));
    for (int i = 0; i < rows - 1; i++) {
        for (int j = 0; j < cols - 1; j++) {
            const int16_t* s = smp + (size_t)i * cols + j;
            const float h00 = vs * (float)s[0], h01 = vs * (float)s[1], h10 = vs * (float)s[cols], h11 = vs * (float)s[cols + 1];
            const float ax = (h10 - h00) / hs, ay = (h11 - h10) / hs, bx = (h11 - h01) / hs, by = (h01 - h00) / hs;
            const float na = 1.0f / sqrtf(ax * ax + ay * ay + 1.0f), nb = 1.0f / sqrtf(bx * bx + by * by + 1.0f);
            nz[(size_t)i * (cols - 1) + j] = fminf(na, nb);
        }
    }
    // separable window max / min: rows first, then columns
    std::vector<float> hrow((size_t)crows * cols), nrow((size_t)crows * (cols - 1));
    for (int I = 0; I < crows; I++) {
        const int lo = I * B - dil < 0 ? 0 : I * B - dil;
        const int hi = I * B + B - 1 + dil + 1 > rows - 1 ? rows - 1 : I * B + B - 1 + dil + 1;   // samples lo..hi, cells lo..hi-1
        for (int j = 0; j < cols; j++) {
            float m = -1e30f;
            for (int i = lo; i <= hi; i++) m = fmaxf(m, vs * (float)smp[(size_t)i * cols + j]);
            hrow[(size_t)I * cols + j] = m;
        }
        for (int j = 0; j < cols - 1; j++) {
            float m = 1.0f;
            for (int i = lo; i < hi; i++) m = fminf(m, nz[(size_t)i * (cols - 1) + j]);
            nrow[(size_t)I * (cols - 1) + j] = m;
        }
    }
    out.assign((size_t)crows * ccols * 2, 0.0f);
    for (int I = 0; I < crows; I++) {
        for (int J = 0; J < ccols; J++) {
            const int lo = J * B - dil < 0 ? 0 : J * B - dil;
            const int hi = J * B + B - 1 + dil + 1 > cols - 1 ? cols - 1 : J * B + B - 1 + dil + 1;
            float mh = -1e30f, mn = 1.0f;
            for (int j = lo; j <= hi; j++) mh = fmaxf(mh, hrow[(size_t)I * cols + j]);
            for (int j = lo; j < hi; j++) mn = fminf(mn, nrow[(size_t)I * (cols - 1) + j]);
            out[2 * ((size_t)I * ccols + J)] = mh + 1e-4f;
            out[2 * ((size_t)I * ccols + J) + 1] = mn * 0.9999f;
        }
    }
}

inline int contact_slots(const b2g_sim_params& s) {
    const int n = s.max_contacts_per_chain > 0 ? s.max_contacts_per_chain : B2G_DEFAULT_CONTACTS_PER_CHAIN;
    return n > B2G_MAX_CONTACTS_PER_CHAIN ? B2G_MAX_CONTACTS_PER_CHAIN : n;
}

inline void pack_dev_params(const b2g_sim_params& s, const b2g_heightfield* hf, const int16_t* hf_dev, DevParams& d, const float* hfc_dev = nullptr,
                            int hfc_rows = 0, int hfc_cols = 0) {
    memset(&d, 0, sizeof(d));
    const int sub = s.substeps > 0 ? s.substeps : 1;
    d.h = s.dt / (float)sub;
    d.substeps = sub;
    for (int i = 0; i < 3; i++) d.g[i] = s.gravity[i];
    d.npos = s.num_position_iterations; d.nvel = s.num_velocity_iterations;
    d.contact_offset = s.contact_offset; d.max_depen = s.max_depenetration_velocity;
    d.mu_ground = (hf && hf_dev) ? hf->friction : s.plane_dynamic_friction;
    d.has_ground = s.has_ground;
    d.limit_kp = s.joint_limit_stiffness; d.limit_kd = s.joint_limit_damping;
    d.max_contacts = contact_slots(s);
    d.self_collide = s.self_collision != 0;
    d.max_lin_vel = s.max_linear_velocity > 0.0f ? s.max_linear_velocity : 0.0f;
    d.max_ang_vel = s.max_angular_velocity > 0.0f ? s.max_angular_velocity : 0.0f;
    {
        const char* ba = getenv("B2G_BLOCK_ALIGN");
        d.block_align = ba ? atoi(ba) : 1;      // default: the block's warps re-align at the start of every sub-step (B2G_BLOCK_ALIGN=0: never)
    }
    if (hf && hf_dev) {
        d.hf = hf_dev; d.hf_rows = hf->rows; d.hf_cols = hf->cols; d.hf_hs = hf->horizontal_scale; d.hf_vs = hf->vertical_scale;
        d.hf_ox = hf->origin_x; d.hf_oy = hf->origin_y;
        d.hfc = hfc_dev; d.hfc_rows = hfc_rows; d.hfc_cols = hfc_cols;
    }
}

}  // namespace b2g
