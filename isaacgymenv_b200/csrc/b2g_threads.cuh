// Per-thread bodies of the kernels (one thread = one lane = one chain of one environment).
// Kept in a header so the very same code is compiled into the CUDA kernels (b200gym.cu) and into the
// host lane emulator used by the CPU tests (tests/emu/emu.cpp, B2G_HOST_EMU).
#pragma once

#include "b2g_dynamics.cuh"

namespace b2g {

struct SimArgs {
    const DevModel* M;
    DevParams P;
    int n_envs;
    float* root;             // (N,13)
    float* dof;              // (N,nd,2)
    const float* target;     // (N,nd)
    const float* actuation;  // (N,nd)
    float* dof_force;        // (N,nd)
    float* contact;          // (N,nb,3)
    const float* friction;   // (N) per-env shape friction
    const float* env_scale = nullptr;   // (N,4) per-env scale of link masses, drive stiffness, drive damping, (spare); null = ones
    const float* link_scale = nullptr;  // (N,nd+1,B2G_LINK_SCALE_COLS) per-link randomisation rows (B2G_T_LINK_SCALE); null = none
};

// the randomisation scales of environment `env` as the sub-step reads them
B2G_HD B2G_INL EnvDr env_dr(const SimArgs& A, int env, bool live, bool with_friction = true) {
    return load_env_dr(with_friction ? A.friction : nullptr, A.env_scale, env, live, A.link_scale, A.M->n_dof + 1);
}

struct TaskArgs {
    b2g_anymal_cfg cfg;
    b2g_cartpole_cfg ccfg;
    b2g_houndarm_cfg hcfg;
    unsigned long long seed;
    const float* actions_in;   // (N,na)
    float* obs;                // (N,12+3*nd)
    float* obs_clamped;
    float* rew;                // (N)
    long long* reset;          // (N) int64
    long long* progress;       // (N)
    long long* timeout;        // (N)
    float* commands;           // (N,3)
    float* actions;            // (N,na)
    int* reset_count;          // (N)
    const float* rand_override; // (N, 2*nd+3) or null
    int post_only;              // 1: skip the physics, run post_physics_step on the tensors as they are (parity tests)
};

// DOF range of a lane: the whole chain, or (segment variant) the lane's piece of at most three links
template <int LANES, int NL>
B2G_HD B2G_INL void lane_span(const DevModel* M, int lane, int& len, int& d0) {
    if (is_segmented<LANES, NL>()) {
        len = lane < M->n_seg ? M->seg_len[lane] : 0;
        d0 = lane < M->n_seg ? M->seg_start[lane] : 0;
    } else {
        len = lane < M->n_chains ? M->chain_len[lane] : 0;
        d0 = lane < M->n_chains ? M->chain_start[lane] : 0;
    }
}

template <int NL>
B2G_HD B2G_INL void load_state(const SimArgs& A, int env, int len, int d0, LaneState<NL>& st) {
    const float* r = A.root + (size_t)env * 13;
    st.rp = V3{r[0], r[1], r[2]};
    st.qx = r[3]; st.qy = r[4]; st.qz = r[5]; st.qw = r[6];
    st.rv = V3{r[7], r[8], r[9]};
    st.rw = V3{r[10], r[11], r[12]};
    const int nd = A.M->n_dof;
#pragma unroll
    for (int j = 0; j < NL; j++) {
        st.q[j] = 0; st.qd[j] = 0; st.tgt[j] = 0; st.act[j] = 0; st.frc[j] = 0;
        if (j < len) {
            const size_t k = (size_t)env * nd + d0 + j;
            st.q[j] = A.dof[2 * k];
            st.qd[j] = A.dof[2 * k + 1];
        }
    }
}

template <int NL>
B2G_HD B2G_INL void store_state(const SimArgs& A, int env, int lane, int len, int d0, const LaneState<NL>& st, bool fixed) {
    const int nd = A.M->n_dof;
    if (lane == 0 && !fixed) {
        float* r = A.root + (size_t)env * 13;
        r[0] = st.rp.x; r[1] = st.rp.y; r[2] = st.rp.z;
        r[3] = st.qx; r[4] = st.qy; r[5] = st.qz; r[6] = st.qw;
        r[7] = st.rv.x; r[8] = st.rv.y; r[9] = st.rv.z;
        r[10] = st.rw.x; r[11] = st.rw.y; r[12] = st.rw.z;
    }
#pragma unroll
    for (int j = 0; j < NL; j++) {
        if (j < len) {
            const size_t k = (size_t)env * nd + d0 + j;
            A.dof[2 * k] = st.q[j];
            A.dof[2 * k + 1] = st.qd[j];
            A.dof_force[k] = st.frc[j];
        }
    }
}

// gym.simulate: `substeps` sub-steps for one lane.
template <int LANES, int NL, bool FIXED, bool HF>
B2G_HD B2G_INL void simulate_thread(const SimArgs& A, int env, int lane, bool valid, ScratchStrided sc, float* bf) {
    const DevModel* M = A.M;
    int len, d0;
    lane_span<LANES, NL>(M, lane, len, d0);
    LaneState<NL> st;
    load_state<NL>(A, env, len, d0, st);
    const int nd = M->n_dof;
#pragma unroll
    for (int j = 0; j < NL; j++) {
        if (j < len) {
            const size_t k = (size_t)env * nd + d0 + j;
            st.tgt[j] = A.target[k];
            st.act[j] = A.actuation[k];
        }
    }
    const EnvDr mu_shape = env_dr(A, valid ? env : 0, valid);
    for (int s = 0; s < A.P.substeps; s++)
        substep<LANES, NL, FIXED, HF, false, (LANES == 4 && NL == 3) || (LANES == 1 && NL == 2)>(M, A.P, lane, len, d0, st, mu_shape, s == A.P.substeps - 1, sc, bf);
    if (valid) {
        store_state<NL>(A, env, lane, len, d0, st, FIXED);
        const int nb3 = M->n_bodies * 3;
        for (int i = lane; i < nb3; i += LANES) A.contact[(size_t)env * nb3 + i] = bf[i];
    }
}

// Philox4x32-10 (Salmon et al., SC'11); checked against the Random123 known-answer vectors in the tests.
B2G_HD B2G_INL void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)c0 * 0xD2511F53ull;
        const unsigned long long p1 = (unsigned long long)c2 * 0xCD9E8D57ull;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
        const unsigned n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
        const unsigned n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// i-th uniform in [0,1) of env `env` at its `rc`-th reset
B2G_HD B2G_INL float reset_uniform(const TaskArgs& T, int env, int rc, int i, int n_draws) {
    if (T.rand_override) return T.rand_override[(size_t)env * n_draws + i];
    unsigned o[4];
    philox4x32((unsigned)env, (unsigned)rc, (unsigned)(i >> 2), 0u, (unsigned)(T.seed & 0xffffffffull), (unsigned)(T.seed >> 32), o);
    return (float)(o[i & 3] >> 8) * (1.0f / 16777216.0f);
}

// utils/torch_jit_utils.py:93-103 (same arithmetic form as the reference)
B2G_HD B2G_INL V3 quat_rotate_inverse(float qx, float qy, float qz, float qw, V3 v) {
    const V3 qv = V3{qx, qy, qz};
    const V3 a = v * (2.0f * qw * qw - 1.0f);
    const V3 b = cross(qv, v) * qw * 2.0f;
    const V3 c = qv * dot(qv, v) * 2.0f;
    return a - b + c;
}
// utils/torch_jit_utils.py:80-90
B2G_HD B2G_INL V3 quat_rotate(float qx, float qy, float qz, float qw, V3 v) {
    const V3 qv = V3{qx, qy, qz};
    const V3 a = v * (2.0f * qw * qw - 1.0f);
    const V3 b = cross(qv, v) * qw * 2.0f;
    const V3 c = qv * dot(qv, v) * 2.0f;
    return a + b + c;
}

// (upper - lower) * u + lower exactly as torch evaluates it (utils/torch_jit_utils.py:215-218): a rounded multiply
// followed by a rounded add -- no FMA contraction, so reset draws are bit-identical to the reference formula
B2G_HD B2G_INL float rand_range(float lower, float upper, float u) {
#if defined(B2G_HOST_EMU)
    volatile float p = (upper - lower) * u;
    return p + lower;
#else
    return __fadd_rn(__fmul_rn(upper - lower, u), lower);
#endif
}

// reset_idx of the flat task for one lane (tasks/anymal.py:278-304): state <- init root, dof draws, commands
template <int NL>
B2G_HD B2G_INL void anymal_reset_lane(const TaskArgs& T, int env, int lane, int len, int d0, int nd, int rc, LaneState<NL>& st, float* cmd) {
    const int n_draws = 2 * nd + 3;
    const float* ir = T.cfg.init_root;
    st.rp = V3{ir[0], ir[1], ir[2]};
    st.qx = ir[3]; st.qy = ir[4]; st.qz = ir[5]; st.qw = ir[6];
    st.rv = V3{ir[7], ir[8], ir[9]};
    st.rw = V3{ir[10], ir[11], ir[12]};
#pragma unroll
    for (int j = 0; j < NL; j++) {
        if (j < len) {
            const int d = d0 + j;
            const float up = reset_uniform(T, env, rc, d, n_draws), uv = reset_uniform(T, env, rc, nd + d, n_draws);
            st.q[j] = T.cfg.default_dof_pos[d] * rand_range(0.5f, 1.5f, up);
            st.qd[j] = rand_range(-0.1f, 0.1f, uv);
        }
    }
    cmd[0] = rand_range(T.cfg.cmd_x[0], T.cfg.cmd_x[1], reset_uniform(T, env, rc, 2 * nd + 0, n_draws));
    cmd[1] = rand_range(T.cfg.cmd_y[0], T.cfg.cmd_y[1], reset_uniform(T, env, rc, 2 * nd + 1, n_draws));
    cmd[2] = rand_range(T.cfg.cmd_yaw[0], T.cfg.cmd_yaw[1], reset_uniform(T, env, rc, 2 * nd + 2, n_draws));
}

// The reset_idx(all) the task constructor performs (tasks/anymal.py:146).
template <int LANES, int NL>
B2G_HD B2G_INL void anymal_reset_all_thread(const SimArgs& A, const TaskArgs& T, int env, int lane, bool valid) {
    const DevModel* M = A.M;
    int len, d0;
    lane_span<LANES, NL>(M, lane, len, d0);
    LaneState<NL> st;
    load_state<NL>(A, env, len, d0, st);
    float cmd[3];
    const int rc = T.reset_count[env];
    anymal_reset_lane<NL>(T, env, lane, len, d0, M->n_dof, rc, st, cmd);
    Grp<LANES>::sync();
    if (valid) {
        store_state<NL>(A, env, lane, len, d0, st, false);
        if (lane == 0) {
            T.commands[(size_t)env * 3 + 0] = cmd[0]; T.commands[(size_t)env * 3 + 1] = cmd[1]; T.commands[(size_t)env * 3 + 2] = cmd[2];
            T.progress[env] = 0;
            T.reset[env] = 1;
            T.reset_count[env] = rc + 1;
        }
    }
}

// VecTask.step for the flat Anymal / Hound task, fused:
//   clamp actions, position targets   (vec_task.py:374, tasks/anymal.py:226-229)
//   gym.simulate                       (vec_task.py:379-382)
//   post_physics_step                  (tasks/anymal.py:231-239: progress, reset_idx, observations, reward)
//   timeout_buf, clamp obs             (vec_task.py:394,402)
template <int LANES, int NL, bool HF>
B2G_HD B2G_INL void anymal_step_thread(const SimArgs& A, const TaskArgs& T, int env, int lane, bool valid, ScratchStrided sc, float* bf) {
    const DevModel* M = A.M;
    const int nd = M->n_dof;
    int len, d0;
    lane_span<LANES, NL>(M, lane, len, d0);
    const b2g_anymal_cfg& C = T.cfg;
    LaneState<NL> st;
    load_state<NL>(A, env, len, d0, st);
    // post-physics inputs are fetched now so their (possibly DRAM) latency hides behind the physics
    const long long progress_in = T.progress[env];
    const long long reset_in = T.reset[env];
    const float cmd_in[3] = {T.commands[(size_t)env * 3], T.commands[(size_t)env * 3 + 1], T.commands[(size_t)env * 3 + 2]};
    const int rc_in = T.reset_count[env];
    float act[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) {
        act[j] = 0.0f;
        if (j < len) {
            const int d = d0 + j;
            float a = T.actions_in[(size_t)env * nd + d];
            a = fminf(fmaxf(a, -C.clip_actions), C.clip_actions);
            act[j] = a;
            st.tgt[j] = C.action_scale * a + C.default_dof_pos[d];
        }
    }
    if (!T.post_only) {
        const EnvDr mu_shape = env_dr(A, valid ? env : 0, valid);
#pragma unroll 1
        for (int s = 0; s < A.P.substeps; s++)
            substep<LANES, NL, false, HF, false, (LANES == 4 && NL == 3), !(LANES == 4 && NL == 3)>(M, A.P, lane, len, d0, st, mu_shape, s == A.P.substeps - 1, sc, bf);      // quadruped variant: no self-collision code (b2g_task_anymal_create refuses the flag)
    } else {
        // torques and contact forces come from the sim tensors instead of a physics step
#pragma unroll
        for (int j = 0; j < NL; j++)
            if (j < len) st.frc[j] = A.dof_force[(size_t)env * nd + d0 + j];
        const int nb3p = M->n_bodies * 3;
        for (int i = lane; i < nb3p; i += LANES) bf[i] = A.contact[(size_t)env * nb3p + i];
        Grp<LANES>::sync();
    }

    // ---- post_physics_step ----
    long long progress = progress_in + 1;
    float cmd[3] = {cmd_in[0], cmd_in[1], cmd_in[2]};
    const bool do_reset = reset_in != 0;
    int rc = 0;
    if (do_reset) {
        rc = rc_in;
        anymal_reset_lane<NL>(T, env, lane, len, d0, nd, rc, st, cmd);
        progress = 0;
    }
    // observations (tasks/anymal.py:354-386)
    const V3 lin = quat_rotate_inverse(st.qx, st.qy, st.qz, st.qw, st.rv);
    const V3 ang = quat_rotate_inverse(st.qx, st.qy, st.qz, st.qw, st.rw);
    const V3 pg = quat_rotate(st.qx, st.qy, st.qz, st.qw, V3{0.0f, 0.0f, -1.0f});
    // reward (tasks/anymal.py:311-351)
    float tq2 = 0.0f;
#pragma unroll
    for (int j = 0; j < NL; j++)
        if (j < len) tq2 += st.frc[j] * st.frc[j];
    tq2 = Grp<LANES>::sum(tq2);
    const float ex = cmd[0] - lin.x, ey = cmd[1] - lin.y, ez = cmd[2] - ang.z;
    const float lin_err = ex * ex + ey * ey, ang_err = ez * ez;
    float rew = expf(-lin_err / 0.25f) * C.rew_lin_vel_xy + expf(-ang_err / 0.25f) * C.rew_ang_vel_z + tq2 * C.rew_torque;
    rew = fmaxf(rew, 0.0f);
    // termination: base or knee contact force above 1 N, or time-out
    bool term = false;
    {
        const float* f = bf + C.base_body * 3;
        term = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 1.0f;
        for (int k = 0; k < C.n_knee; k++) {
            const float* g = bf + C.knee_bodies[k] * 3;
            term = term || (sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) > 1.0f);
        }
    }
    const bool time_out = progress >= C.max_episode_length - 1;
    const bool reset = term || time_out;
    if (valid) {
        store_state<NL>(A, env, lane, len, d0, st, false);
        const int nb3 = M->n_bodies * 3;
        for (int i = lane; i < nb3; i += LANES) A.contact[(size_t)env * nb3 + i] = bf[i];
        const int no = 12 + 3 * nd;
        float* o = T.obs + (size_t)env * no;
        float* oc = T.obs_clamped + (size_t)env * no;
        const float clip = C.clip_obs;
#define B2G_PUT(idx, val) { const float vv = (val); o[idx] = vv; oc[idx] = fminf(fmaxf(vv, -clip), clip); }
        if (lane == 0) {
            B2G_PUT(0, lin.x * C.lin_vel_scale) B2G_PUT(1, lin.y * C.lin_vel_scale) B2G_PUT(2, lin.z * C.lin_vel_scale)
            B2G_PUT(3, ang.x * C.ang_vel_scale) B2G_PUT(4, ang.y * C.ang_vel_scale) B2G_PUT(5, ang.z * C.ang_vel_scale)
            B2G_PUT(6, pg.x) B2G_PUT(7, pg.y) B2G_PUT(8, pg.z)
            B2G_PUT(9, cmd[0] * C.lin_vel_scale) B2G_PUT(10, cmd[1] * C.lin_vel_scale) B2G_PUT(11, cmd[2] * C.ang_vel_scale)
            T.rew[env] = rew;
            T.reset[env] = reset ? 1 : 0;
            T.progress[env] = progress;
            T.timeout[env] = (time_out && reset) ? 1 : 0;
            T.commands[(size_t)env * 3 + 0] = cmd[0]; T.commands[(size_t)env * 3 + 1] = cmd[1]; T.commands[(size_t)env * 3 + 2] = cmd[2];
            if (do_reset) T.reset_count[env] = rc + 1;
        }
#pragma unroll
        for (int j = 0; j < NL; j++) {
            if (j < len) {
                const int d = d0 + j;
                B2G_PUT(12 + d, (st.q[j] - C.default_dof_pos[d]) * C.dof_pos_scale)
                B2G_PUT(12 + nd + d, st.qd[j] * C.dof_vel_scale)
                B2G_PUT(12 + 2 * nd + d, act[j])
                T.actions[(size_t)env * nd + d] = act[j];
            }
        }
#undef B2G_PUT
    }
}

// VecTask.step for Cartpole, fused (tasks/cartpole.py:159-163 effort, :166-175 post_physics_step, :144-158 reset_idx,
// :131-142 observations, :180-196 reward/reset; vec_task.py:394,402 tail).  One thread per environment.
B2G_HD B2G_INL void cartpole_step_thread(const SimArgs& A, const TaskArgs& T, int env, bool valid, ScratchStrided sc, float* bf) {
    const DevModel* M = A.M;
    const b2g_cartpole_cfg& C = T.ccfg;
    LaneState<2> st;
    load_state<2>(A, env, 2, 0, st);
    float a = T.actions_in[env];
    a = fminf(fmaxf(a, -C.clip_actions), C.clip_actions);
    st.act[0] = a * C.max_push_effort;
    st.act[1] = 0.0f;
    if (!T.post_only) {
        for (int s = 0; s < A.P.substeps; s++)
            substep<1, 2, true, false, false, true>(M, A.P, 0, 2, 0, st, env_dr(A, valid ? env : 0, valid, false), s == A.P.substeps - 1, sc, bf);
    }
    long long progress = T.progress[env] + 1;
    long long reset_prev = T.reset[env];
    int rc = 0;
    const bool do_reset = reset_prev != 0;
    if (do_reset) {
        rc = T.reset_count[env];
        // positions = 0.2 * (rand - 0.5), velocities = 0.5 * (rand - 0.5); draw order: pos[0], pos[1], vel[0], vel[1]
        st.q[0] = 0.2f * (reset_uniform(T, env, rc, 0, 4) - 0.5f);
        st.q[1] = 0.2f * (reset_uniform(T, env, rc, 1, 4) - 0.5f);
        st.qd[0] = 0.5f * (reset_uniform(T, env, rc, 2, 4) - 0.5f);
        st.qd[1] = 0.5f * (reset_uniform(T, env, rc, 3, 4) - 0.5f);
        reset_prev = 0;
        progress = 0;
    }
    const float cart_pos = st.q[0], cart_vel = st.qd[0], pole_angle = st.q[1], pole_vel = st.qd[1];
    float rew = 1.0f - pole_angle * pole_angle - 0.01f * fabsf(cart_vel) - 0.005f * fabsf(pole_vel);
    const bool out_x = fabsf(cart_pos) > C.reset_dist, out_a = fabsf(pole_angle) > 1.57079632679489661923f;
    if (out_x) rew = -2.0f;
    if (out_a) rew = -2.0f;
    const bool time_out = progress >= C.max_episode_length - 1;
    const long long reset = (out_x || out_a || time_out) ? 1 : reset_prev;
    if (valid) {
        store_state<2>(A, env, 0, 2, 0, st, true);
        float* o = T.obs + (size_t)env * 4;
        float* oc = T.obs_clamped + (size_t)env * 4;
        const float v[4] = {cart_pos, cart_vel, pole_angle, pole_vel};
        for (int k = 0; k < 4; k++) { o[k] = v[k]; oc[k] = fminf(fmaxf(v[k], -C.clip_obs), C.clip_obs); }
        T.rew[env] = rew;
        T.reset[env] = reset;
        T.progress[env] = progress;
        T.timeout[env] = (time_out && reset != 0) ? 1 : 0;
        T.actions[env] = a;
        if (do_reset) T.reset_count[env] = rc + 1;
    }
}


// ================================================================================================
// Rough-terrain locomotion task (AnymalTerrain / HoundTerrain), two kernels per step
// ================================================================================================
struct TerrainArgs {
    b2g_terrain_cfg cfg;
    const float* actions_in;     // (N,na)
    float* obs;                  // (N,no)
    float* obs_clamped;
    float* rew;
    long long* reset;            // (N) 0/1
    long long* progress;
    long long* timeout;          // (N) previous step's value is an input of the reward (anymal_terrain.py:367)
    float* commands;             // (N,4) x, y, yaw rate (heading-derived), heading
    float* actions;              // (N,na)
    float* torques;              // (N,nd)
    float* last_actions;
    float* last_dof_vel;
    float* feet_air_time;        // (N,4)
    float* episode_sums;         // (13,N)
    float* env_origins;          // (N,3)
    long long* terrain_levels;   // (N)
    long long* terrain_types;    // (N)
    const float* terrain_origins;   // (rows,cols,3) or null
    const short* height_samples;    // (hs_rows,hs_cols) or null
    float* scratch;              // (N,9) base lin vel, base ang vel, projected gravity of this step (pre-reset)
    float* resetw;               // (N) reset ? |cmd_xy|^2 : 0      (curriculum scalar, quirk Q10)
    float* report;               // (13,N) episode sums of the envs that reset this step (else 0)
    float* measured;             // (N,n_height_points)
    // cross-block reductions of the step (device kernels only; the host emulator computes them in its driver)
    float* cnorm = nullptr;            // curriculum scalar sqrt(sum resetw), written by the last block of k_terrain_phys
    float* extras_part = nullptr;      // (blocks,16) per-block sums for extras["episode"]
    float* extras = nullptr;           // (15) means over the envs that reset, mean terrain level, reset count
    unsigned* tickets = nullptr;       // [0]: k_terrain_phys, [1]: k_terrain_post block-arrival counters
    long long* step_ctr_advance = nullptr;   // counter the last block of the step increments (device-side step counter)
    float* arm_mm;               // (N,36) arm block of the mass matrix as of the last post_physics refresh (pre-reset)
    float* arm_jac;              // (N,36) Jacobian slice the task takes: base columns of body `jac_body`
    float* eef_state;            // (N,13) end-effector rigid-body row (never refreshed unless cfg.refresh_eef)
    float* arm_commands;         // (N,3)
    int* reset_count;
    const float* reset_override; // (N,29)
    const float* noise_override; // (N,no)
    const float* push_override;  // (N,2)
    long long common_step;       // counter AFTER this step's increment
    const long long* step_ctr = nullptr;   // when set, the counter is read from device memory instead (CUDA-graph replay)
    int init_done;
    int post_only;
    unsigned long long seed;
};

B2G_HD B2G_INL float terrain_uniform(const TerrainArgs& T, const float* ovr, int ncol, int env, unsigned c1, unsigned stream, int i) {
    if (ovr) return ovr[(size_t)env * ncol + i];
    unsigned o[4];
    philox4x32((unsigned)env, c1, (unsigned)(i >> 2), stream, (unsigned)(T.seed & 0xffffffffull), (unsigned)(T.seed >> 32), o);
    return (float)(o[i & 3] >> 8) * (1.0f / 16777216.0f);
}

// utils/torch_jit_utils.py:70-77
B2G_HD B2G_INL V3 quat_apply_v(float qx, float qy, float qz, float qw, V3 b) {
    const V3 xyz = V3{qx, qy, qz};
    const V3 t = cross(xyz, b) * 2.0f;
    return b + t * qw + cross(xyz, t);
}
// tasks/anymal_terrain.py:683-687 as TorchScript runs it (C fmod)
B2G_HD B2G_INL float wrap_to_pi_f(float a) {
    a = fmodf(a, 6.283185307179586f);
    if (a > 3.141592653589793f) a -= 6.283185307179586f;
    return a;
}
B2G_HD B2G_INL float norm3(const float* f) { return sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]); }


// ---- hound + arm helpers (tasks/useful_hound.py) ----
// general 6x6 inverse, Gauss-Jordan with partial pivoting (what torch.inverse does for the OSC law, :668-670).
// The OSC law inverts M and J M^-1 J^T; on the arms of this repo (a 19 g last link) the second matrix reaches condition numbers of
// 1e8-1e10, where a float32 evaluation -- the reference's included -- is dominated by round-off.  The handful of 6x6 operations
// per policy step are therefore carried out in float64 (a few thousand DP flops per environment), inputs and outputs float32.
typedef double osc_real;
// n x n inverse, row-major with leading dimension n (n <= 7)
B2G_HD inline void inv_n(const osc_real* a, osc_real* inv, int n) {
    osc_real m[7][14];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { m[i][j] = a[i * n + j]; m[i][n + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < n; c++) {
        int piv = c;
        osc_real best = fabs(m[c][c]);
        for (int r = c + 1; r < n; r++) if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); piv = r; }
        if (piv != c) for (int j = 0; j < 2 * n; j++) { const osc_real t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        const osc_real d = 1.0 / m[c][c];
        for (int j = 0; j < 2 * n; j++) m[c][j] *= d;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const osc_real f = m[r][c];
            for (int j = 0; j < 2 * n; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = m[i][n + j];
}
// o (r x c) = op(a) op(b), inner dimension k; a is stored (r x k) or, with ta, (k x r); b (k x c) or, with tb, (c x k); o may alias a or b
B2G_HD inline void mat_mul(const osc_real* a, const osc_real* b, osc_real* o, int r, int k, int c, bool ta = false, bool tb = false) {
    osc_real t[49];
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) {
            osc_real acc = 0.0;
            for (int q = 0; q < k; q++) acc += (ta ? a[q * r + i] : a[i * k + q]) * (tb ? b[j * k + q] : b[q * c + j]);
            t[i * c + j] = acc;
        }
    for (int i = 0; i < r * c; i++) o[i] = t[i];
}
// _compute_osc_torques (tasks/useful_hound.py:660-691, hound_arm.py / manipulator.py:534-560), split in two because the mass-matrix
// block, the Jacobian slice, the commanded pose change and the end-effector velocity row do not change inside one policy step:
//   prepare (once per step):  u_task = J^T Lambda (kp dpose - kd eef_vel),  N = (I - J^T Lambda J M^-1) M
//   apply   (every decimation step):  u = clamp(u_task + N (kd_null (-qd) + kp_null wrap(q_default - q)), +-effort)
// NJ joints (6: the hound's arm; 7: the Franka), six task dimensions; mm is NJ x NJ, J is 6 x NJ, both dense row-major.
template <int NJ>
struct OscPreparedN {
    float u_task[NJ];
    float N[NJ * NJ];
};
typedef OscPreparedN<6> OscPrepared;
template <int NJ>
B2G_HD inline void osc_prepare_n(const float* mm_f, const float* j_f, const float* dpose, const float* eef_vel, float kp, OscPreparedN<NJ>& P) {
    const osc_real kd = 2.0 * sqrt((osc_real)kp);
    osc_real mm[NJ * NJ], j[6 * NJ], mm_inv[NJ * NJ], t[49], m_eef_inv[36], m_eef[36], j_eef_inv[6 * NJ];
    for (int i = 0; i < NJ * NJ; i++) mm[i] = mm_f[i];
    for (int i = 0; i < 6 * NJ; i++) j[i] = j_f[i];
    inv_n(mm, mm_inv, NJ);
    mat_mul(j, mm_inv, t, 6, NJ, NJ);                       // J M^-1
    mat_mul(t, j, m_eef_inv, 6, NJ, 6, false, true);        // J M^-1 J^T
    inv_n(m_eef_inv, m_eef, 6);
    osc_real w[6], v[6];
    for (int i = 0; i < 6; i++) w[i] = (osc_real)kp * dpose[i] - kd * eef_vel[i];
    for (int i = 0; i < 6; i++) { osc_real acc = 0; for (int k = 0; k < 6; k++) acc += m_eef[i * 6 + k] * w[k]; v[i] = acc; }
    for (int i = 0; i < NJ; i++) { osc_real acc = 0; for (int k = 0; k < 6; k++) acc += j[k * NJ + i] * v[k]; P.u_task[i] = (float)acc; }   // J^T (Lambda w)
    mat_mul(m_eef, j, t, 6, 6, NJ);
    mat_mul(t, mm_inv, j_eef_inv, 6, NJ, NJ);
    mat_mul(j, j_eef_inv, t, NJ, 6, NJ, true, false);       // J^T j_eef_inv
    for (int i = 0; i < NJ * NJ; i++) t[i] = ((i / NJ == i % NJ) ? 1.0 : 0.0) - t[i];
    mat_mul(t, mm, t, NJ, NJ, NJ);
    for (int i = 0; i < NJ * NJ; i++) P.N[i] = (float)t[i];
}
template <int NJ>
B2G_HD inline void osc_apply_n(const OscPreparedN<NJ>& P, const float* q, const float* qd, float kp_null, const float* effort, float* u, const float* q_default = nullptr) {
    const float kdn = 2.0f * sqrtf(kp_null);
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float un[NJ];
    for (int i = 0; i < NJ; i++) {
        float a = (q_default ? q_default[i] : 0.0f) - q[i] + pi;
        a = a - two_pi * floorf(a / two_pi);            // python-style remainder (eager torch %)
        un[i] = kdn * -qd[i] + kp_null * (a - pi);
    }
    for (int i = 0; i < NJ; i++) {
        float acc = P.u_task[i];
        for (int k = 0; k < NJ; k++) acc += P.N[i * NJ + k] * un[k];
        u[i] = fminf(fmaxf(acc, -effort[i]), effort[i]);
    }
}
B2G_HD inline void osc_prepare(const float* mm_f, const float* j_f, const float* dpose, const float* eef_vel, float kp, OscPrepared& P) {
    osc_prepare_n<6>(mm_f, j_f, dpose, eef_vel, kp, P);
}
B2G_HD inline void osc_apply(const OscPrepared& P, const float* q, const float* qd, float kp_null, const float* effort, float* u) {
    osc_apply_n<6>(P, q, qd, kp_null, effort, u);
}
// Kinematics + composite-rigid-body pass over one chain: joint-space mass-matrix block of the chain (n x n, row-major in
// mm[ld * ld], leading dimension ld >= n), the position (relative to the root origin) of link `want_link` (chain-local index, -1 = root) and its world rotation.
B2G_HD inline void chain_crba(const DevModel* M, int d0, int n, const float* rootq, const float* q, float* mm, int want_link, V3* want_pos, M3* want_rot,
                              const float* qd, V3 root_w, V3 root_v, SV* want_vel, float mass_scale = 1.0f, const float* link_scale = nullptr, int ld = 6) {
    M3 R = quat_to_m3(rootq[0], rootq[1], rootq[2], rootq[3]);
    V3 p = V3{0, 0, 0};
    SV vel = SV{root_w, root_v};
    SV S[B2G_MAX_FIXED_CHAIN_LEN];
    SI I[B2G_MAX_FIXED_CHAIN_LEN];
    if (want_link < 0) { *want_pos = p; *want_rot = R; *want_vel = vel; }
    for (int j = 0; j < n; j++) {
        const DevDof& D = M->dof[d0 + j];
        M3 jr;
        for (int k = 0; k < 9; k++) jr.m[k] = D.jrot[k];
        const V3 ax = V3{D.axis[0], D.axis[1], D.axis[2]};
        const M3 RJ = mul(R, jr);
        const V3 pj = p + mul(R, V3{D.jpos[0], D.jpos[1], D.jpos[2]});
        const V3 axw = mul(RJ, ax);
        if (D.type == B2G_JOINT_REVOLUTE) { R = mul(RJ, axis_angle_m3(ax, q[j])); p = pj; S[j] = SV{axw, cross(pj, axw)}; }
        else { R = RJ; p = pj + axw * q[j]; S[j] = SV{V3{0, 0, 0}, axw}; }
        vel = vel + S[j] * qd[j];
        const V3 cw = p + mul(R, V3{D.com[0], D.com[1], D.com[2]});
        const float ms = link_scale ? mass_scale * link_scale[B2G_LINK_SCALE_COLS * (1 + d0 + j)] : mass_scale;
        I[j] = rigid_inertia(D.mass * ms, cw, rotate_sym(R, S3{D.inertia[0] * ms, D.inertia[1] * ms, D.inertia[2] * ms, D.inertia[3] * ms, D.inertia[4] * ms,
                                                            D.inertia[5] * ms}));
        if (j == want_link) { *want_pos = p; *want_rot = R; *want_vel = vel; }
    }
    for (int j = n - 2; j >= 0; j--) I[j] += I[j + 1];
    for (int i = 0; i < ld * ld; i++) mm[i] = 0.0f;
    for (int i = 0; i < n; i++) {
        const SV F = mul(I[i], S[i]);
        mm[i * ld + i] = dot(S[i], F) + M->dof[d0 + i].armature;
        for (int k = 0; k < i; k++) { const float v = dot(S[k], F); mm[i * ld + k] = v; mm[k * ld + i] = v; }
    }
}
// what post_physics_step's refresh_jacobian / refresh_mass_matrix leave for the next step's OSC (pre-reset state), plus the
// optional live end-effector row
// q / qd = the arm's joint state (whole chain), st = the caller's lane state (root part used)
template <int NL>
B2G_HD inline void arm_refresh(const SimArgs& A, const TerrainArgs& T, int env, int d0, int len, const LaneState<NL>& st, const float* q, const float* qd) {
    const DevModel* M = A.M;
    const b2g_terrain_cfg& C = T.cfg;
    const float rq[4] = {st.qx, st.qy, st.qz, st.qw};
    float mm[36];
    // Jacobian slice jacobian[:, jac_body, :, :6] of a floating-base actor = the base's six columns (linear rows first):
    // [[I, -[r]x], [0, I]] with r = position of that body relative to the base origin
    const int jl = M->body_link[C.jac_body];
    const int jloc = (jl == 0) ? -1 : jl - 1 - d0;
    V3 lp; M3 lr; SV lv;
    chain_crba(M, d0, len, rq, q, mm, jloc, &lp, &lr, qd, st.rw, st.rv, &lv, A.env_scale ? A.env_scale[(size_t)env * 4] : 1.0f,
               A.link_scale ? A.link_scale + (size_t)env * (M->n_dof + 1) * B2G_LINK_SCALE_COLS : nullptr);
    const V3 r = lp + mul(lr, V3{M->body_pos[C.jac_body][0], M->body_pos[C.jac_body][1], M->body_pos[C.jac_body][2]});
    float* J = T.arm_jac + (size_t)env * 36;
    for (int i = 0; i < 36; i++) J[i] = 0.0f;
    for (int i = 0; i < 6; i++) J[i * 6 + i] = 1.0f;
    J[0 * 6 + 4] = r.z; J[0 * 6 + 5] = -r.y; J[1 * 6 + 3] = -r.z; J[1 * 6 + 5] = r.x; J[2 * 6 + 3] = r.y; J[2 * 6 + 4] = -r.x;
    float* MM = T.arm_mm + (size_t)env * 36;
    for (int i = 0; i < 36; i++) MM[i] = mm[i];
    if (C.refresh_eef) {
        const int el = M->body_link[C.eef_body];
        const int eloc = (el == 0) ? -1 : el - 1 - d0;
        float dummy[36];
        V3 ep; M3 er; SV ev;
        chain_crba(M, d0, len, rq, q, dummy, eloc, &ep, &er, qd, st.rw, st.rv, &ev);
        const V3 off = mul(er, V3{M->body_pos[C.eef_body][0], M->body_pos[C.eef_body][1], M->body_pos[C.eef_body][2]});
        const V3 pos = ep + off;
        const V3 lin = ev.v + cross(ev.w, pos);
        float* e = T.eef_state + (size_t)env * 13;
        e[0] = st.rp.x + pos.x; e[1] = st.rp.y + pos.y; e[2] = st.rp.z + pos.z;
        // orientation: body rotation as a quaternion (xyzw)
        const M3 Rb = mul(er, quat_to_m3(M->body_quat[C.eef_body][0], M->body_quat[C.eef_body][1], M->body_quat[C.eef_body][2], M->body_quat[C.eef_body][3]));
        const float tr = Rb.m[0] + Rb.m[4] + Rb.m[8];
        float qx, qy, qz, qw;
        if (tr > 0.0f) { float s2 = sqrtf(tr + 1.0f) * 2.0f; qw = 0.25f * s2; qx = (Rb.m[7] - Rb.m[5]) / s2; qy = (Rb.m[2] - Rb.m[6]) / s2; qz = (Rb.m[3] - Rb.m[1]) / s2; }
        else if (Rb.m[0] > Rb.m[4] && Rb.m[0] > Rb.m[8]) { float s2 = sqrtf(1.0f + Rb.m[0] - Rb.m[4] - Rb.m[8]) * 2.0f; qw = (Rb.m[7] - Rb.m[5]) / s2; qx = 0.25f * s2; qy = (Rb.m[1] + Rb.m[3]) / s2; qz = (Rb.m[2] + Rb.m[6]) / s2; }
        else if (Rb.m[4] > Rb.m[8]) { float s2 = sqrtf(1.0f + Rb.m[4] - Rb.m[0] - Rb.m[8]) * 2.0f; qw = (Rb.m[2] - Rb.m[6]) / s2; qx = (Rb.m[1] + Rb.m[3]) / s2; qy = 0.25f * s2; qz = (Rb.m[5] + Rb.m[7]) / s2; }
        else { float s2 = sqrtf(1.0f + Rb.m[8] - Rb.m[0] - Rb.m[4]) * 2.0f; qw = (Rb.m[3] - Rb.m[1]) / s2; qx = (Rb.m[2] + Rb.m[6]) / s2; qy = (Rb.m[5] + Rb.m[7]) / s2; qz = 0.25f * s2; }
        e[3] = qx; e[4] = qy; e[5] = qz; e[6] = qw;
        e[7] = lin.x; e[8] = lin.y; e[9] = lin.z; e[10] = ev.w.x; e[11] = ev.w.y; e[12] = ev.w.z;
    }
}

// ---- Houndarm (tasks/hound_arm.py): the whole VecTask.step of the fixed-base arm reach task, one thread per environment ----
// kinematics of the single chain: link poses relative to the root origin (world axes), link spatial velocities, joint axes
template <int NJ>
B2G_HD inline void arm_chain_kin(const DevModel* M, const LaneState<NJ>& st, int n, M3* Rl, V3* pl, SV* vl, V3* axw, V3* pj) {
    Rl[0] = quat_to_m3(st.qx, st.qy, st.qz, st.qw);
    pl[0] = V3{0, 0, 0};
    vl[0] = sv0();
    for (int j = 0; j < n; j++) {
        const DevDof& D = M->dof[j];
        M3 jr;
        for (int k = 0; k < 9; k++) jr.m[k] = D.jrot[k];
        const V3 ax = V3{D.axis[0], D.axis[1], D.axis[2]};
        const M3 RJ = mul(Rl[j], jr);
        pj[j] = pl[j] + mul(Rl[j], V3{D.jpos[0], D.jpos[1], D.jpos[2]});
        axw[j] = mul(RJ, ax);
        SV S;
        if (D.type == B2G_JOINT_REVOLUTE) { Rl[j + 1] = mul(RJ, axis_angle_m3(ax, st.q[j])); pl[j + 1] = pj[j]; S = SV{axw[j], cross(pj[j], axw[j])}; }
        else { Rl[j + 1] = RJ; pl[j + 1] = pj[j] + axw[j] * st.q[j]; S = SV{V3{0, 0, 0}, axw[j]}; }
        vl[j + 1] = vl[j] + S * st.qd[j];
    }
}
// rigid-body state row of API body b (same conventions as body_state_env): pos3, quat xyzw, linear velocity, angular velocity
template <int NJ>
B2G_HD inline void arm_body_row(const DevModel* M, int b, const LaneState<NJ>& st, const M3* Rl, const V3* pl, const SV* vl, float* o) {
    const int l = M->body_link[b];
    const V3 p = pl[l] + mul(Rl[l], V3{M->body_pos[b][0], M->body_pos[b][1], M->body_pos[b][2]});
    const M3 R = mul(Rl[l], quat_to_m3(M->body_quat[b][0], M->body_quat[b][1], M->body_quat[b][2], M->body_quat[b][3]));
    float qx, qy, qz, qw;
    const float tr = R.m[0] + R.m[4] + R.m[8];
    if (tr > 0.0f) { float s = sqrtf(tr + 1.0f) * 2.0f; qw = 0.25f * s; qx = (R.m[7] - R.m[5]) / s; qy = (R.m[2] - R.m[6]) / s; qz = (R.m[3] - R.m[1]) / s; }
    else if (R.m[0] > R.m[4] && R.m[0] > R.m[8]) { float s = sqrtf(1.0f + R.m[0] - R.m[4] - R.m[8]) * 2.0f; qw = (R.m[7] - R.m[5]) / s; qx = 0.25f * s; qy = (R.m[1] + R.m[3]) / s; qz = (R.m[2] + R.m[6]) / s; }
    else if (R.m[4] > R.m[8]) { float s = sqrtf(1.0f + R.m[4] - R.m[0] - R.m[8]) * 2.0f; qw = (R.m[2] - R.m[6]) / s; qx = (R.m[1] + R.m[3]) / s; qy = 0.25f * s; qz = (R.m[5] + R.m[7]) / s; }
    else { float s = sqrtf(1.0f + R.m[8] - R.m[0] - R.m[4]) * 2.0f; qw = (R.m[3] - R.m[1]) / s; qx = (R.m[2] + R.m[6]) / s; qy = (R.m[5] + R.m[7]) / s; qz = 0.25f * s; }
    const V3 v = vl[l].v + cross(vl[l].w, p);
    o[0] = st.rp.x + p.x; o[1] = st.rp.y + p.y; o[2] = st.rp.z + p.z;
    o[3] = qx; o[4] = qy; o[5] = qz; o[6] = qw;
    o[7] = v.x; o[8] = v.y; o[9] = v.z;
    o[10] = vl[l].w.x; o[11] = vl[l].w.y; o[12] = vl[l].w.z;
}

// NJ = 6: Houndarm, 7: Manipulator (the same task on the Franka; b2g_houndarm_cfg::default_dof_pos / n_reset_tail).  Six actions either way.
template <int NJ>
B2G_HD B2G_INL void houndarm_step_thread(const SimArgs& A, const TaskArgs& T, int env, bool valid, ScratchStrided sc, float* bf) {
    const DevModel* M = A.M;
    const b2g_houndarm_cfg& C = T.hcfg;
    const int n = M->chain_len[0];
    constexpr int NA = 6;      // actions: the commanded end-effector pose change
    LaneState<NJ> st;
    load_state<NJ>(A, env, n, 0, st);
    float act[NA];
    for (int j = 0; j < NA; j++) {
        const float a = T.actions_in[(size_t)env * NA + j];
        act[j] = fminf(fmaxf(a, -C.clip_actions), C.clip_actions);
    }
    for (int j = 0; j < NJ; j++) {
        st.act[j] = 0.0f;
        st.tgt[j] = 0.0f;
        st.frc[j] = (j < n) ? A.dof_force[(size_t)env * n + j] : 0.0f;
    }
    M3 Rl[NJ + 1];
    V3 pl[NJ + 1], axw[NJ], pj[NJ];
    SV vl[NJ + 1];
    if (!T.post_only) {
        // pre_physics_step (:495-507): OSC torques from the state as it is now
        arm_chain_kin<NJ>(M, st, n, Rl, pl, vl, axw, pj);
        float eef[13], J[6 * NJ], mm[NJ * NJ], dpose[6], effort[NJ], u[NJ];
        arm_body_row<NJ>(M, C.eef_body, st, Rl, pl, vl, eef);
        const int l = M->body_link[C.jac_body];
        const V3 pb = pl[l] + mul(Rl[l], V3{M->body_pos[C.jac_body][0], M->body_pos[C.jac_body][1], M->body_pos[C.jac_body][2]});
        for (int i = 0; i < 6 * NJ; i++) J[i] = 0.0f;
        for (int d = 0; d <= l - 1; d++) {
            if (M->dof[d].type == B2G_JOINT_REVOLUTE) {
                const V3 lin = cross(axw[d], pb - pj[d]);
                J[0 * NJ + d] = lin.x; J[1 * NJ + d] = lin.y; J[2 * NJ + d] = lin.z;
                J[3 * NJ + d] = axw[d].x; J[4 * NJ + d] = axw[d].y; J[5 * NJ + d] = axw[d].z;
            } else {
                J[0 * NJ + d] = axw[d].x; J[1 * NJ + d] = axw[d].y; J[2 * NJ + d] = axw[d].z;
            }
        }
        const float rq[4] = {st.qx, st.qy, st.qz, st.qw};
        V3 dp; M3 dr; SV dv;
        chain_crba(M, 0, n, rq, st.q, mm, -1, &dp, &dr, st.qd, V3{0, 0, 0}, V3{0, 0, 0}, &dv, A.env_scale ? A.env_scale[(size_t)(valid ? env : 0) * 4] : 1.0f,
                   A.link_scale ? A.link_scale + (size_t)(valid ? env : 0) * (M->n_dof + 1) * B2G_LINK_SCALE_COLS : nullptr, NJ);
        for (int i = n; i < NJ; i++) mm[i * NJ + i] = 1.0f;      // a shorter chain: keep the padded block invertible
        for (int i = 0; i < 6; i++) dpose[i] = act[i] * C.cmd_limit[i] / C.action_scale;
        for (int i = 0; i < NJ; i++) effort[i] = (i < n) ? M->dof[i].effort : 0.0f;
        OscPreparedN<NJ> P;
        osc_prepare_n<NJ>(mm, J, dpose, eef + 7, C.kp, P);
        osc_apply_n<NJ>(P, st.q, st.qd, C.kp_null, effort, u, C.default_dof_pos);
        for (int j = 0; j < NJ; j++) st.act[j] = (j < n) ? u[j] : 0.0f;
        const EnvDr dr_env = env_dr(A, valid ? env : 0, valid);
#pragma unroll 1
        for (int s = 0; s < A.P.substeps; s++)
            substep<1, NJ, true, false, false, false>(M, A.P, 0, n, 0, st, dr_env, s == A.P.substeps - 1, sc, bf);
    }
    // post_physics_step (:509-517)
    long long progress = T.progress[env] + 1;
    long long reset_prev = T.reset[env];
    int rc = 0;
    float cmd[3] = {T.commands[(size_t)env * 3 + 0], T.commands[(size_t)env * 3 + 1], T.commands[(size_t)env * 3 + 2]};
    const bool do_reset = reset_prev != 0;
    if (do_reset) {          // reset_idx (:394-459); draw order: command x, y, z, then one joint-noise uniform per DOF
        rc = T.reset_count[env];
        const int nd = 3 + n;
        for (int k = 0; k < 3; k++) cmd[k] = rand_range(C.cmd_range[2 * k], C.cmd_range[2 * k + 1], reset_uniform(T, env, rc, k, nd));
        for (int j = 0; j < n; j++) {
            const float r = reset_uniform(T, env, rc, 3 + j, nd);
            const float p = C.default_dof_pos[j] + C.dof_noise * 2.0f * (r - 0.5f);
            st.q[j] = fminf(fmaxf(p, M->dof[j].lower), M->dof[j].upper);
            if (j >= n - C.n_reset_tail) st.q[j] = C.default_dof_pos[j];      // manipulator.py:417 (after the clamp, no noise)
            st.qd[j] = 0.0f;
        }
        progress = 0;
        reset_prev = 0;
    }
    // compute_observations (:383-392) on the refreshed state, compute_reward (:550-567)
    arm_chain_kin<NJ>(M, st, n, Rl, pl, vl, axw, pj);
    float eef[13];
    arm_body_row<NJ>(M, C.eef_body, st, Rl, pl, vl, eef);
    const float dx = eef[0] - cmd[0], dy = eef[1] - cmd[1], dz = eef[2] - cmd[2];
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    float vsq = 0.0f;
    for (int k = 7; k < 13; k++) vsq += eef[k] * eef[k];
    float rew = (1.0f - tanhf(10.0f * dist)) * C.dist_scale + (1.0f - tanhf(10.0f * sqrtf(vsq))) * (dist < 0.02f ? 1.0f : 0.0f) * C.vel_scale;
    rew = fmaxf(rew, 0.0f);
    const bool time_out = progress >= C.max_episode_length - 1;
    const long long reset = time_out ? 1 : reset_prev;
    if (valid) {
        store_state<NJ>(A, env, 0, n, 0, st, true);
        float* o = T.obs + (size_t)env * 10;
        float* oc = T.obs_clamped + (size_t)env * 10;
        for (int k = 0; k < 7; k++) o[k] = eef[k];
        for (int k = 0; k < 3; k++) o[7 + k] = cmd[k];
        for (int k = 0; k < 10; k++) oc[k] = fminf(fmaxf(o[k], -C.clip_obs), C.clip_obs);
        T.rew[env] = rew;
        T.reset[env] = reset;
        T.progress[env] = progress;
        T.timeout[env] = (time_out && reset != 0) ? 1 : 0;
        for (int j = 0; j < NA; j++) T.actions[(size_t)env * NA + j] = act[j];
        for (int k = 0; k < 3; k++) T.commands[(size_t)env * 3 + k] = cmd[k];
        if (do_reset) T.reset_count[env] = rc + 1;
    }
}

// kernel 1: pre_physics_step (decimation loop) + extra sim step + post_physics_step up to and including the reward
// ARM = false compiles the operational-space arm law out (quadruped-only robots: no float64 code, no OSC stack frame)
template <int LANES, int NL, bool HF, bool ARM = (NL >= 6 || is_segmented<LANES, NL>())>
B2G_HD B2G_INL void terrain_phys_thread(const SimArgs& A, const TerrainArgs& T, int env, int lane, bool valid, ScratchStrided sc, float* bf) {
    const DevModel* M = A.M;
    const b2g_terrain_cfg& C = T.cfg;
    const int nd = M->n_dof;
    int len, d0;
    lane_span<LANES, NL>(M, lane, len, d0);
    LaneState<NL> st;
    load_state<NL>(A, env, len, d0, st);
    const long long progress = T.progress[env] + 1;
    const bool timeout_prev = T.timeout[env] != 0;
    float cmd[4] = {T.commands[(size_t)env * 4], T.commands[(size_t)env * 4 + 1], T.commands[(size_t)env * 4 + 2], T.commands[(size_t)env * 4 + 3]};
    float act[NL], tq[NL], lact[NL], lqd[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) {
        act[j] = 0; tq[j] = 0; lact[j] = 0; lqd[j] = 0;
        if (j < len) {
            const size_t k = (size_t)env * nd + d0 + j;
            float a = T.actions_in[k];
            a = fminf(fmaxf(a, -C.clip_actions), C.clip_actions);
            act[j] = a;
            lact[j] = T.last_actions[k];
            lqd[j] = T.last_dof_vel[k];
        }
    }
    // arm: lane arm_chain holds the chain (segment variant: its first three joints, lane arm2 the rest) and evaluates the torque law
    constexpr bool SEG = is_segmented<LANES, NL>();
    const bool has_arm = ARM && (C.arm_chain >= 0);
    const bool is_arm = has_arm && (lane == C.arm_chain);
    const int arm2 = (SEG && has_arm) ? M->seg_child[C.arm_chain] : -1;
    const bool is_arm2 = arm2 >= 0 && lane == arm2;
    const int arm_d0 = has_arm ? M->chain_start[C.arm_chain] : 0, arm_len = has_arm ? M->chain_len[C.arm_chain] : 0;
    // the six arm values of a per-DOF array, complete in lane arm_chain (segment variant: a warp-wide exchange, every lane calls it)
    auto arm_gather = [&](const float* x, float* o6) {
        for (int j = 0; j < 6; j++) o6[j] = 0.0f;
#pragma unroll
        for (int j = 0; j < NL; j++)
            if (j < 6 && j < len) o6[j] = x[j];
        if (SEG) {
#pragma unroll
            for (int j = 0; j < NL; j++) {
                const float v = Grp<LANES>::bcast(x[j], arm2 >= 0 ? arm2 : lane);
                if (arm2 >= 0 && NL + j < 6) o6[NL + j] = v;
            }
        }
    };
    // arm lane: OSC torque from the stored mass-matrix / Jacobian slices, the end-effector velocity row and the live arm DOFs
    OscPrepared osc;
    if (SEG ? has_arm : is_arm) {
        float a6[6], dpose[6];
        arm_gather(act, a6);
        for (int j = 0; j < 6; j++) dpose[j] = (j < arm_len) ? a6[j] * C.arm_cmd_limit[j] / C.arm_action_scale : 0.0f;
        if (is_arm) osc_prepare(T.arm_mm + (size_t)env * 36, T.arm_jac + (size_t)env * 36, dpose, T.eef_state + (size_t)env * 13 + 7, C.arm_kp, osc);
    }
    // whole-chain variant: called by the arm lane; segment variant: called by every lane of the warp (has_arm is uniform)
    auto arm_osc = [&](float* out) {
        float qa[6], qda[6], eff[6], u[6];
        arm_gather(st.q, qa);
        arm_gather(st.qd, qda);
        for (int j = 0; j < 6; j++) { eff[j] = (j < arm_len) ? M->dof[arm_d0 + j].effort : 0.0f; u[j] = 0.0f; }
        if (is_arm) osc_apply(osc, qa, qda, C.arm_kp_null, eff, u);
#pragma unroll
        for (int j = 0; j < NL; j++)
            if (is_arm && j < 6 && j < len) out[j] = u[j];
        if (SEG) {
#pragma unroll
            for (int j = 0; j < NL; j++) {
                const float v = Grp<LANES>::bcast(NL + j < 6 ? u[NL + j] : 0.0f, C.arm_chain);
                if (is_arm2 && j < len) out[j] = v;
            }
        }
    };
    if (T.post_only == 2) {        // OSC probe (parity tests): one evaluation of the torque law, nothing else
        if (SEG ? has_arm : is_arm) {
            arm_osc(tq);
            if (valid && (is_arm || is_arm2))
                for (int j = 0; j < len; j++) T.torques[(size_t)env * nd + d0 + j] = tq[j];
        }
        return;
    }
    if (!T.post_only) {
        const EnvDr mu_shape = env_dr(A, valid ? env : 0, valid);
        const int total = C.decimation + C.extra_sim_steps;
#pragma unroll 1
        for (int it = 0; it < total; it++) {
            if (it < C.decimation) {      // tasks/anymal_terrain.py:444-445: fresh explicit PD torque, clipped
                if (SEG ? has_arm : is_arm) arm_osc(tq);      // tasks/useful_hound.py:704-716: operational-space torques for the arm, every decimation step
                if (!is_arm && !is_arm2) {
#pragma unroll
                    for (int j = 0; j < NL; j++) {
                        if (j < len) {
                            const float t = C.kp * (C.action_scale * act[j] + C.default_dof_pos[d0 + j] - st.q[j]) - C.kd * st.qd[j];
                            tq[j] = fminf(fmaxf(t, -C.torque_limit), C.torque_limit);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NL; j++) st.act[j] = tq[j];
            }
#pragma unroll 1
            for (int s = 0; s < A.P.substeps; s++)
                substep<LANES, NL, false, HF, false, (LANES == 4 && NL == 3)>(M, A.P, lane, len, d0, st, mu_shape,
                                                                             it == total - 1 && s == A.P.substeps - 1, sc, bf);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NL; j++)
            if (j < len) { tq[j] = T.torques[(size_t)env * nd + d0 + j]; st.frc[j] = A.dof_force[(size_t)env * nd + d0 + j]; }
        const int nb3p = M->n_bodies * 3;
        for (int i = lane; i < nb3p; i += LANES) bf[i] = A.contact[(size_t)env * nb3p + i];
        Grp<LANES>::sync();
    }
    // ---- post_physics_step (tasks/anymal_terrain.py:453-471) ----
    const long long common_step = T.step_ctr ? *T.step_ctr : T.common_step;
    if (C.push_interval > 0 && (common_step % C.push_interval) == 0) {      // push_robots :437-439
        st.rv.x = rand_range(-1.0f, 1.0f, terrain_uniform(T, T.push_override, 2, env, (unsigned)common_step, 3u, 0));
        st.rv.y = rand_range(-1.0f, 1.0f, terrain_uniform(T, T.push_override, 2, env, (unsigned)common_step, 3u, 1));
    }
    const V3 lin = quat_rotate_inverse(st.qx, st.qy, st.qz, st.qw, st.rv);
    const V3 ang = quat_rotate_inverse(st.qx, st.qy, st.qz, st.qw, st.rw);
    const V3 pg = quat_rotate_inverse(st.qx, st.qy, st.qz, st.qw, V3{0.0f, 0.0f, -1.0f});
    const V3 fwd = quat_apply_v(st.qx, st.qy, st.qz, st.qw, V3{1.0f, 0.0f, 0.0f});
    const float heading = atan2f(fwd.y, fwd.x);
    cmd[2] = fminf(fmaxf(0.5f * wrap_to_pi_f(cmd[3] - heading), -1.0f), 1.0f);
    // check_termination
    bool term = norm3(bf + C.base_body * 3) > 1.0f;
    if (C.hound_termination || !C.allow_knee_contacts)
        for (int k = 0; k < C.n_knee; k++) term = term || (norm3(bf + C.knee_bodies[k] * 3) > 1.0f);
    if (C.hound_termination)
        for (int k = 0; k < C.n_term_extra; k++) term = term || (norm3(bf + C.term_extra_bodies[k] * 3) > 1.0f);
    const bool reset = term || (progress >= C.max_episode_length - 1);
    const int nctrl = C.n_ctrl_dof > 0 ? C.n_ctrl_dof : nd;
    // compute_reward :315-382 -- per-lane partial sums, then one shuffle reduction each
    float s_tq = 0, s_jacc = 0, s_arate = 0, s_hip = 0, s_coll = 0, s_stumble = 0, s_air = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) {
        if (j < len) {
            s_tq += tq[j] * tq[j];
            const float dv = lqd[j] - st.qd[j], da = lact[j] - act[j];
            if (d0 + j < nctrl) s_jacc += dv * dv;
            s_arate += da * da;
            const int d = d0 + j;
            if (d == C.hip_dofs[0] || d == C.hip_dofs[1] || d == C.hip_dofs[2] || d == C.hip_dofs[3]) s_hip += fabsf(st.q[j] - C.default_dof_pos[d]);
        }
    }
    for (int k = lane; k < C.n_knee; k += LANES) s_coll += (norm3(bf + C.knee_bodies[k] * 3) > 1.0f) ? 1.0f : 0.0f;
    if (C.arm_chain >= 0)      // tasks/useful_hound.py:524-525
        for (int k = lane; k < C.n_term_extra; k += LANES) s_coll += (norm3(bf + C.term_extra_bodies[k] * 3) > 1.0f) ? 1.0f : 0.0f;
    for (int k = lane; k < C.n_feet; k += LANES) {
        const float* f = bf + C.feet_bodies[k] * 3;
        s_stumble += ((sqrtf(f[0] * f[0] + f[1] * f[1]) > 5.0f) && (fabsf(f[2]) < 1.0f)) ? 1.0f : 0.0f;
        const bool contact = f[2] > 1.0f;
        float air = T.feet_air_time[(size_t)env * 4 + k];
        const bool first = (air > 0.0f) && contact;
        air += C.dt;
        if (first) s_air += air - 0.5f;
        if (contact) air = 0.0f;
        if (valid) T.feet_air_time[(size_t)env * 4 + k] = air;
    }
    s_tq = Grp<LANES>::sum(s_tq); s_jacc = Grp<LANES>::sum(s_jacc); s_arate = Grp<LANES>::sum(s_arate); s_hip = Grp<LANES>::sum(s_hip);
    s_coll = Grp<LANES>::sum(s_coll); s_stumble = Grp<LANES>::sum(s_stumble); s_air = Grp<LANES>::sum(s_air);
    const float ex = cmd[0] - lin.x, ey = cmd[1] - lin.y, ez = cmd[2] - ang.z;
    const float r_lin_xy = expf(-(ex * ex + ey * ey) / 0.25f) * C.rew[1];
    const float r_ang_z = expf(-(ez * ez) / 0.25f) * C.rew[3];
    const float r_lin_z = lin.z * lin.z * C.rew[2];
    const float r_ang_xy = (ang.x * ang.x + ang.y * ang.y) * C.rew[4];
    const float r_orient = (pg.x * pg.x + pg.y * pg.y) * C.rew[5];
    const float dz = st.rp.z - C.base_height_target;
    const float r_height = dz * dz * C.rew[8];
    const float r_torque = s_tq * C.rew[6];
    const float r_jacc = s_jacc * C.rew[7];
    const float r_coll = s_coll * C.rew[10];
    const float r_stumble = s_stumble * C.rew[11];
    const float r_arate = s_arate * C.rew[12];
    float r_air = s_air * C.rew[9];
    if (!(sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]) > 0.1f)) r_air = 0.0f;
    const float r_hip = s_hip * C.rew[13];
    float rew = r_lin_xy + r_ang_z + r_lin_z + r_ang_xy + r_orient + r_height + r_torque + r_jacc + r_coll + r_arate + r_air + r_hip + r_stumble;
    rew = fmaxf(rew, 0.0f);
    if (reset && !timeout_prev) rew += C.rew[0];
    float q6[6], qd6[6];
    if (SEG ? has_arm : is_arm) { arm_gather(st.q, q6); arm_gather(st.qd, qd6); }
    if (valid) {
        store_state<NL>(A, env, lane, len, d0, st, false);
        const int nb3 = M->n_bodies * 3;
        for (int i = lane; i < nb3; i += LANES) A.contact[(size_t)env * nb3 + i] = bf[i];
#pragma unroll
        for (int j = 0; j < NL; j++) {
            if (j < len) {
                const size_t k = (size_t)env * nd + d0 + j;
                T.actions[k] = act[j];
                T.torques[k] = tq[j];
            }
        }
        if (is_arm) arm_refresh<NL>(A, T, env, arm_d0, arm_len, st, q6, qd6);      // refresh_jacobian / refresh_mass_matrix (useful_hound.py:731-732)
        if (lane == 0) {
            T.rew[env] = rew;
            T.reset[env] = reset ? 1 : 0;
            T.progress[env] = progress;
            T.commands[(size_t)env * 4 + 2] = cmd[2];
            T.resetw[env] = reset ? (cmd[0] * cmd[0] + cmd[1] * cmd[1]) : 0.0f;
            float* sc9 = T.scratch + (size_t)env * 9;
            sc9[0] = lin.x; sc9[1] = lin.y; sc9[2] = lin.z; sc9[3] = ang.x; sc9[4] = ang.y; sc9[5] = ang.z; sc9[6] = pg.x; sc9[7] = pg.y; sc9[8] = pg.z;
            // episode sums, key order of the reference dict (anymal_terrain.py:156-158)
            const float terms[13] = {r_lin_xy, r_lin_z, r_ang_z, r_ang_xy, r_orient, r_torque, r_jacc, r_height, r_air, r_coll, r_stumble, r_arate, r_hip};
            for (int k = 0; k < 13; k++) T.episode_sums[(size_t)k * A.n_envs + env] += terms[k];
        }
    }
}

// kernel 2: reset_idx (+ terrain curriculum), observations (+ height scan, noise), history, time-outs
// NSUB threads work on one environment: the first LANES of them are the chain lanes (reset draws, DOF columns), all of them
// share the element-wise passes (height scan, noise + clamp, report rows).  The device kernel runs 16 per environment -- the
// pass is a latency chain per element, so the extra warps are what hides it; the host emulator runs NSUB = LANES.
template <int LANES, int NL, int NSUB = LANES>
B2G_HD B2G_INL void terrain_post_thread(const SimArgs& A, const TerrainArgs& T, int env, int lane, bool valid, float curriculum_norm,
                                        float* rep_row = nullptr) {   // rep_row: 16 floats of this environment for the block's extras sums
    const DevModel* M = A.M;
    const b2g_terrain_cfg& C = T.cfg;
    const int nd = M->n_dof, N = A.n_envs;
    int len, d0;
    lane_span<LANES, NL>(M, lane, len, d0);
    const bool reset = T.reset[env] != 0;
    float root[13];
    for (int k = 0; k < 13; k++) root[k] = A.root[(size_t)env * 13 + k];
    float cmd[4] = {T.commands[(size_t)env * 4], T.commands[(size_t)env * 4 + 1], T.commands[(size_t)env * 4 + 2], T.commands[(size_t)env * 4 + 3]};
    long long progress = T.progress[env];
    const int rc = T.reset_count[env];
    float q[NL], qd[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) {
        q[j] = 0; qd[j] = 0;
        if (j < len) { q[j] = A.dof[2 * ((size_t)env * nd + d0 + j)]; qd[j] = A.dof[2 * ((size_t)env * nd + d0 + j) + 1]; }
    }
    float origin[3] = {T.env_origins[(size_t)env * 3], T.env_origins[(size_t)env * 3 + 1], T.env_origins[(size_t)env * 3 + 2]};
    long long level = T.terrain_levels[env];
    const long long type = T.terrain_types[env];
    // every lane has read the per-env inputs it replicates; only now may lane 0 overwrite them
    Grp<NSUB>::sync();
    const bool has_arm = C.arm_chain >= 0;
    const bool is_arm = has_arm && (lane == C.arm_chain || (is_segmented<LANES, NL>() && lane == M->seg_child[C.arm_chain]));
    const int nctrl = C.n_ctrl_dof > 0 ? C.n_ctrl_dof : nd;
    if (reset) {        // reset_idx, tasks/anymal_terrain.py:384-425 (useful_hound.py:569-637); draw order of the random calls
        const int ndraw = 2 * nctrl + 5 + (has_arm ? 6 : 0);
        const int arm_col = 2 * nctrl + (C.custom_origins ? 2 : 0);
#pragma unroll
        for (int j = 0; j < NL; j++) {
            if (j < len) {
                const int d = d0 + j;
                if (!is_arm) {
                    q[j] = C.default_dof_pos[d] * rand_range(0.5f, 1.5f, terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, d));
                    qd[j] = rand_range(-0.1f, 0.1f, terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, nctrl + d));
                } else {      // arm joints <- clamp(0 + noise * 2 (u - 0.5), lower, upper), zero velocity (useful_hound.py:594-601)
                    const float u = terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, arm_col + d - M->chain_start[C.arm_chain]);
                    const float v = 0.0f + C.arm_dof_noise * 2.0f * (u - 0.5f);
                    q[j] = fminf(fmaxf(v, M->dof[d].lower), M->dof[d].upper);
                    qd[j] = 0.0f;
                }
            }
        }
        int col = 2 * nctrl;
        if (C.custom_origins) {
            if (T.init_done && C.curriculum) {      // update_terrain_level :427-435
                const float dx = root[0] - origin[0], dy = root[1] - origin[1];
                const float dist = sqrtf(dx * dx + dy * dy);
                if (dist < curriculum_norm * C.max_episode_length_s * 0.25f) level -= 1;
                if (dist > C.env_length / 2) level += 1;
                if (level < 0) level = 0;
                level = level % C.env_rows;
                const float* o = T.terrain_origins + ((size_t)level * C.env_cols + (size_t)type) * 3;
                origin[0] = o[0]; origin[1] = o[1]; origin[2] = o[2];
                if (valid && lane == 0) {
                    T.terrain_levels[env] = level;
                    T.env_origins[(size_t)env * 3] = origin[0]; T.env_origins[(size_t)env * 3 + 1] = origin[1]; T.env_origins[(size_t)env * 3 + 2] = origin[2];
                }
            }
            for (int k = 0; k < 13; k++) root[k] = C.init_root[k];
            root[0] += origin[0]; root[1] += origin[1]; root[2] += origin[2];
            root[0] += rand_range(-0.5f, 0.5f, terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, col));
            root[1] += rand_range(-0.5f, 0.5f, terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, col + 1));
            col += 2;
        } else {
            for (int k = 0; k < 13; k++) root[k] = C.init_root[k];
        }
        if (has_arm) col += 6;
        cmd[0] = rand_range(C.cmd_x[0], C.cmd_x[1], terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, col));
        cmd[1] = rand_range(C.cmd_y[0], C.cmd_y[1], terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, col + 1));
        cmd[3] = rand_range(C.cmd_yaw[0], C.cmd_yaw[1], terrain_uniform(T, T.reset_override, ndraw, env, (unsigned)rc, 1u, col + 2));
        if (!(sqrtf(cmd[0] * cmd[0] + cmd[1] * cmd[1]) > 0.25f)) { cmd[0] = 0; cmd[1] = 0; cmd[2] = 0; cmd[3] = 0; }   // :412 (all four columns)
        progress = 0;
    }
    // ---- observations (tasks/anymal_terrain.py:302-313) ----
    const int nhp = C.n_hx * C.n_hy;
    const int no = 12 + 2 * nctrl + nhp + nd + (has_arm ? 10 : 0);
    float* o = T.obs + (size_t)env * no;
    float* oc = T.obs_clamped + (size_t)env * no;
    const unsigned step = (unsigned)(T.step_ctr ? *T.step_ctr : T.common_step);
    // values are stored raw; the noise pass below adds the uniform noise and writes the clamped copy.  The noise stream draws
    // column i from word i & 3 of Philox block i >> 2, so that pass walks the row in groups of four columns: one Philox
    // evaluation per group instead of one per column (same stream, same values)
    auto put = [&](int idx, float val, float nscale) {
        (void)nscale;
        if (valid) o[idx] = val;
    };
    if (has_arm && lane == 0) {      // tasks/useful_hound.py:493-496: end-effector position, orientation, arm command
        const float* e = T.eef_state + (size_t)env * 13;
        const int b0 = 12 + 2 * nctrl + nhp + nd;
        for (int k = 0; k < 7; k++) put(b0 + k, e[k], 0.0f);
        for (int k = 0; k < 3; k++) put(b0 + 7 + k, T.arm_commands[(size_t)env * 3 + k], 0.0f);
    }
    if (lane == 0) {
        const float* s9 = T.scratch + (size_t)env * 9;
        for (int k = 0; k < 3; k++) put(k, s9[k] * C.lin_vel_scale, C.noise_lin_vel);
        for (int k = 0; k < 3; k++) put(3 + k, s9[3 + k] * C.ang_vel_scale, C.noise_ang_vel);
        for (int k = 0; k < 3; k++) put(6 + k, s9[6 + k], C.noise_gravity);
        put(9, cmd[0] * C.lin_vel_scale, 0.0f);
        put(10, cmd[1] * C.lin_vel_scale, 0.0f);
        put(11, cmd[2] * C.ang_vel_scale, 0.0f);
    }
#pragma unroll
    for (int j = 0; j < NL; j++) {
        if (j < len) {
            const int d = d0 + j;
            const size_t k = (size_t)env * nd + d;
            const float a = T.actions[k];
            if (d < nctrl) {
                put(12 + d, q[j] * C.dof_pos_scale, C.noise_dof_pos);
                put(12 + nctrl + d, qd[j] * C.dof_vel_scale, C.noise_dof_vel);
            }
            put(12 + 2 * nctrl + nhp + d, a, 0.0f);
            if (valid) {
                T.last_actions[k] = a;          // :484-485 (post-reset values for envs that reset)
                T.last_dof_vel[k] = qd[j];
                if (reset) { A.dof[2 * k] = q[j]; A.dof[2 * k + 1] = qd[j]; }
            }
        }
    }
    // height scan: yaw-only rotation of the grid, truncating index, min of two diagonal samples (:515-538)
    {
        float yz = root[5], yw = root[6];
        const float yn = fmaxf(sqrtf(yz * yz + yw * yw), 1e-9f);
        yz /= yn; yw /= yn;
        for (int p = lane; p < nhp; p += NSUB) {
            float h = 0.0f;
            if (T.height_samples) {
                const V3 pt = quat_apply_v(0.0f, 0.0f, yz, yw, V3{C.hx[p / C.n_hy], C.hy[p % C.n_hy], 0.0f});
                // (points + border).long() then clip (anymal_terrain.py:529-533); the float clamp keeps the 32-bit conversion exact
                int px = (int)fminf(fmaxf((pt.x + root[0] + C.border_size) / C.hscale, -1.0f), (float)C.hs_rows);
                int py = (int)fminf(fmaxf((pt.y + root[1] + C.border_size) / C.hscale, -1.0f), (float)C.hs_cols);
                px = px < 0 ? 0 : (px > C.hs_rows - 2 ? C.hs_rows - 2 : px);
                py = py < 0 ? 0 : (py > C.hs_cols - 2 ? C.hs_cols - 2 : py);
                const short h1 = T.height_samples[(size_t)px * C.hs_cols + py], h2 = T.height_samples[(size_t)(px + 1) * C.hs_cols + py + 1];
                h = (float)(h1 < h2 ? h1 : h2) * C.vscale;
            }
            if (valid) T.measured[(size_t)env * nhp + p] = h;
            const float v = fminf(fmaxf(root[2] - 0.5f - h, -1.0f), 1.0f) * C.height_meas_scale;
            put(12 + 2 * nctrl + p, v, C.noise_height);
        }
    }
    // noise + clamp pass (noise scales per column block: anymal_terrain.py:163-178)
    Grp<NSUB>::sync();
    {
        const int c_pos = 12, c_vel = 12 + nctrl, c_h = 12 + 2 * nctrl, c_end = c_h + nhp;
        for (int g = lane; g * 4 < no; g += NSUB) {
            unsigned w[4] = {0u, 0u, 0u, 0u};
            const bool noisy = C.add_noise && g * 4 < c_end;   // columns past the height scan (actions, arm state) carry no noise
            if (noisy && !T.noise_override)
                philox4x32((unsigned)env, step, (unsigned)g, 2u, (unsigned)(T.seed & 0xffffffffull), (unsigned)(T.seed >> 32), w);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int idx = g * 4 + c;
                if (idx < no) {
                    const float nscale = idx < 3 ? C.noise_lin_vel : idx < 6 ? C.noise_ang_vel : idx < 9 ? C.noise_gravity : idx < c_pos ? 0.0f
                                       : idx < c_vel ? C.noise_dof_pos : idx < c_h ? C.noise_dof_vel : idx < c_end ? C.noise_height : 0.0f;
                    float val = o[idx];
                    if (C.add_noise && nscale != 0.0f) {
                        const float u = T.noise_override ? T.noise_override[(size_t)env * no + idx] : (float)(w[c] >> 8) * (1.0f / 16777216.0f);
                        val += (2.0f * u - 1.0f) * nscale;
                    }
                    if (valid) { o[idx] = val; oc[idx] = fminf(fmaxf(val, -C.clip_obs), C.clip_obs); }
                }
            }
        }
    }
    if (valid) {
        for (int k = lane; k < C.n_feet; k += NSUB)
            if (reset) T.feet_air_time[(size_t)env * 4 + k] = 0.0f;
        for (int k = lane; k < 13; k += NSUB) {
            float* es = T.episode_sums + (size_t)k * N + env;
            const float rv = reset ? *es : 0.0f;
            T.report[(size_t)k * N + env] = rv;
            if (rep_row) rep_row[k] = rv;
            if (reset) *es = 0.0f;
        }
        if (rep_row && lane == 0) { rep_row[13] = reset ? 1.0f : 0.0f; rep_row[14] = (float)level; }
        if (lane == 0) {
            if (reset) {
                for (int k = 0; k < 13; k++) A.root[(size_t)env * 13 + k] = root[k];
                for (int k = 0; k < 4; k++) T.commands[(size_t)env * 4 + k] = cmd[k];
                T.reset_count[env] = rc + 1;
            }
            T.progress[env] = progress;
            T.timeout[env] = (progress >= C.max_episode_length - 1 && reset) ? 1 : 0;
        }
    }
}


// refresh_mass_matrix_tensors: joint-space block of the mass matrix, (nd x nd) per environment (Isaac Gym convention used by
// tasks/useful_hound.py:452-455).  DOFs of different chains couple only through the root, which is not part of this block,
// so the matrix is block diagonal over the chains.  One thread per environment.
B2G_HD inline void mass_matrix_env(const DevModel* M, const float* root, const float* dof, float* out, float mass_scale = 1.0f,
                                   const float* link_scale = nullptr) {
    const int nd = M->n_dof;
    for (int i = 0; i < nd * nd; i++) out[i] = 0.0f;
    for (int c = 0; c < M->n_chains; c++) {
        const int d0 = M->chain_start[c], n = M->chain_len[c];
        constexpr int LD = B2G_MAX_FIXED_CHAIN_LEN;
        float q[LD], qd[LD], mm[LD * LD];
        for (int j = 0; j < n; j++) { q[j] = dof[2 * (d0 + j)]; qd[j] = 0.0f; }
        V3 p; M3 r; SV v;
        chain_crba(M, d0, n, root + 3, q, mm, -1, &p, &r, qd, V3{0, 0, 0}, V3{0, 0, 0}, &v, mass_scale, link_scale, LD);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) out[(d0 + i) * nd + d0 + j] = mm[i * LD + j];
    }
}

// refresh_jacobian_tensors: geometric Jacobian of every API body, rows = (linear 3, angular 3) of the body-frame origin in the
// world frame; columns = [6 base columns (linear, angular) for a floating base] + one per DOF.  Floating base: (nb, 6, 6+nd);
// fixed base: (nb-1, 6, nd) with the root body left out (Isaac Gym convention).  One thread per environment.
B2G_HD inline void jacobian_env(const DevModel* M, const float* root, const float* dof, float* out) {
    const int nd = M->n_dof, nb = M->n_bodies;
    const int ncol = nd + (M->fixed_base ? 0 : 6), row0 = M->fixed_base ? 1 : 0;
    M3 Rl[B2G_MAX_LINKS];
    V3 pl[B2G_MAX_LINKS], axw[B2G_MAX_DOF], pj[B2G_MAX_DOF];
    Rl[0] = quat_to_m3(root[3], root[4], root[5], root[6]);
    pl[0] = V3{0, 0, 0};
    for (int c = 0; c < M->n_chains; c++)
        for (int j = 0; j < M->chain_len[c]; j++) {
            const int d = M->chain_start[c] + j, l = d + 1, p = (j == 0) ? 0 : l - 1;
            const DevDof& D = M->dof[d];
            M3 jr;
            for (int k = 0; k < 9; k++) jr.m[k] = D.jrot[k];
            const V3 ax = V3{D.axis[0], D.axis[1], D.axis[2]};
            const M3 RJ = mul(Rl[p], jr);
            pj[d] = pl[p] + mul(Rl[p], V3{D.jpos[0], D.jpos[1], D.jpos[2]});
            axw[d] = mul(RJ, ax);
            if (D.type == B2G_JOINT_REVOLUTE) { Rl[l] = mul(RJ, axis_angle_m3(ax, dof[2 * d])); pl[l] = pj[d]; }
            else { Rl[l] = RJ; pl[l] = pj[d] + axw[d] * dof[2 * d]; }
        }
    for (int b = row0; b < nb; b++) {
        float* J = out + (size_t)(b - row0) * 6 * ncol;
        for (int i = 0; i < 6 * ncol; i++) J[i] = 0.0f;
        const int l = M->body_link[b];
        const V3 pb = pl[l] + mul(Rl[l], V3{M->body_pos[b][0], M->body_pos[b][1], M->body_pos[b][2]});
        int c0 = 0;
        if (!M->fixed_base) {
            for (int i = 0; i < 6; i++) J[i * ncol + i] = 1.0f;
            J[0 * ncol + 4] = pb.z; J[0 * ncol + 5] = -pb.y; J[1 * ncol + 3] = -pb.z; J[1 * ncol + 5] = pb.x; J[2 * ncol + 3] = pb.y; J[2 * ncol + 4] = -pb.x;
            c0 = 6;
        }
        if (l > 0) {
            // DOFs on the path root -> link l: those of its chain up to and including DOF l-1
            int dfirst = 0;
            for (int c = 0; c < M->n_chains; c++)
                if (l - 1 >= M->chain_start[c] && l - 1 < M->chain_start[c] + M->chain_len[c]) dfirst = M->chain_start[c];
            for (int d = dfirst; d <= l - 1; d++) {
                if (M->dof[d].type == B2G_JOINT_REVOLUTE) {
                    const V3 lin = cross(axw[d], pb - pj[d]);
                    J[0 * ncol + c0 + d] = lin.x; J[1 * ncol + c0 + d] = lin.y; J[2 * ncol + c0 + d] = lin.z;
                    J[3 * ncol + c0 + d] = axw[d].x; J[4 * ncol + c0 + d] = axw[d].y; J[5 * ncol + c0 + d] = axw[d].z;
                } else {
                    J[0 * ncol + c0 + d] = axw[d].x; J[1 * ncol + c0 + d] = axw[d].y; J[2 * ncol + c0 + d] = axw[d].z;
                }
            }
        }
    }
}

// refresh_rigid_body_state_tensor: world pose + velocity of every API body (N,nb,13). One thread per env.
B2G_HD B2G_INL void body_state_env(const DevModel* M, const float* root, const float* dof, float* out) {
    const int nd = M->n_dof;
    M3 Rl[B2G_MAX_LINKS];
    V3 pl[B2G_MAX_LINKS];
    SV vl[B2G_MAX_LINKS];
    Rl[0] = quat_to_m3(root[3], root[4], root[5], root[6]);
    pl[0] = V3{0, 0, 0};
    vl[0] = M->fixed_base ? sv0() : SV{V3{root[10], root[11], root[12]}, V3{root[7], root[8], root[9]}};
    for (int c = 0; c < M->n_chains; c++) {
        for (int j = 0; j < M->chain_len[c]; j++) {
            const int d = M->chain_start[c] + j, l = d + 1, p = (j == 0) ? 0 : l - 1;
            const DevDof& D = M->dof[d];
            M3 jr;
            for (int k = 0; k < 9; k++) jr.m[k] = D.jrot[k];
            const V3 ax = V3{D.axis[0], D.axis[1], D.axis[2]};
            const M3 RJ = mul(Rl[p], jr);
            const V3 pj = pl[p] + mul(Rl[p], V3{D.jpos[0], D.jpos[1], D.jpos[2]});
            const V3 axw = mul(RJ, ax);
            SV S;
            if (D.type == B2G_JOINT_REVOLUTE) { Rl[l] = mul(RJ, axis_angle_m3(ax, dof[2 * d])); pl[l] = pj; S = SV{axw, cross(pj, axw)}; }
            else { Rl[l] = RJ; pl[l] = pj + axw * dof[2 * d]; S = SV{V3{0, 0, 0}, axw}; }
            vl[l] = vl[p] + S * dof[2 * d + 1];
        }
    }
    (void)nd;
    for (int b = 0; b < M->n_bodies; b++) {
        const int l = M->body_link[b];
        const V3 off = mul(Rl[l], V3{M->body_pos[b][0], M->body_pos[b][1], M->body_pos[b][2]});
        const V3 p = pl[l] + off;
        const M3 R = mul(Rl[l], quat_to_m3(M->body_quat[b][0], M->body_quat[b][1], M->body_quat[b][2], M->body_quat[b][3]));
        // rotation matrix -> quaternion (xyzw)
        float qx, qy, qz, qw;
        const float tr = R.m[0] + R.m[4] + R.m[8];
        if (tr > 0.0f) { float s = sqrtf(tr + 1.0f) * 2.0f; qw = 0.25f * s; qx = (R.m[7] - R.m[5]) / s; qy = (R.m[2] - R.m[6]) / s; qz = (R.m[3] - R.m[1]) / s; }
        else if (R.m[0] > R.m[4] && R.m[0] > R.m[8]) { float s = sqrtf(1.0f + R.m[0] - R.m[4] - R.m[8]) * 2.0f; qw = (R.m[7] - R.m[5]) / s; qx = 0.25f * s; qy = (R.m[1] + R.m[3]) / s; qz = (R.m[2] + R.m[6]) / s; }
        else if (R.m[4] > R.m[8]) { float s = sqrtf(1.0f + R.m[4] - R.m[0] - R.m[8]) * 2.0f; qw = (R.m[2] - R.m[6]) / s; qx = (R.m[1] + R.m[3]) / s; qy = 0.25f * s; qz = (R.m[5] + R.m[7]) / s; }
        else { float s = sqrtf(1.0f + R.m[8] - R.m[0] - R.m[4]) * 2.0f; qw = (R.m[3] - R.m[1]) / s; qx = (R.m[2] + R.m[6]) / s; qy = (R.m[5] + R.m[7]) / s; qz = 0.25f * s; }
        const V3 v = vl[l].v + cross(vl[l].w, p);   // velocity of the body-frame origin
        float* o = out + (size_t)b * 13;
        o[0] = root[0] + p.x; o[1] = root[1] + p.y; o[2] = root[2] + p.z;
        o[3] = qx; o[4] = qy; o[5] = qz; o[6] = qw;
        o[7] = v.x; o[8] = v.y; o[9] = v.z;
        o[10] = vl[l].w.x; o[11] = vl[l].w.y; o[12] = vl[l].w.z;
    }
}

}  // namespace b2g
