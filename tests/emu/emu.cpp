// TEST INFRASTRUCTURE: host lane emulator.  Compiles the *same* per-thread kernel bodies as the CUDA
// library (isaacgymenv_b200/csrc/b2g_threads.cuh) for the CPU, running the LANES cooperating lanes of an
// environment as lock-stepped host threads.  It lets the CPU test-suite (-m "not gpu") exercise the
// kernel logic against the oracle without a GPU.  It is NOT part of the product and is never loaded by
// isaacgymenv_b200; the product has no CPU path.
#define B2G_HOST_EMU 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "b2g_host_pack.h"
#include "b2g_threads.cuh"

namespace b2g {
thread_local EmuCtx emu_ctx;
long long emu_hfc_tests = 0, emu_hfc_skips = 0;
}
using namespace b2g;

namespace {

struct Variant { int lanes, nl; bool fixed; bool seg = false; };

Variant pick(const b2g_model& m) {
    int maxlen = 0;
    for (int c = 0; c < m.n_chains; c++) maxlen = m.chain_len[c] > maxlen ? m.chain_len[c] : maxlen;
    int minlen = 1 << 20;
    for (int c = 0; c < m.n_chains; c++) minlen = m.chain_len[c] < minlen ? m.chain_len[c] : minlen;
    // the specialised variants assume FULL chains (every lane has exactly NL links)
    if (m.fixed_base && m.n_chains == 1 && maxlen == 2) return {1, 2, true};
    if (!m.fixed_base && m.n_chains == 4 && maxlen == 3 && minlen == 3) return {4, 3, false};
    int pieces = m.n_chains;      // same rule as b200gym.cu::pick_variant
    for (int c = 0; c < m.n_chains; c++) pieces += m.chain_len[c] > kSegLinks ? 1 : 0;
    const char* seg = getenv("B2G_SEGMENTS");
    if (seg && seg[0] == '1' && !m.fixed_base && m.n_chains > 0 && pieces <= B2G_MAX_CHAINS) return {8, 3, false, true};
    if (m.fixed_base) return {8, B2G_MAX_FIXED_CHAIN_LEN, true};
    return {8, 6, false};
}

template <class F>
void run_group(int lanes, F&& body) {
    EmuGroup g;
    g.lanes = lanes; g.count = 0; g.sense = 0;
    std::vector<std::thread> th;
    for (int l = 0; l < lanes; l++)
        th.emplace_back([&, l]() {
            emu_ctx.g = &g; emu_ctx.lane = l; emu_ctx.local_sense = 0;
            body(l);
        });
    for (auto& t : th) t.join();
}

template <int LANES, int NL, bool FIXED, bool HF>
void sim_env(const SimArgs& A, int env) {
    std::vector<float> scratch((size_t)LANES * MAXC * CF_COUNT), bf((size_t)A.M->n_bodies * 3);
    std::vector<float> links(link_store_floats(1, B2G_MAX_DOF)), anc((size_t)B2G_MAX_CHAINS * kAncFloats);
    run_group(LANES, [&](int lane) {
        ScratchStrided sc{scratch.data() + lane, LANES}; sc.links = links.data(); sc.anc = anc.data();
        simulate_thread<LANES, NL, FIXED, HF>(A, env, lane, true, sc, bf.data());
    });
}

template <int LANES, int NL, bool FIXED>
void probe_env(const SimArgs& A, int env, float* qdd, float* a0) {
    std::vector<float> scratch((size_t)LANES * MAXC * CF_COUNT), bf((size_t)A.M->n_bodies * 3);
    std::vector<float> links(link_store_floats(1, B2G_MAX_DOF)), anc((size_t)B2G_MAX_CHAINS * kAncFloats);
    run_group(LANES, [&](int lane) {
        const DevModel* M = A.M;
        int len, d0;
        lane_span<LANES, NL>(M, lane, len, d0);
        LaneState<NL> st;
        load_state<NL>(A, env, len, d0, st);
        for (int j = 0; j < NL; j++) if (j < len) st.act[j] = A.actuation[(size_t)env * M->n_dof + d0 + j];
        ScratchStrided sc{scratch.data() + lane, LANES}; sc.links = links.data(); sc.anc = anc.data();
        substep<LANES, NL, FIXED, false, true>(M, A.P, lane, len, d0, st, env_dr(A, env, true, false), false, sc, bf.data());
        for (int j = 0; j < NL; j++) if (j < len) qdd[(size_t)env * M->n_dof + d0 + j] = st.frc[j];
        if (lane == 0) {
            float* o = a0 + (size_t)env * 6;
            o[0] = st.rw.x; o[1] = st.rw.y; o[2] = st.rw.z; o[3] = st.rv.x; o[4] = st.rv.y; o[5] = st.rv.z;
        }
    });
}

template <int LANES, int NL, bool HF>
void anymal_env(const SimArgs& A, const TaskArgs& T, int env, int mode) {
    std::vector<float> scratch((size_t)LANES * MAXC * CF_COUNT), bf((size_t)A.M->n_bodies * 3);
    std::vector<float> links(link_store_floats(1, B2G_MAX_DOF)), anc((size_t)B2G_MAX_CHAINS * kAncFloats);
    run_group(LANES, [&](int lane) {
        ScratchStrided sc{scratch.data() + lane, LANES}; sc.links = links.data(); sc.anc = anc.data();
        if (mode == 0) anymal_reset_all_thread<LANES, NL>(A, T, env, lane, true);
        else anymal_step_thread<LANES, NL, HF>(A, T, env, lane, true, sc, bf.data());
    });
}

}  // namespace

// per-env domain-randomisation scales (N,4) used by the next emu_simulate / emu_forward_dynamics calls; null = ones
static const float* g_env_scale = nullptr;
static const float* g_link_scale = nullptr;     // (N, nd + 1, 6) per-link rows (B2G_T_LINK_SCALE); null = none
// contact statistics of the emulated sub-steps (same counters as b2g_sim_contact_stats)
static unsigned long long g_stats[4] = {0, 0, 0, 0};

extern "C" {

void emu_set_env_scale(const float* p) { g_env_scale = p; }
void emu_set_link_scale(const float* p) { g_link_scale = p; }

void emu_contact_stats(long long* out, int reset) {
    for (int i = 0; i < 4; i++) out[i] = (long long)g_stats[i];
    if (reset) for (int i = 0; i < 4; i++) g_stats[i] = 0;
}

void emu_hfc_stats(long long* out, int reset) {
    out[0] = b2g::emu_hfc_tests; out[1] = b2g::emu_hfc_skips;
    if (reset) b2g::emu_hfc_tests = b2g::emu_hfc_skips = 0;
}

int emu_simulate(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, const b2g_heightfield* hf,
                 const int16_t* hfs, const float* friction, int n_envs, float* root, float* dof, const float* target,
                 const float* actuation, float* dof_force, float* contact) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    SimArgs A;
    A.M = dm;
    pack_dev_params(*sp, hf, hfs, A.P);
    A.P.stats = g_stats;
    std::vector<float> hfc;   // coarse heightfield bound, as the library builds it (B2G_NO_HFC=1: exhaustive candidate tests)
    if (hf && hfs && !(getenv("B2G_NO_HFC") && getenv("B2G_NO_HFC")[0] == '1')) {
        build_hf_coarse(hfs, hf->rows, hf->cols, hf->horizontal_scale, hf->vertical_scale, max_link_radius(*dm), hfc, A.P.hfc_rows, A.P.hfc_cols);
        A.P.hfc = hfc.data();
    }
    A.n_envs = n_envs; A.root = root; A.dof = dof; A.target = target; A.actuation = actuation;
    A.dof_force = dof_force; A.contact = contact; A.friction = friction; A.env_scale = g_env_scale; A.link_scale = g_link_scale;
    const Variant v = pick(*m);
    const bool HFm = hf && hfs;
    for (int e = 0; e < n_envs; e++) {
        if (v.lanes == 1) sim_env<1, 2, true, false>(A, e);
        else if (v.lanes == 4) { if (HFm) sim_env<4, 3, false, true>(A, e); else sim_env<4, 3, false, false>(A, e); }
        else if (v.fixed) sim_env<8, 7, true, false>(A, e);
        else if (v.seg) { if (HFm) sim_env<8, 3, false, true>(A, e); else sim_env<8, 3, false, false>(A, e); }
        else { if (HFm) sim_env<8, 6, false, true>(A, e); else sim_env<8, 6, false, false>(A, e); }
    }
    delete dm;
    return 0;
}

int emu_forward_dynamics(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, int n_envs, float* root,
                         float* dof, const float* tau, float* qdd, float* a0) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    SimArgs A;
    A.M = dm;
    pack_dev_params(*sp, nullptr, nullptr, A.P);
    A.P.stats = g_stats;
    A.n_envs = n_envs; A.root = root; A.dof = dof; A.target = tau; A.actuation = tau; A.dof_force = nullptr; A.contact = nullptr; A.friction = nullptr; A.env_scale = g_env_scale; A.link_scale = g_link_scale;
    const Variant v = pick(*m);
    for (int e = 0; e < n_envs; e++) {
        if (v.lanes == 1) probe_env<1, 2, true>(A, e, qdd, a0);
        else if (v.lanes == 4) probe_env<4, 3, false>(A, e, qdd, a0);
        else if (v.fixed) probe_env<8, 7, true>(A, e, qdd, a0);
        else if (v.seg) probe_env<8, 3, false>(A, e, qdd, a0);
        else probe_env<8, 6, false>(A, e, qdd, a0);
    }
    delete dm;
    return 0;
}

// mode 0: reset_all, mode 1: step, mode 2: post_physics_step only
int emu_anymal(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, const b2g_anymal_cfg* cfg, int mode,
               int n_envs, float* root, float* dof, float* dof_force, float* contact, const float* actions_in, float* obs,
               float* obs_clamped, float* rew, long long* reset, long long* progress, long long* timeout, float* commands,
               float* actions, int* reset_count, const float* rand_override) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    SimArgs A;
    A.M = dm;
    pack_dev_params(*sp, nullptr, nullptr, A.P);
    A.P.stats = g_stats;
    A.n_envs = n_envs; A.root = root; A.dof = dof; A.target = nullptr; A.actuation = nullptr; A.dof_force = dof_force;
    A.contact = contact; A.friction = nullptr;
    TaskArgs T;
    T.cfg = *cfg; T.seed = cfg->seed; memset(&T.ccfg, 0, sizeof(T.ccfg)); memset(&T.hcfg, 0, sizeof(T.hcfg)); T.actions_in = actions_in; T.obs = obs; T.obs_clamped = obs_clamped; T.rew = rew; T.reset = reset;
    T.progress = progress; T.timeout = timeout; T.commands = commands; T.actions = actions; T.reset_count = reset_count;
    T.rand_override = rand_override;
    T.post_only = (mode == 2);
    const Variant v = pick(*m);
    if (v.fixed) { delete dm; return -2; }
    for (int e = 0; e < n_envs; e++) {
        if (v.lanes == 4) anymal_env<4, 3, false>(A, T, e, mode);
        else if (v.seg) anymal_env<8, 3, false>(A, T, e, mode);
        else anymal_env<8, 6, false>(A, T, e, mode);
    }
    delete dm;
    return 0;
}

// rough-terrain task: mode 1 = step, mode 2 = post_physics_step only.  ptrs: see TerrainArgs.
struct EmuTerrainBufs {
    float *root, *dof, *dof_force, *contact;
    const float* actions_in;
    float *obs, *obs_clamped, *rew;
    long long *reset, *progress, *timeout;
    float *commands, *actions, *torques, *last_actions, *last_dof_vel, *feet_air_time, *episode_sums, *env_origins;
    long long *terrain_levels, *terrain_types;
    const float* terrain_origins;
    const short* height_samples;
    float *scratch, *resetw, *report, *measured;
    float *arm_mm, *arm_jac, *eef_state, *arm_commands;
    int* reset_count;
    const float *reset_override, *noise_override, *push_override;
    const float* friction;
};

int emu_terrain(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, const b2g_heightfield* hf, const int16_t* hfs,
                const b2g_terrain_cfg* cfg, int mode, int n_envs, long long common_step, int init_done, const EmuTerrainBufs* B) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    SimArgs A;
    A.M = dm;
    pack_dev_params(*sp, hf, hfs, A.P);
    A.P.stats = g_stats;
    std::vector<float> hfc;   // coarse heightfield bound, as the library builds it (B2G_NO_HFC=1: exhaustive candidate tests)
    if (hf && hfs && !(getenv("B2G_NO_HFC") && getenv("B2G_NO_HFC")[0] == '1')) {
        build_hf_coarse(hfs, hf->rows, hf->cols, hf->horizontal_scale, hf->vertical_scale, max_link_radius(*dm), hfc, A.P.hfc_rows, A.P.hfc_cols);
        A.P.hfc = hfc.data();
    }
    A.n_envs = n_envs; A.root = B->root; A.dof = B->dof; A.target = nullptr; A.actuation = nullptr; A.dof_force = B->dof_force;
    A.contact = B->contact; A.friction = B->friction;
    TerrainArgs T;
    T.cfg = *cfg; T.actions_in = B->actions_in; T.obs = B->obs; T.obs_clamped = B->obs_clamped; T.rew = B->rew; T.reset = B->reset;
    T.progress = B->progress; T.timeout = B->timeout; T.commands = B->commands; T.actions = B->actions; T.torques = B->torques;
    T.last_actions = B->last_actions; T.last_dof_vel = B->last_dof_vel; T.feet_air_time = B->feet_air_time; T.episode_sums = B->episode_sums;
    T.env_origins = B->env_origins; T.terrain_levels = B->terrain_levels; T.terrain_types = B->terrain_types; T.terrain_origins = B->terrain_origins;
    T.height_samples = B->height_samples; T.scratch = B->scratch; T.resetw = B->resetw; T.report = B->report; T.measured = B->measured;
    T.arm_mm = B->arm_mm; T.arm_jac = B->arm_jac; T.eef_state = B->eef_state; T.arm_commands = B->arm_commands;
    T.reset_count = B->reset_count; T.reset_override = B->reset_override; T.noise_override = B->noise_override; T.push_override = B->push_override;
    T.common_step = common_step; T.init_done = init_done; T.post_only = (mode == 2) ? 1 : (mode == 3 ? 2 : 0); T.seed = cfg->seed;
    const Variant v = pick(*m);
    if (v.fixed) { delete dm; return -2; }
    const bool HFm = hf && hfs;
    for (int e = 0; e < n_envs; e++) {
        std::vector<float> scratch((size_t)v.lanes * MAXC * CF_COUNT), bf((size_t)m->n_bodies * 3);
    std::vector<float> links(link_store_floats(1, B2G_MAX_DOF)), anc((size_t)B2G_MAX_CHAINS * kAncFloats);
        run_group(v.lanes, [&](int lane) {
            ScratchStrided sc{scratch.data() + lane, v.lanes}; sc.links = links.data(); sc.anc = anc.data();
            if (v.lanes == 4) { if (HFm) terrain_phys_thread<4, 3, true>(A, T, e, lane, true, sc, bf.data()); else terrain_phys_thread<4, 3, false>(A, T, e, lane, true, sc, bf.data()); }
            else if (v.seg) { if (HFm) terrain_phys_thread<8, 3, true>(A, T, e, lane, true, sc, bf.data()); else terrain_phys_thread<8, 3, false>(A, T, e, lane, true, sc, bf.data()); }
            else { if (HFm) terrain_phys_thread<8, 6, true>(A, T, e, lane, true, sc, bf.data()); else terrain_phys_thread<8, 6, false>(A, T, e, lane, true, sc, bf.data()); }
        });
    }
    if (mode == 3) { delete dm; return 0; }      // OSC probe: first kernel only
    float cnorm = 0.0f;
    if (cfg->custom_origins && cfg->curriculum && init_done) {
        // same summation order as the kernel: 64 strided partial sums, then a pairwise tree
        float red[64];
        for (int t = 0; t < 64; t++) { float acc = 0; for (int i = t; i < n_envs; i += 64) acc += B->resetw[i]; red[t] = acc; }
        for (int o = 32; o > 0; o >>= 1) for (int t = 0; t < o; t++) red[t] += red[t + o];
        cnorm = sqrtf(red[0]);
    }
    for (int e = 0; e < n_envs; e++) {
        run_group(v.lanes, [&](int lane) {
            if (v.lanes == 4) terrain_post_thread<4, 3>(A, T, e, lane, true, cnorm);
            else if (v.seg) terrain_post_thread<8, 3>(A, T, e, lane, true, cnorm);
            else terrain_post_thread<8, 6>(A, T, e, lane, true, cnorm);
        });
    }
    delete dm;
    return 0;
}

// refresh_jacobian_tensors / refresh_mass_matrix_tensors for n_envs environments
int emu_jac_mm(const b2g_model* m, const b2g_dof_props* dp, int n_envs, const float* root, const float* dof, float* jac, int jac_per_env, float* mm) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    for (int e = 0; e < n_envs; e++) {
        jacobian_env(dm, root + (size_t)e * 13, dof + (size_t)e * m->n_dof * 2, jac + (size_t)e * jac_per_env);
        mass_matrix_env(dm, root + (size_t)e * 13, dof + (size_t)e * m->n_dof * 2, mm + (size_t)e * m->n_dof * m->n_dof);
    }
    delete dm;
    return 0;
}

// Cartpole: mode 1 = step, mode 2 = post_physics_step only
int emu_cartpole(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, const b2g_cartpole_cfg* cfg, int mode,
                 int n_envs, float* root, float* dof, float* dof_force, float* contact, const float* actions_in, float* obs,
                 float* obs_clamped, float* rew, long long* reset, long long* progress, long long* timeout, float* actions,
                 int* reset_count, const float* rand_override) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    SimArgs A;
    A.M = dm;
    pack_dev_params(*sp, nullptr, nullptr, A.P);
    A.P.stats = g_stats;
    A.n_envs = n_envs; A.root = root; A.dof = dof; A.target = nullptr; A.actuation = nullptr; A.dof_force = dof_force;
    A.contact = contact; A.friction = nullptr;
    TaskArgs T;
    memset(&T.cfg, 0, sizeof(T.cfg));
    memset(&T.hcfg, 0, sizeof(T.hcfg));
    T.ccfg = *cfg; T.seed = cfg->seed; T.actions_in = actions_in; T.obs = obs; T.obs_clamped = obs_clamped; T.rew = rew; T.reset = reset;
    T.progress = progress; T.timeout = timeout; T.commands = nullptr; T.actions = actions; T.reset_count = reset_count;
    T.rand_override = rand_override; T.post_only = (mode == 2);
    std::vector<float> scratch((size_t)MAXC * CF_COUNT), bf((size_t)m->n_bodies * 3);
    std::vector<float> links(link_store_floats(1, B2G_MAX_DOF)), anc((size_t)B2G_MAX_CHAINS * kAncFloats);
    EmuGroup g; g.lanes = 1; g.count = 0; g.sense = 0;
    emu_ctx.g = &g; emu_ctx.lane = 0; emu_ctx.local_sense = 0;
    for (int e = 0; e < n_envs; e++) {
        ScratchStrided sc{scratch.data(), 1}; sc.links = links.data(); sc.anc = anc.data();
        cartpole_step_thread(A, T, e, true, sc, bf.data());
    }
    delete dm;
    return 0;
}

// Houndarm fused step: mode 1 = step, 2 = post_physics_step only
int emu_houndarm(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, const b2g_houndarm_cfg* cfg, int mode,
                 int n_envs, float* root, float* dof, float* dof_force, float* contact, const float* actions_in, float* obs,
                 float* obs_clamped, float* rew, long long* reset, long long* progress, long long* timeout, float* commands, float* actions,
                 int* reset_count, const float* rand_override) {
    DevModel* dm = new DevModel;
    const char* why;
    if (pack_dev_model(*m, *dp, *dm, &why) != 0) { delete dm; return -1; }
    if (!m->fixed_base || m->n_chains != 1 || m->n_dof > B2G_MAX_FIXED_CHAIN_LEN) { delete dm; return -2; }
    SimArgs A;
    A.M = dm;
    pack_dev_params(*sp, nullptr, nullptr, A.P);
    A.P.stats = g_stats;
    A.n_envs = n_envs; A.root = root; A.dof = dof; A.target = nullptr; A.actuation = nullptr; A.dof_force = dof_force;
    A.contact = contact; A.friction = nullptr; A.env_scale = g_env_scale; A.link_scale = g_link_scale;
    TaskArgs T;
    memset(&T.cfg, 0, sizeof(T.cfg));
    memset(&T.ccfg, 0, sizeof(T.ccfg));
    T.hcfg = *cfg; T.seed = cfg->seed; T.actions_in = actions_in; T.obs = obs; T.obs_clamped = obs_clamped; T.rew = rew; T.reset = reset;
    T.progress = progress; T.timeout = timeout; T.commands = commands; T.actions = actions; T.reset_count = reset_count;
    T.rand_override = rand_override; T.post_only = (mode == 2);
    std::vector<float> scratch((size_t)MAXC * CF_COUNT), bf((size_t)m->n_bodies * 3);
    std::vector<float> links(link_store_floats(1, B2G_MAX_DOF)), anc((size_t)B2G_MAX_CHAINS * kAncFloats);
    EmuGroup g; g.lanes = 1; g.count = 0; g.sense = 0;
    emu_ctx.g = &g; emu_ctx.lane = 0; emu_ctx.local_sense = 0;
    for (int e = 0; e < n_envs; e++) {
        ScratchStrided sc{scratch.data(), 1}; sc.links = links.data(); sc.anc = anc.data();
        if (m->n_dof > 6) houndarm_step_thread<7>(A, T, e, true, sc, bf.data());
        else houndarm_step_thread<6>(A, T, e, true, sc, bf.data());
    }
    delete dm;
    return 0;
}

}  // extern "C"
