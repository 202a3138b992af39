// libb200gym.so -- kernels + C ABI (include/b200gym.h).
//
// One CUDA thread per (environment, chain); LANES consecutive threads of a warp form one environment.
// Kernels: simulate (gym.simulate), fused Anymal/Hound step (pre_physics + simulate + post_physics),
// reset_all, forward-dynamics probe, rigid-body-state refresh, indexed row copies.
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see __graft_entry__.build()).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200gym.h"
#include "b2g_host_pack.h"
#include "b2g_threads.cuh"

using namespace b2g;

namespace {

#ifndef B2G_BLOCK
#define B2G_BLOCK 64
#endif
constexpr int kBlock = B2G_BLOCK;
// experiment switch: B2G_SPARSE = 2 leaves the upper half of every warp idle (half as many environments per warp, twice as
// many warps): less divergence per warp, a second warp per scheduler at the 4096-env headline
#ifndef B2G_SPARSE
#define B2G_SPARSE 1
#endif
constexpr int kSparse = B2G_SPARSE;
template <int LANES> struct EnvsPerBlock { static constexpr int value = (kBlock / 32) * (32 / LANES / kSparse); };

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace

namespace b2g {
// shared with the other translation units of the library (b2g_policy.cu)
int fail_msg(int code, const char* msg) { return fail(code, "%s", msg); }
}  // namespace b2g

namespace {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return fail(B2G_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct Variant {
    int lanes, nl;
    bool fixed;
    bool seg = false;      // <8,3>: chains cut into pieces of at most three links (DevModel::seg_*)
};

Variant pick_variant(const b2g_model& m) {
    int maxlen = 0;
    for (int c = 0; c < m.n_chains; c++) maxlen = m.chain_len[c] > maxlen ? m.chain_len[c] : maxlen;
    int minlen = 1 << 20;
    for (int c = 0; c < m.n_chains; c++) minlen = m.chain_len[c] < minlen ? m.chain_len[c] : minlen;
    // the specialised variants assume FULL chains (every lane has exactly NL links)
    if (m.fixed_base && m.n_chains == 1 && maxlen == 2) return {1, 2, true};
    if (!m.fixed_base && m.n_chains == 4 && maxlen == 3 && minlen == 3) return {4, 3, false};
    // B2G_SEGMENTS=1: chains cut into pieces of at most three links that own a lane each (<8,3>, link state in registers) when the
    // pieces fit the eight lanes.  Parity-green and measured: 438 us per UsefulHound step against 397 us for the
    // whole-chain kernels -- the recursions stay serial over six links -- so it is an opt-in experiment, not the default.
    int pieces = m.n_chains;
    for (int c = 0; c < m.n_chains; c++) pieces += m.chain_len[c] > kSegLinks ? 1 : 0;
    const char* seg = getenv("B2G_SEGMENTS");
    if (seg && seg[0] == '1' && !m.fixed_base && m.n_chains > 0 && pieces <= B2G_MAX_CHAINS) return {8, 3, false, true};
    if (m.fixed_base) return {8, B2G_MAX_FIXED_CHAIN_LEN, true};      // one lane in use: the arm's single chain (<= 7 links)
    return {8, 6, false};
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <int LANES>
__device__ __forceinline__ void thread_ids(int n_envs, int nb, float* smem, int& env, int& lane, bool& valid, ScratchStrided& sc, float*& bf, int nd, int slots, int nstore = 0) {
    const int tid = threadIdx.x;
    constexpr int EPW = 32 / LANES / kSparse, EPB = EnvsPerBlock<LANES>::value;
    const int wl = tid & 31;
    const bool live = wl < EPW * LANES;
    int eib = (tid >> 5) * EPW + (live ? wl / LANES : 0);
    lane = tid % LANES;
    if (LANES == kSplit8) {      // Grp<8> layout: chains 0-3 in lanes 0-15 (environment-major), chains 4-7 in lanes 16-31 (chain-major)
        static_assert(LANES != 8 || kSparse == 1, "the 8-lane layout assumes full warps");
        eib = (tid >> 5) * EPW + (wl < 16 ? wl >> 2 : wl & 3);
        lane = wl < 16 ? (wl & 3) : 4 + ((wl - 16) >> 2);
    }
    const int e = blockIdx.x * EPB + eib;
    valid = live && e < n_envs;
    env = (e < n_envs) ? e : n_envs - 1;
    sc.base = smem + tid;
    sc.stride = kBlock;
    const int grp = LANES == kSplit8 ? eib : tid / LANES;
    bf = smem + kBlock * slots * CF_COUNT + grp * nb * 3;     // one accumulator per lane group, idle groups included
    // link store of the rolled long-chain variants (b2g_dynamics.cuh::links_in_shared): one LinkData per DOF per lane group
    sc.links = smem + ((kBlock * slots * CF_COUNT + (kBlock / LANES) * nb * 3 + 3) & ~3) + link_store_floats(grp, nd);     // 16-byte aligned
    // ancestor store of the segment variant (b2g_dynamics.cuh::kAncFloats per piece that has a child), same 16-byte aligned base
    sc.anc = smem + ((kBlock * slots * CF_COUNT + (kBlock / LANES) * nb * 3 + 3) & ~3) + (kLinksShared ? link_store_floats(kBlock / LANES, nd) : 0) + grp * nstore * kAncFloats;
}

// ---- host mirror (b2g_task_step_host): the step's outputs (obs_clamped | rew | reset | timeout, the b2g_task_host_layout arena)
// are stored straight into the caller's page-locked buffer by the SMs and a sequence number is published last, so the host
// neither queues a copy command nor waits on the stream: it polls one word of host memory.
struct HostMirror {
    unsigned char* dst = nullptr;         // device alias of the caller's pinned buffer (null = no mirror)
    const unsigned char* src = nullptr;   // the device arena
    unsigned long long off[4] = {0, 0, 0, 0};
    int num_obs = 0;
    unsigned* done_ctr = nullptr;         // device: blocks that have finished their stores
    unsigned* flag = nullptr;             // device alias of the pinned sequence word
    unsigned seq = 0;
};

// coalesced copy of `bytes` (multiple of 4) by the whole block; 16-byte units when both ends allow it.  Loads bypass L1
// (the data was written by other threads of this block, or by a previous kernel).
__device__ __forceinline__ void mirror_span(unsigned char* dst, const unsigned char* src, size_t bytes) {
    if ((((size_t)dst | (size_t)src | bytes) & 15) == 0) {
        for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(dst)[i] = __ldcg(reinterpret_cast<const uint4*>(src) + i);
    } else {
        for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x)
            reinterpret_cast<unsigned*>(dst)[i] = __ldcg(reinterpret_cast<const unsigned*>(src) + i);
    }
}

// every thread fences its own host stores; the block that arrives last publishes the sequence number
__device__ __forceinline__ void mirror_publish(const HostMirror& H) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(H.done_ctr, 1u);
        if (prev == gridDim.x - 1) {
            *H.done_ctr = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(H.flag) = H.seq;
        }
    }
}

// tail of a fused step kernel: this block's environments [e0, e0 + cnt) go to the host as soon as the block is done, so the
// PCIe transfer overlaps the blocks that are still computing
__device__ __forceinline__ void mirror_block(const HostMirror& H, int e0, int cnt) {
    __syncthreads();   // the block's own global stores are visible to all its threads
    if (cnt > 0) {
        const size_t ob = (size_t)H.num_obs * 4;
        mirror_span(H.dst + H.off[0] + e0 * ob, H.src + H.off[0] + e0 * ob, cnt * ob);
        mirror_span(H.dst + H.off[1] + (size_t)e0 * 4, H.src + H.off[1] + (size_t)e0 * 4, (size_t)cnt * 4);
        mirror_span(H.dst + H.off[2] + (size_t)e0 * 8, H.src + H.off[2] + (size_t)e0 * 8, (size_t)cnt * 8);
        mirror_span(H.dst + H.off[3] + (size_t)e0 * 8, H.src + H.off[3] + (size_t)e0 * 8, (size_t)cnt * 8);
    }
    mirror_publish(H);
}

// stand-alone mirror for the step kernels that do not carry the tail: whole arena, grid-stride
__global__ void __launch_bounds__(256) k_mirror_host(HostMirror H, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(H.dst)[i] = __ldcg(reinterpret_cast<const uint4*>(H.src) + i);
    mirror_publish(H);
}

template <int LANES, int NL, bool FIXED, bool HF>
__global__ void __launch_bounds__(kBlock) k_simulate(SimArgs A) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<LANES>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts, A.M->n_seg_store);
    simulate_thread<LANES, NL, FIXED, HF>(A, env, lane, valid, sc, bf);
}

// MINB = minimum resident blocks per SM the register allocation must allow: 1 = no cap (233 registers, 4 blocks/SM: lowest
// latency per warp, the 4096-env headline), 6 = 168 registers with a few spills (3 warps per sub-partition) for grids that
// exceed one wave of the uncapped variant
template <int LANES, int NL, bool HF, int MINB = 1>
__global__ void __launch_bounds__(kBlock, MINB) k_anymal_step(SimArgs A, TaskArgs T, HostMirror H) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<LANES>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts, A.M->n_seg_store);
    anymal_step_thread<LANES, NL, HF>(A, T, env, lane, valid, sc, bf);
    if (H.dst) {
        constexpr int EPB = EnvsPerBlock<LANES>::value;
        const int e0 = blockIdx.x * EPB, left = A.n_envs - e0;
        mirror_block(H, e0, left < EPB ? left : EPB);
    }
}

template <int NJ>
__global__ void __launch_bounds__(kBlock) k_houndarm_step(SimArgs A, TaskArgs T) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<1>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts);
    houndarm_step_thread<NJ>(A, T, env, valid, sc, bf);
}

__global__ void __launch_bounds__(kBlock) k_cartpole_step(SimArgs A, TaskArgs T) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<1>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts);
    cartpole_step_thread(A, T, env, valid, sc, bf);
}

// Grid-wide "who finishes last": every thread's global writes are fenced, one thread per block takes a ticket; returns true
// (block-uniformly) in the block that arrived last, with the counter already re-armed for the next launch.
__device__ __forceinline__ bool last_block_arrives(unsigned* ticket) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ticket = 0u;
    }
    __syncthreads();
    const bool r = is_last != 0;
    if (r) __threadfence();
    return r;
}

template <int LANES, int NL, bool HF>
__global__ void __launch_bounds__(kBlock) k_terrain_phys(SimArgs A, TerrainArgs T) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<LANES>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts, A.M->n_seg_store);
    terrain_phys_thread<LANES, NL, HF>(A, T, env, lane, valid, sc, bf);
    // curriculum scalar for ALL resetting envs (reference quirk: torch.norm without dim, anymal_terrain.py:432): the last
    // block to arrive sums the N values in a fixed order (kBlock strided partial sums, pairwise tree) -> deterministic
    if (T.cnorm && T.cfg.custom_origins && T.cfg.curriculum && T.init_done) {
        __shared__ float red[kBlock];
        if (last_block_arrives(T.tickets + 0)) {
            float acc = 0.0f;     // 16 loads in flight, added in index order (the + 0.0f of a missing tail element is exact)
            for (int i0 = threadIdx.x; i0 < A.n_envs; i0 += kBlock * 16) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; u++) v[u] = i0 + u * kBlock < A.n_envs ? __ldcg(T.resetw + i0 + u * kBlock) : 0.0f;
#pragma unroll
                for (int u = 0; u < 16; u++) acc += v[u];
            }
            red[threadIdx.x] = acc;
            __syncthreads();
            for (int o = kBlock / 2; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) *T.cnorm = sqrtf(red[0]);
        }
    }
}

// 16 threads per environment (kPostSub), 16 environments per block
#ifndef B2G_POST_SUB
#define B2G_POST_SUB 16
#endif
constexpr int kPostSub = B2G_POST_SUB, kPostBlock = 256, kPostEnvs = kPostBlock / kPostSub;
template <int LANES, int NL>
__global__ void __launch_bounds__(kPostBlock) k_terrain_post(SimArgs A, TerrainArgs T, HostMirror H) {
    static_assert(LANES <= kPostSub, "chain lanes are the first sub-threads of an environment");
    // one scalar for ALL resetting envs, reduced once by the last block of k_terrain_phys
    const float cnorm = (T.cfg.custom_origins && T.cfg.curriculum && T.init_done) ? *T.cnorm : 0.0f;
    const int tid = threadIdx.x;
    constexpr int EPB = kPostEnvs;
    const int e = blockIdx.x * EPB + tid / kPostSub;
    const bool valid = e < A.n_envs;
    __shared__ float rep[kPostEnvs][16];      // this block's report rows | reset flag | terrain level
    __shared__ float red[15][kPostBlock];
    terrain_post_thread<LANES, NL, kPostSub>(A, T, valid ? e : A.n_envs - 1, tid % kPostSub, valid, cnorm, valid ? rep[tid / kPostSub] : nullptr);
    if (H.dst) {      // b2g_task_step_host: this block's rows cross PCIe while the other blocks are still working
        const int m0 = blockIdx.x * EPB, left = A.n_envs - m0;
        mirror_block(H, m0, left < EPB ? left : EPB);
    }
    // extras["episode"] (anymal_terrain.py:420-425): means over the envs that reset this step.  Per-block sums in env order,
    // then the last block to arrive adds the blocks in a fixed order (strided partial sums, pairwise tree): deterministic, no
    // extra launch
    if (!T.extras) return;
    __syncthreads();
    if (tid < 15) {
        const int cnt = A.n_envs - blockIdx.x * EPB < EPB ? A.n_envs - blockIdx.x * EPB : EPB;
        float acc = 0.0f;
        for (int i = 0; i < cnt; i++) acc += rep[i][tid];
        T.extras_part[(size_t)blockIdx.x * 16 + tid] = acc;
    }
    if (last_block_arrives(T.tickets + 1)) {
        float acc[15];
#pragma unroll
        for (int k = 0; k < 15; k++) acc[k] = 0.0f;
        for (int b = tid; b < (int)gridDim.x; b += kPostBlock) {
            float v[15];
#pragma unroll
            for (int k = 0; k < 15; k++) v[k] = __ldcg(T.extras_part + (size_t)b * 16 + k);
#pragma unroll
            for (int k = 0; k < 15; k++) acc[k] += v[k];
        }
#pragma unroll
        for (int k = 0; k < 15; k++) red[k][tid] = acc[k];
        __syncthreads();
        for (int o = kPostBlock / 2; o > 0; o >>= 1) {
            if (tid < o) {
#pragma unroll
                for (int k = 0; k < 15; k++) red[k][tid] += red[k][tid + o];
            }
            __syncthreads();
        }
        const float cnt = red[13][0];
        if (tid < 13 && cnt > 0.0f) T.extras[tid] = red[tid][0] / cnt * (1.0f / T.cfg.max_episode_length_s);
        if (tid == 13 && cnt > 0.0f) { T.extras[13] = red[14][0] / (float)A.n_envs; T.extras[14] = cnt; }
        if (tid == 0 && T.step_ctr_advance) *T.step_ctr_advance += 1;     // end of the step: the next step sees the next counter value
    }
}

template <int LANES, int NL>
__global__ void __launch_bounds__(kBlock) k_anymal_reset_all(SimArgs A, TaskArgs T) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<LANES>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts, A.M->n_seg_store);
    anymal_reset_all_thread<LANES, NL>(A, T, env, lane, valid);
}

template <int LANES, int NL, bool FIXED>
__global__ void __launch_bounds__(kBlock) k_probe(SimArgs A, float* qdd, float* a0) {
    extern __shared__ __align__(16) float smem[];
    int env, lane; bool valid; ScratchStrided sc; float* bf;
    thread_ids<LANES>(A.n_envs, A.M->n_bodies, smem, env, lane, valid, sc, bf, A.M->n_dof, A.P.max_contacts, A.M->n_seg_store);
    const DevModel* M = A.M;
    int len, d0;
    lane_span<LANES, NL>(M, lane, len, d0);
    LaneState<NL> st;
    load_state<NL>(A, env, len, d0, st);
#pragma unroll
    for (int j = 0; j < NL; j++)
        if (j < len) st.act[j] = A.actuation[(size_t)env * M->n_dof + d0 + j];
    substep<LANES, NL, FIXED, false, true>(M, A.P, lane, len, d0, st, env_dr(A, valid ? env : 0, valid, false), false, sc, bf);
    if (valid) {
#pragma unroll
        for (int j = 0; j < NL; j++)
            if (j < len) qdd[(size_t)env * M->n_dof + d0 + j] = st.frc[j];
        if (lane == 0) {
            float* o = a0 + (size_t)env * 6;
            o[0] = st.rw.x; o[1] = st.rw.y; o[2] = st.rw.z; o[3] = st.rv.x; o[4] = st.rv.y; o[5] = st.rv.z;
        }
    }
}

__global__ void k_body_state(const DevModel* M, const float* root, const float* dof, float* out, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    body_state_env(M, root + (size_t)e * 13, dof + (size_t)e * M->n_dof * 2, out + (size_t)e * M->n_bodies * 13);
}

__global__ void k_mass_matrix(const DevModel* M, const float* root, const float* dof, float* out, int n, const float* env_scale,
                              const float* link_scale) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    mass_matrix_env(M, root + (size_t)e * 13, dof + (size_t)e * M->n_dof * 2, out + (size_t)e * M->n_dof * M->n_dof,
                    env_scale ? env_scale[(size_t)e * 4] : 1.0f, link_scale ? link_scale + (size_t)e * (M->n_dof + 1) * B2G_LINK_SCALE_COLS : nullptr);
}

__global__ void k_jacobian(const DevModel* M, const float* root, const float* dof, float* out, int n, int per_env) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    jacobian_env(M, root + (size_t)e * 13, dof + (size_t)e * M->n_dof * 2, out + (size_t)e * per_env);
}

// rows idx[0..n) of src -> dst, row = `row` floats (gym.set_*_tensor_indexed)
__global__ void k_copy_rows(float* dst, const float* src, const int* idx, int n, int row) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * row) return;
    const int r = idx[t / row], c = t % row;
    dst[(size_t)r * row + c] = src[(size_t)r * row + c];
}

__global__ void k_fill_rows(float* dst, const float* row_vals, int n, int row) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * row) return;
    dst[t] = row_vals[t % row];
}

__global__ void k_fill_i64(long long* dst, long long v, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = v;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// the sim object
// ------------------------------------------------------------------------------------------------
struct b2g_sim {
    int device = 0;
    int n_sm = 148;
    b2g_sim_params params{};
    b2g_model model{};
    b2g_dof_props props{};
    bool has_model = false, prepared = false;
    b2g_heightfield hf{};
    int16_t* d_hf = nullptr;
    bool has_hf = false;
    std::vector<int16_t> h_hf;     // host copy of the samples: the coarse bound is rebuilt when the model arrives later
    float* d_hfc = nullptr;        // coarse conservative bound (build_hf_coarse)
    int hfc_rows = 0, hfc_cols = 0;
    float link_rmax = 0.0f;
    int n_envs = 0;
    float root_pose[7] = {0, 0, 0, 0, 0, 0, 1};
    Variant v{4, 3, false};
    DevModel* d_model = nullptr;
    int n_seg_store = 0;      // pieces with a child (DevModel::n_seg_store): sizes the segment variant's shared-memory ancestor store
    float* t[B2G_T_COUNT] = {nullptr};
    // task
    bool has_task = false;
    int task_kind = 0;             // 1 = flat locomotion (Anymal/Hound), 2 = Cartpole, 3 = rough terrain, 4 = Houndarm
    int num_obs = 0, num_act = 0, n_draws = 0, n_cmd = 3;
    unsigned long long seed = 0;
    b2g_anymal_cfg acfg{};
    b2g_cartpole_cfg ccfg{};
    b2g_houndarm_cfg hcfg{};
    b2g_terrain_cfg tcfg{};
    float *torques = nullptr, *last_actions = nullptr, *last_dof_vel = nullptr, *feet_air_time = nullptr, *episode_sums = nullptr;
    float *env_origins = nullptr, *terrain_origins = nullptr, *scratch9 = nullptr, *resetw = nullptr, *report = nullptr, *measured = nullptr;
    float *noise_override = nullptr, *push_override = nullptr, *extras = nullptr;
    float *cnorm = nullptr, *extras_part = nullptr;   // cross-block reductions of the terrain step
    unsigned* tickets = nullptr;
    float *arm_mm = nullptr, *arm_jac = nullptr, *eef_state = nullptr, *arm_commands = nullptr;
    long long *terrain_levels = nullptr, *terrain_types = nullptr;
    int16_t* height_samples = nullptr;
    long long common_step = 0;
    long long* step_ctr = nullptr;   // device copy of the counter the NEXT step uses (b2g_task_terrain_device_step)
    int auto_step = 0;
    int env_scale_used = 0;          // set when B2G_T_ENV_SCALE is first acquired: until then the kernels skip the loads
    int link_scale_used = 0;         // same for B2G_T_LINK_SCALE
    int init_done = 0;
    float *obs = nullptr, *obs_clamped = nullptr, *rew = nullptr, *commands = nullptr, *actions = nullptr, *rand_override = nullptr;
    long long *reset = nullptr, *progress = nullptr, *timeout = nullptr;
    int* reset_count = nullptr;
    float* actions_in = nullptr;   // staging for step_host
    // obs_clamped | rew | reset | timeout live back to back in ONE allocation so that step_host can return them with a
    // single device-to-host copy when the caller's buffers follow the same layout (b2g_task_host_layout)
    unsigned char* out_arena = nullptr;
    size_t out_off[4] = {0, 0, 0, 0}, out_total = 0;
    unsigned* host_flag = nullptr;      // page-locked sequence word the mirror kernels publish (b2g_task_step_host)
    unsigned* host_flag_dev = nullptr;  // its device alias
    unsigned* done_ctr = nullptr;       // device: finished-block counter of the mirror
    unsigned host_seq = 0;
    int use_rand_override = 0;
    int64_t launches = 0;
    unsigned long long* d_stats = nullptr;   // contact statistics (DevParams::stats), 4 counters; B2G_CONTACT_STATS=0 switches them off
};

namespace {

size_t smem_bytes(const b2g_sim* s) {
    const int epb = kBlock / s->v.lanes;
    const size_t links = kLinksShared && s->v.nl > 3 ? 4 + link_store_floats(epb, s->model.n_dof) : 0;     // links_in_shared variants
    const size_t anc = s->v.seg ? 4 + (size_t)epb * s->n_seg_store * kAncFloats : 0;     // segment variant's ancestor store
    return sizeof(float) * ((size_t)kBlock * contact_slots(s->params) * CF_COUNT + (size_t)epb * s->model.n_bodies * 3 + links + anc);
}

int grid_size(const b2g_sim* s) {
    const int epb = (kBlock / 32) * (32 / s->v.lanes / kSparse);
    return (s->n_envs + epb - 1) / epb;
}

SimArgs make_args(const b2g_sim* s) {
    SimArgs A;
    A.M = s->d_model;
    pack_dev_params(s->params, s->has_hf ? &s->hf : nullptr, s->d_hf, A.P, s->d_hfc, s->hfc_rows, s->hfc_cols);
    A.P.stats = s->d_stats;
    A.n_envs = s->n_envs;
    A.root = s->t[B2G_T_ROOT_STATE];
    A.dof = s->t[B2G_T_DOF_STATE];
    A.target = s->t[B2G_T_DOF_TARGET];
    A.actuation = s->t[B2G_T_DOF_ACTUATION];
    A.dof_force = s->t[B2G_T_DOF_FORCE];
    A.contact = s->t[B2G_T_NET_CONTACT];
    A.friction = s->t[B2G_T_FRICTION];
    A.env_scale = s->env_scale_used ? s->t[B2G_T_ENV_SCALE] : nullptr;   // ones until somebody acquires the tensor
    A.link_scale = s->link_scale_used ? s->t[B2G_T_LINK_SCALE] : nullptr;
    return A;
}

TaskArgs make_task_args(const b2g_sim* s, const float* actions_in, int post_only = 0) {
    TaskArgs T;
    T.post_only = post_only;
    T.cfg = s->acfg;
    T.ccfg = s->ccfg;
    T.hcfg = s->hcfg;
    T.seed = s->seed;
    T.actions_in = actions_in;
    T.obs = s->obs; T.obs_clamped = s->obs_clamped; T.rew = s->rew; T.reset = s->reset; T.progress = s->progress;
    T.timeout = s->timeout; T.commands = s->commands; T.actions = s->actions; T.reset_count = s->reset_count;
    T.rand_override = s->use_rand_override ? s->rand_override : nullptr;
    return T;
}

// (Re)build the coarse heightfield bound; needs both the samples and the model's largest link radius.  B2G_NO_HFC=1 keeps
// the kernels on the exhaustive candidate tests (A/B timing; results are identical either way).
int rebuild_hfc(b2g_sim* s) {
    if (s->d_hfc) { cudaFree(s->d_hfc); s->d_hfc = nullptr; }
    s->hfc_rows = s->hfc_cols = 0;
    const char* off = getenv("B2G_NO_HFC");
    if (!s->has_hf || s->h_hf.empty() || (off && off[0] == '1')) return B2G_OK;
    std::vector<float> c;
    int cr = 0, cc = 0;
    build_hf_coarse(s->h_hf.data(), s->hf.rows, s->hf.cols, s->hf.horizontal_scale, s->hf.vertical_scale, s->link_rmax, c, cr, cc);
    cudaError_t e = cudaMalloc(&s->d_hfc, sizeof(float) * c.size());
    if (e == cudaSuccess) e = cudaMemcpy(s->d_hfc, c.data(), sizeof(float) * c.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(B2G_ERR_CUDA, "coarse heightfield upload: %s", cudaGetErrorString(e));
    s->hfc_rows = cr; s->hfc_cols = cc;
    return B2G_OK;
}

int upload_model(b2g_sim* s) {
    DevModel* h = (DevModel*)malloc(sizeof(DevModel));
    if (!h) return fail(B2G_ERR_ARG, "out of host memory");
    const char* why = "";
    if (pack_dev_model(s->model, s->props, *h, &why) != 0) {
        free(h);
        return fail(B2G_ERR_ARG, "bad model: %s", why);
    }
    cudaError_t e = cudaSuccess;
    if (!s->d_model) e = cudaMalloc(&s->d_model, sizeof(DevModel));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_model, h, sizeof(DevModel), cudaMemcpyHostToDevice);
    const float rmax = max_link_radius(*h);
    s->n_seg_store = h->n_seg_store;
    free(h);
    if (e != cudaSuccess) return fail(B2G_ERR_CUDA, "model upload: %s", cudaGetErrorString(e));
    if (rmax != s->link_rmax || (s->has_hf && !s->d_hfc)) {
        s->link_rmax = rmax;
        if (s->has_hf) return rebuild_hfc(s);
    }
    return B2G_OK;
}

size_t tensor_floats(const b2g_sim* s, int kind) {
    const size_t n = s->n_envs, nd = s->model.n_dof, nb = s->model.n_bodies;
    const size_t jcols = nd + (s->model.fixed_base ? 0 : 6);
    switch (kind) {
        case B2G_T_ROOT_STATE: return n * 13;
        case B2G_T_DOF_STATE: return n * nd * 2;
        case B2G_T_NET_CONTACT: return n * nb * 3;
        case B2G_T_DOF_FORCE: return n * nd;
        case B2G_T_RIGID_BODY_STATE: return n * nb * 13;
        case B2G_T_DOF_TARGET: return n * nd;
        case B2G_T_DOF_ACTUATION: return n * nd;
        case B2G_T_JACOBIAN: return n * (nb - (s->model.fixed_base ? 1 : 0)) * 6 * jcols;
        case B2G_T_MASS_MATRIX: return n * nd * nd;
        case B2G_T_FRICTION: return n;
        case B2G_T_ENV_SCALE: return n * 4;
        case B2G_T_LINK_SCALE: return n * (nd + 1) * B2G_LINK_SCALE_COLS;
        default: return 0;
    }
}

void describe(const b2g_sim* s, int kind, b2g_tensor_desc* d) {
    const int64_t n = s->n_envs, nd = s->model.n_dof, nb = s->model.n_bodies;
    const int64_t jcols = nd + (s->model.fixed_base ? 0 : 6);
    d->dtype = 0;
    d->device_id = s->device;
    d->shape[0] = d->shape[1] = d->shape[2] = d->shape[3] = 1;
    switch (kind) {
        case B2G_T_ROOT_STATE: d->ndim = 2; d->shape[0] = n; d->shape[1] = 13; break;
        case B2G_T_DOF_STATE: d->ndim = 2; d->shape[0] = n * nd; d->shape[1] = 2; break;
        case B2G_T_NET_CONTACT: d->ndim = 2; d->shape[0] = n * nb; d->shape[1] = 3; break;
        case B2G_T_DOF_FORCE: d->ndim = 1; d->shape[0] = n * nd; break;
        case B2G_T_RIGID_BODY_STATE: d->ndim = 2; d->shape[0] = n * nb; d->shape[1] = 13; break;
        case B2G_T_DOF_TARGET: d->ndim = 1; d->shape[0] = n * nd; break;
        case B2G_T_DOF_ACTUATION: d->ndim = 1; d->shape[0] = n * nd; break;
        case B2G_T_JACOBIAN: d->ndim = 4; d->shape[0] = n; d->shape[1] = nb - (s->model.fixed_base ? 1 : 0); d->shape[2] = 6; d->shape[3] = jcols; break;
        case B2G_T_MASS_MATRIX: d->ndim = 3; d->shape[0] = n; d->shape[1] = nd; d->shape[2] = nd; break;
        case B2G_T_FRICTION: d->ndim = 1; d->shape[0] = n; break;
        case B2G_T_ENV_SCALE: d->ndim = 2; d->shape[0] = n; d->shape[1] = 4; break;
        case B2G_T_LINK_SCALE: d->ndim = 3; d->shape[0] = n; d->shape[1] = nd + 1; d->shape[2] = B2G_LINK_SCALE_COLS; break;
        default: d->ndim = 0;
    }
}

template <class F>
int with_device(b2g_sim* s, F&& f) {
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != s->device) cudaSetDevice(s->device);
    const int rc = f();
    if (prev != s->device) cudaSetDevice(prev);
    return rc;
}

int launch_simulate(b2g_sim* s, cudaStream_t st) {
    const SimArgs A = make_args(s);
    const int grid = grid_size(s);
    const size_t sm = smem_bytes(s);
    const bool hf = s->has_hf;
    if (s->v.lanes == 1) k_simulate<1, 2, true, false><<<grid, kBlock, sm, st>>>(A);
    else if (s->v.lanes == 4) { if (hf) k_simulate<4, 3, false, true><<<grid, kBlock, sm, st>>>(A); else k_simulate<4, 3, false, false><<<grid, kBlock, sm, st>>>(A); }
    else if (s->v.fixed) k_simulate<8, 7, true, false><<<grid, kBlock, sm, st>>>(A);
    else if (s->v.seg) { if (hf) k_simulate<8, 3, false, true><<<grid, kBlock, sm, st>>>(A); else k_simulate<8, 3, false, false><<<grid, kBlock, sm, st>>>(A); }
    else { if (hf) k_simulate<8, 6, false, true><<<grid, kBlock, sm, st>>>(A); else k_simulate<8, 6, false, false><<<grid, kBlock, sm, st>>>(A); }
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2G_OK;
}

int launch_anymal_step(b2g_sim* s, const float* actions_dev, cudaStream_t st, int post_only = 0, const HostMirror* hm = nullptr) {
    const SimArgs A = make_args(s);
    const TaskArgs T = make_task_args(s, actions_dev, post_only);
    const int grid = grid_size(s);
    const size_t sm = smem_bytes(s);
    if (s->task_kind == 3) {
        TerrainArgs R;
        R.cfg = s->tcfg; R.actions_in = actions_dev; R.obs = s->obs; R.obs_clamped = s->obs_clamped; R.rew = s->rew; R.reset = s->reset;
        R.progress = s->progress; R.timeout = s->timeout; R.commands = s->commands; R.actions = s->actions; R.torques = s->torques;
        R.last_actions = s->last_actions; R.last_dof_vel = s->last_dof_vel; R.feet_air_time = s->feet_air_time; R.episode_sums = s->episode_sums;
        R.env_origins = s->env_origins; R.terrain_levels = s->terrain_levels; R.terrain_types = s->terrain_types; R.terrain_origins = s->terrain_origins;
        R.height_samples = s->height_samples; R.scratch = s->scratch9; R.resetw = s->resetw; R.report = s->report; R.measured = s->measured;
        R.reset_count = s->reset_count;
        R.arm_mm = s->arm_mm; R.arm_jac = s->arm_jac; R.eef_state = s->eef_state; R.arm_commands = s->arm_commands;
        R.reset_override = s->use_rand_override ? s->rand_override : nullptr;
        R.noise_override = s->use_rand_override ? s->noise_override : nullptr;
        R.push_override = s->use_rand_override ? s->push_override : nullptr;
        R.common_step = s->common_step; R.step_ctr = s->auto_step ? s->step_ctr : nullptr; R.init_done = s->init_done; R.post_only = post_only; R.seed = s->seed;
        const bool advance = s->auto_step && post_only == 0;
        R.cnorm = s->cnorm; R.extras_part = s->extras_part; R.extras = s->extras; R.tickets = s->tickets;
        R.step_ctr_advance = advance ? s->step_ctr : nullptr;
        const HostMirror H3 = (hm && post_only == 0) ? *hm : HostMirror{};     // the mirror is the tail of k_terrain_post
        if (s->v.lanes == 4) {
            if (s->has_hf) k_terrain_phys<4, 3, true><<<grid, kBlock, sm, st>>>(A, R); else k_terrain_phys<4, 3, false><<<grid, kBlock, sm, st>>>(A, R);
            if (post_only != 2) k_terrain_post<4, 3><<<(s->n_envs + kPostEnvs - 1) / kPostEnvs, kPostBlock, 0, st>>>(A, R, H3);
        } else if (s->v.seg) {
            if (s->has_hf) k_terrain_phys<8, 3, true><<<grid, kBlock, sm, st>>>(A, R); else k_terrain_phys<8, 3, false><<<grid, kBlock, sm, st>>>(A, R);
            if (post_only != 2) k_terrain_post<8, 3><<<(s->n_envs + kPostEnvs - 1) / kPostEnvs, kPostBlock, 0, st>>>(A, R, H3);
        } else {
            if (s->has_hf) k_terrain_phys<8, 6, true><<<grid, kBlock, sm, st>>>(A, R); else k_terrain_phys<8, 6, false><<<grid, kBlock, sm, st>>>(A, R);
            if (post_only != 2) k_terrain_post<8, 6><<<(s->n_envs + kPostEnvs - 1) / kPostEnvs, kPostBlock, 0, st>>>(A, R, H3);
        }
        if (post_only == 2) {
            s->launches += 1;
            CUDA_TRY(cudaGetLastError());
            return B2G_OK;
        }
        if (advance) s->common_step++;
        s->launches += 2;
        CUDA_TRY(cudaGetLastError());
        return B2G_OK;
    }
    if (s->task_kind == 4) {     // one thread per environment whatever the generic kernels' lane count is
        const int g1 = (s->n_envs + kBlock / kSparse - 1) / (kBlock / kSparse);
        const size_t sm1 = sizeof(float) * ((size_t)kBlock * contact_slots(s->params) * CF_COUNT + (size_t)kBlock * s->model.n_bodies * 3 + (kLinksShared ? 4 + link_store_floats(kBlock, s->model.n_dof) : 0));
        if (s->model.n_dof > 6) k_houndarm_step<7><<<g1, kBlock, sm1, st>>>(A, T);      // Manipulator (7-DOF Franka)
        else k_houndarm_step<6><<<g1, kBlock, sm1, st>>>(A, T);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return B2G_OK;
    }
    if (s->task_kind == 2) {
        k_cartpole_step<<<grid, kBlock, sm, st>>>(A, T);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return B2G_OK;
    }
    const HostMirror H = hm ? *hm : HostMirror{};     // flat tasks: the mirror is the kernel's own tail
    if (s->v.lanes == 4) {
        if (s->has_hf) k_anymal_step<4, 3, true><<<grid, kBlock, sm, st>>>(A, T, H);
        else if (grid > 4 * s->n_sm && sm * 6 <= 200 * 1024) k_anymal_step<4, 3, false, 6><<<grid, kBlock, sm, st>>>(A, T, H);     // more than one wave and six blocks' scratch fit an SM: occupancy build
        else k_anymal_step<4, 3, false><<<grid, kBlock, sm, st>>>(A, T, H);
    }
    else if (s->v.seg) { if (s->has_hf) k_anymal_step<8, 3, true><<<grid, kBlock, sm, st>>>(A, T, H); else k_anymal_step<8, 3, false><<<grid, kBlock, sm, st>>>(A, T, H); }
    else { if (s->has_hf) k_anymal_step<8, 6, true><<<grid, kBlock, sm, st>>>(A, T, H); else k_anymal_step<8, 6, false><<<grid, kBlock, sm, st>>>(A, T, H); }
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2G_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int b2g_abi_version(void) { return B2G_ABI_VERSION; }
const char* b2g_last_error(void) { return g_err; }

// sizeof() of the public PODs, so the Python mirror can verify its layout
int b2g_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(b2g_model);
        case 1: return (int)sizeof(b2g_sim_params);
        case 2: return (int)sizeof(b2g_dof_props);
        case 3: return (int)sizeof(b2g_heightfield);
        case 4: return (int)sizeof(b2g_tensor_desc);
        case 5: return (int)sizeof(b2g_anymal_cfg);
        case 6: return (int)sizeof(b2g_cartpole_cfg);
        case 7: return (int)sizeof(b2g_terrain_cfg);
        case 8: return (int)sizeof(b2g_houndarm_cfg);
        default: return -1;
    }
}

int b2g_sim_create(int device_id, const b2g_sim_params* params, b2g_sim** out) {
    if (!params || !out) return fail(B2G_ERR_ARG, "null argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(B2G_ERR_CUDA, "no usable CUDA device (%s); libb200gym has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device_id < 0 || device_id >= count) return fail(B2G_ERR_ARG, "device %d out of range (%d devices)", device_id, count);
    b2g_sim* s = new b2g_sim();
    s->device = device_id;
    cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, device_id);
    s->params = *params;
    if (s->params.substeps <= 0) s->params.substeps = 1;
    const char* cs = getenv("B2G_CONTACT_STATS");
    if (!(cs && cs[0] == '0')) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(device_id);
        if (cudaMalloc(&s->d_stats, 4 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(s->d_stats, 0, 4 * sizeof(unsigned long long));
        else s->d_stats = nullptr;
        cudaSetDevice(prev);
    }
    *out = s;
    return B2G_OK;
}

int b2g_sim_destroy(b2g_sim* s) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    with_device(s, [&]() {
        for (int k = 0; k < B2G_T_COUNT; k++) if (s->t[k]) cudaFree(s->t[k]);
        void* ptrs[] = {s->d_model, s->d_hf, s->d_hfc, s->obs, s->out_arena, s->commands, s->actions, s->rand_override,
                        s->progress, s->reset_count, s->actions_in, s->torques, s->last_actions, s->last_dof_vel,
                        s->feet_air_time, s->episode_sums, s->env_origins, s->terrain_origins, s->scratch9, s->resetw, s->report, s->measured,
                        s->noise_override, s->push_override, s->extras, s->terrain_levels, s->terrain_types, s->height_samples,
                        s->arm_mm, s->arm_jac, s->eef_state, s->arm_commands, s->step_ctr, s->done_ctr, s->cnorm, s->extras_part, s->tickets, s->d_stats};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (s->host_flag) cudaFreeHost(s->host_flag);
        return 0;
    });
    delete s;
    return B2G_OK;
}

int b2g_sim_set_params(b2g_sim* s, const b2g_sim_params* p) {
    if (!s || !p) return fail(B2G_ERR_ARG, "null argument");
    const int ground = s->params.has_ground;
    const int slots = s->params.max_contacts_per_chain;
    if (s->prepared && contact_slots(*p) != contact_slots(s->params))
        return fail(B2G_ERR_STATE, "max_contacts_per_chain cannot change after b2g_sim_prepare (shared memory is sized there)");
    s->params = *p;
    if (s->params.substeps <= 0) s->params.substeps = 1;
    s->params.has_ground = ground || p->has_ground;
    if (s->prepared) s->params.max_contacts_per_chain = slots;
    return B2G_OK;
}

int b2g_sim_get_params(const b2g_sim* s, b2g_sim_params* out) {
    if (!s || !out) return fail(B2G_ERR_ARG, "null argument");
    *out = s->params;
    return B2G_OK;
}

int b2g_sim_add_ground(b2g_sim* s, float sf, float df, float rest) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    s->params.has_ground = 1;
    s->params.plane_static_friction = sf;
    s->params.plane_dynamic_friction = df;
    s->params.plane_restitution = rest;
    return B2G_OK;
}

int b2g_sim_add_heightfield(b2g_sim* s, const b2g_heightfield* hf, const int16_t* samples) {
    if (!s || !hf || !samples) return fail(B2G_ERR_ARG, "null argument");
    if (hf->rows < 2 || hf->cols < 2 || hf->horizontal_scale <= 0) return fail(B2G_ERR_ARG, "bad heightfield dimensions");
    return with_device(s, [&]() {
        if (s->d_hf) { cudaFree(s->d_hf); s->d_hf = nullptr; }
        const size_t bytes = sizeof(int16_t) * (size_t)hf->rows * hf->cols;
        CUDA_TRY(cudaMalloc(&s->d_hf, bytes));
        CUDA_TRY(cudaMemcpy(s->d_hf, samples, bytes, cudaMemcpyHostToDevice));
        s->hf = *hf;
        s->has_hf = true;
        s->h_hf.assign(samples, samples + (size_t)hf->rows * hf->cols);
        if (s->has_model) return rebuild_hfc(s);
        return (int)B2G_OK;
    });
}

int b2g_sim_add_articulation(b2g_sim* s, const b2g_model* model, const b2g_dof_props* props, int n_envs, const float* root_pose7,
                             float env_spacing, int num_per_row) {
    (void)env_spacing; (void)num_per_row;   // state tensors are env-local (Isaac Gym convention); the grid only matters for rendering
    if (!s || !model || !props) return fail(B2G_ERR_ARG, "null argument");
    if (s->has_model) return fail(B2G_ERR_UNSUPPORTED, "one articulation type per sim");
    if (n_envs <= 0) return fail(B2G_ERR_ARG, "n_envs must be positive");
    s->model = *model;
    s->props = *props;
    s->n_envs = n_envs;
    if (root_pose7) memcpy(s->root_pose, root_pose7, sizeof(float) * 7);
    s->v = pick_variant(*model);
    const int rc = with_device(s, [&]() { return upload_model(s); });
    if (rc != B2G_OK) return rc;
    s->has_model = true;
    return B2G_OK;
}

int b2g_sim_set_dof_props(b2g_sim* s, const b2g_dof_props* props) {
    if (!s || !props) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_model) return fail(B2G_ERR_STATE, "no articulation yet");
    s->props = *props;
    return with_device(s, [&]() { return upload_model(s); });
}

int b2g_sim_prepare(b2g_sim* s) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    if (!s->has_model) return fail(B2G_ERR_STATE, "prepare_sim before any actor was created");
    if (s->prepared) return B2G_OK;
    return with_device(s, [&]() {
        for (int k = 0; k < B2G_T_COUNT; k++) {
            if (k == B2G_T_JACOBIAN || k == B2G_T_MASS_MATRIX) continue;   // allocated on first acquire
            const size_t nf = tensor_floats(s, k);
            CUDA_TRY(cudaMalloc(&s->t[k], sizeof(float) * (nf ? nf : 1)));
            CUDA_TRY(cudaMemset(s->t[k], 0, sizeof(float) * (nf ? nf : 1)));
        }
        // initial root state = the create_actor pose, zero velocity; friction 1
        float row[13] = {0};
        memcpy(row, s->root_pose, sizeof(float) * 7);
        float* d_row = nullptr;
        CUDA_TRY(cudaMalloc(&d_row, sizeof(row)));
        CUDA_TRY(cudaMemcpy(d_row, row, sizeof(row), cudaMemcpyHostToDevice));
        k_fill_rows<<<(s->n_envs * 13 + 255) / 256, 256>>>(s->t[B2G_T_ROOT_STATE], d_row, s->n_envs, 13);
        const float one = 1.0f;
        CUDA_TRY(cudaMemcpy(d_row, &one, sizeof(float), cudaMemcpyHostToDevice));
        k_fill_rows<<<(s->n_envs + 255) / 256, 256>>>(s->t[B2G_T_FRICTION], d_row, s->n_envs, 1);
        k_fill_rows<<<(s->n_envs * 4 + 255) / 256, 256>>>(s->t[B2G_T_ENV_SCALE], d_row, s->n_envs * 4, 1);
        {
            const float row[B2G_LINK_SCALE_COLS] = {1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f};
            float* d_row6 = nullptr;
            CUDA_TRY(cudaMalloc(&d_row6, sizeof(row)));
            CUDA_TRY(cudaMemcpy(d_row6, row, sizeof(row), cudaMemcpyHostToDevice));
            const int nls = s->n_envs * (s->model.n_dof + 1);
            k_fill_rows<<<(nls * B2G_LINK_SCALE_COLS + 255) / 256, 256>>>(s->t[B2G_T_LINK_SCALE], d_row6, nls, B2G_LINK_SCALE_COLS);
            CUDA_TRY(cudaDeviceSynchronize());
            cudaFree(d_row6);
        }
        s->launches += 3;
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(d_row);
        // opt in to the dynamic shared memory the kernels need
        const int sm = (int)smem_bytes(s);
        cudaFuncSetAttribute(k_simulate<1, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<4, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<4, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<8, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<8, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<8, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<8, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_terrain_phys<8, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_terrain_phys<8, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_probe<8, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<8, 7, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<8, 6, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_simulate<8, 6, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<4, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<4, 3, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<4, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<8, 6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_anymal_step<8, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_terrain_phys<4, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_terrain_phys<4, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_terrain_phys<8, 6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_terrain_phys<8, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_cartpole_step, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_probe<1, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_probe<4, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_probe<8, 7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        cudaFuncSetAttribute(k_probe<8, 6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        s->prepared = true;
        return (int)B2G_OK;
    });
}

int b2g_sim_tensor(b2g_sim* s, int kind, b2g_tensor_desc* out) {
    if (!s || !out) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "acquire_*_tensor before prepare_sim");
    if (kind < 0 || kind >= B2G_T_COUNT) return fail(B2G_ERR_ARG, "unknown tensor kind %d", kind);
    if (!s->t[kind]) {
        const int rc = with_device(s, [&]() {
            const size_t nf = tensor_floats(s, kind);
            CUDA_TRY(cudaMalloc(&s->t[kind], sizeof(float) * (nf ? nf : 1)));
            CUDA_TRY(cudaMemset(s->t[kind], 0, sizeof(float) * (nf ? nf : 1)));
            return (int)B2G_OK;
        });
        if (rc != B2G_OK) return rc;
    }
    if (kind == B2G_T_ENV_SCALE) s->env_scale_used = 1;
    if (kind == B2G_T_LINK_SCALE) s->link_scale_used = 1;
    describe(s, kind, out);
    out->data = s->t[kind];
    return B2G_OK;
}

int b2g_sim_simulate(b2g_sim* s, void* stream) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    if (!s->prepared) return fail(B2G_ERR_STATE, "simulate before prepare_sim");
    return with_device(s, [&]() { return launch_simulate(s, (cudaStream_t)stream); });
}

int b2g_sim_refresh(b2g_sim* s, int kind, void* stream) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    if (!s->prepared) return fail(B2G_ERR_STATE, "refresh before prepare_sim");
    if (kind == B2G_T_RIGID_BODY_STATE) {
        return with_device(s, [&]() {
            k_body_state<<<(s->n_envs + 63) / 64, 64, 0, (cudaStream_t)stream>>>(s->d_model, s->t[B2G_T_ROOT_STATE], s->t[B2G_T_DOF_STATE],
                                                                                s->t[B2G_T_RIGID_BODY_STATE], s->n_envs);
            s->launches++;
            CUDA_TRY(cudaGetLastError());
            return (int)B2G_OK;
        });
    }
    if (kind == B2G_T_JACOBIAN || kind == B2G_T_MASS_MATRIX) {
        if (!s->t[kind]) return B2G_OK;      // never acquired: nothing to refresh
        return with_device(s, [&]() {
            cudaStream_t st = (cudaStream_t)stream;
            if (kind == B2G_T_MASS_MATRIX)
                k_mass_matrix<<<(s->n_envs + 63) / 64, 64, 0, st>>>(s->d_model, s->t[B2G_T_ROOT_STATE], s->t[B2G_T_DOF_STATE], s->t[kind], s->n_envs,
                                                                            s->env_scale_used ? s->t[B2G_T_ENV_SCALE] : nullptr,
                                                                            s->link_scale_used ? s->t[B2G_T_LINK_SCALE] : nullptr);
            else
                k_jacobian<<<(s->n_envs + 63) / 64, 64, 0, st>>>(s->d_model, s->t[B2G_T_ROOT_STATE], s->t[B2G_T_DOF_STATE], s->t[kind], s->n_envs,
                                                                 (int)(tensor_floats(s, kind) / s->n_envs));
            s->launches++;
            CUDA_TRY(cudaGetLastError());
            return (int)B2G_OK;
        });
    }
    return B2G_OK;   // live state: nothing to do
}

int b2g_sim_set_indexed(b2g_sim* s, int kind, const void* src, const int32_t* idx, int n, void* stream) {
    if (!s || !src || (!idx && n > 0)) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "set_*_indexed before prepare_sim");
    if (kind < 0 || kind >= B2G_T_COUNT || !s->t[kind]) return fail(B2G_ERR_ARG, "unknown tensor kind %d", kind);
    if (n <= 0 || src == s->t[kind]) return B2G_OK;   // the task edited the live tensor in place
    int row = 0;
    const int nd = s->model.n_dof;
    switch (kind) {
        case B2G_T_ROOT_STATE: row = 13; break;
        case B2G_T_DOF_STATE: row = nd * 2; break;
        case B2G_T_DOF_TARGET: case B2G_T_DOF_ACTUATION: row = nd; break;
        default: return fail(B2G_ERR_UNSUPPORTED, "indexed set not supported for tensor kind %d", kind);
    }
    return with_device(s, [&]() {
        k_copy_rows<<<(n * row + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s->t[kind], (const float*)src, idx, n, row);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return (int)B2G_OK;
    });
}

int b2g_sim_set_tensor(b2g_sim* s, int kind, const void* src, void* stream) {
    if (!s || !src) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "set_*_tensor before prepare_sim");
    if (kind < 0 || kind >= B2G_T_COUNT || !s->t[kind]) return fail(B2G_ERR_ARG, "unknown tensor kind %d", kind);
    if (src == s->t[kind]) return B2G_OK;
    return with_device(s, [&]() {
        CUDA_TRY(cudaMemcpyAsync(s->t[kind], src, sizeof(float) * tensor_floats(s, kind), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return (int)B2G_OK;
    });
}

int b2g_sim_forward_dynamics(b2g_sim* s, float* qdd, float* a0, void* stream) {
    if (!s || !qdd || !a0) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "forward_dynamics before prepare_sim");
    return with_device(s, [&]() {
        const SimArgs A = make_args(s);
        const int grid = grid_size(s);
        const size_t sm = smem_bytes(s);
        cudaStream_t st = (cudaStream_t)stream;
        if (s->v.lanes == 1) k_probe<1, 2, true><<<grid, kBlock, sm, st>>>(A, qdd, a0);
        else if (s->v.lanes == 4) k_probe<4, 3, false><<<grid, kBlock, sm, st>>>(A, qdd, a0);
        else if (s->v.fixed) k_probe<8, 7, true><<<grid, kBlock, sm, st>>>(A, qdd, a0);
        else if (s->v.seg) k_probe<8, 3, false><<<grid, kBlock, sm, st>>>(A, qdd, a0);
        else k_probe<8, 6, false><<<grid, kBlock, sm, st>>>(A, qdd, a0);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return (int)B2G_OK;
    });
}

// ---- fused flat task ----
static int alloc_task_buffers(b2g_sim* s, int num_obs, int num_act, int n_draws) {
    return with_device(s, [&]() {
        const size_t n = s->n_envs;
        s->num_obs = num_obs; s->num_act = num_act; s->n_draws = n_draws;
        CUDA_TRY(cudaMalloc(&s->obs, sizeof(float) * n * num_obs));
        auto up = [](size_t b) { return (b + 255) / 256 * 256; };
        s->out_off[0] = 0;
        s->out_off[1] = s->out_off[0] + up(sizeof(float) * n * num_obs);
        s->out_off[2] = s->out_off[1] + up(sizeof(float) * n);
        s->out_off[3] = s->out_off[2] + up(sizeof(long long) * n);
        s->out_total = s->out_off[3] + up(sizeof(long long) * n);
        CUDA_TRY(cudaMalloc(&s->out_arena, s->out_total));
        CUDA_TRY(cudaMemset(s->out_arena, 0, s->out_total));
        s->obs_clamped = reinterpret_cast<float*>(s->out_arena + s->out_off[0]);
        s->rew = reinterpret_cast<float*>(s->out_arena + s->out_off[1]);
        s->reset = reinterpret_cast<long long*>(s->out_arena + s->out_off[2]);
        s->timeout = reinterpret_cast<long long*>(s->out_arena + s->out_off[3]);
        CUDA_TRY(cudaMalloc(&s->commands, sizeof(float) * n * 4));
        CUDA_TRY(cudaMalloc(&s->actions, sizeof(float) * n * num_act));
        CUDA_TRY(cudaMalloc(&s->actions_in, sizeof(float) * n * num_act));
        CUDA_TRY(cudaMalloc(&s->done_ctr, sizeof(unsigned)));
        CUDA_TRY(cudaMemset(s->done_ctr, 0, sizeof(unsigned)));
        CUDA_TRY(cudaHostAlloc(&s->host_flag, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        *s->host_flag = 0;
        CUDA_TRY(cudaHostGetDevicePointer(&s->host_flag_dev, s->host_flag, 0));
        CUDA_TRY(cudaMalloc(&s->rand_override, sizeof(float) * n * n_draws));
        CUDA_TRY(cudaMalloc(&s->progress, sizeof(long long) * n));
        CUDA_TRY(cudaMalloc(&s->reset_count, sizeof(int) * n));
        CUDA_TRY(cudaMemset(s->obs, 0, sizeof(float) * n * num_obs));
        CUDA_TRY(cudaMemset(s->obs_clamped, 0, sizeof(float) * n * num_obs));
        CUDA_TRY(cudaMemset(s->rew, 0, sizeof(float) * n));
        CUDA_TRY(cudaMemset(s->commands, 0, sizeof(float) * n * 4));
        CUDA_TRY(cudaMemset(s->actions, 0, sizeof(float) * n * num_act));
        CUDA_TRY(cudaMemset(s->rand_override, 0, sizeof(float) * n * n_draws));
        CUDA_TRY(cudaMemset(s->progress, 0, sizeof(long long) * n));
        CUDA_TRY(cudaMemset(s->timeout, 0, sizeof(long long) * n));
        CUDA_TRY(cudaMemset(s->reset_count, 0, sizeof(int) * n));
        k_fill_i64<<<((int)n + 255) / 256, 256>>>(s->reset, 1, (int)n);   // reset_buf starts at ones (vec_task.py:316)
        s->launches++;
        CUDA_TRY(cudaDeviceSynchronize());
        s->has_task = true;
        return (int)B2G_OK;
    });
}

int b2g_task_anymal_create(b2g_sim* s, const b2g_anymal_cfg* cfg) {
    if (!s || !cfg) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "task created before prepare_sim");
    if (s->model.fixed_base) return fail(B2G_ERR_UNSUPPORTED, "the flat locomotion task needs a floating base");
    if (s->params.self_collision && s->v.lanes == 4)
        return fail(B2G_ERR_UNSUPPORTED, "the fused flat-terrain step of the quadrupeds is built without self-collision (the reference's flat tasks create their actors "
                                         "with collision filter 1 = off); clear b2g_sim_params.self_collision or use the generic path");
    if (cfg->base_body < 0 || cfg->base_body >= s->model.n_bodies || cfg->n_knee < 0 || cfg->n_knee > 8)
        return fail(B2G_ERR_ARG, "bad base/knee body indices");
    for (int k = 0; k < cfg->n_knee; k++)
        if (cfg->knee_bodies[k] < 0 || cfg->knee_bodies[k] >= s->model.n_bodies) return fail(B2G_ERR_ARG, "bad knee body index");
    if (s->has_task && s->task_kind != 1) return fail(B2G_ERR_STATE, "another task already lives on this sim");
    s->acfg = *cfg;
    s->seed = cfg->seed;
    s->task_kind = 1;
    s->n_cmd = 3;
    if (s->has_task) return B2G_OK;
    const int nd = s->model.n_dof;
    return alloc_task_buffers(s, 12 + 3 * nd, nd, 2 * nd + 3);
}

int b2g_task_cartpole_create(b2g_sim* s, const b2g_cartpole_cfg* cfg) {
    if (!s || !cfg) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "task created before prepare_sim");
    if (!s->model.fixed_base || s->model.n_dof != 2 || s->v.lanes != 1) return fail(B2G_ERR_UNSUPPORTED, "Cartpole needs the fixed-base 2-DOF cart-pole model");
    if (s->has_task && s->task_kind != 2) return fail(B2G_ERR_STATE, "another task already lives on this sim");
    s->ccfg = *cfg;
    s->seed = cfg->seed;
    s->task_kind = 2;
    if (s->has_task) return B2G_OK;
    return alloc_task_buffers(s, 4, 1, 4);
}

int b2g_task_houndarm_create(b2g_sim* s, const b2g_houndarm_cfg* cfg) {
    if (!s || !cfg) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "task created before prepare_sim");
    if (!s->model.fixed_base || s->model.n_chains != 1 || s->model.n_dof < 1 || s->model.n_dof > B2G_MAX_FIXED_CHAIN_LEN)
        return fail(B2G_ERR_UNSUPPORTED, "Houndarm / Manipulator need a fixed-base single chain of at most 7 DOF");
    if (cfg->n_reset_tail < 0 || cfg->n_reset_tail > s->model.n_dof) return fail(B2G_ERR_ARG, "n_reset_tail out of range");
    if (cfg->eef_body < 0 || cfg->eef_body >= s->model.n_bodies || cfg->jac_body < 0 || cfg->jac_body >= s->model.n_bodies)
        return fail(B2G_ERR_ARG, "eef_body / jac_body out of range");
    if (!(cfg->action_scale != 0.0f)) return fail(B2G_ERR_ARG, "action_scale must be non-zero");
    if (s->has_task && s->task_kind != 4) return fail(B2G_ERR_STATE, "another task already lives on this sim");
    s->hcfg = *cfg;
    s->seed = cfg->seed;
    s->task_kind = 4;
    if (s->has_task) return B2G_OK;
    const size_t sm1 = sizeof(float) * ((size_t)kBlock * contact_slots(s->params) * CF_COUNT + (size_t)kBlock * s->model.n_bodies * 3 + (kLinksShared ? 4 + link_store_floats(kBlock, s->model.n_dof) : 0));
    cudaFuncSetAttribute(k_houndarm_step<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    cudaFuncSetAttribute(k_houndarm_step<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    return alloc_task_buffers(s, 10, 6, 3 + s->model.n_dof);      // six actions (end-effector pose change), 3 command + n joint-noise draws per reset
}

int b2g_task_terrain_create(b2g_sim* s, const b2g_terrain_cfg* cfg, const int16_t* hs_host, const float* origins_host) {
    if (!s || !cfg) return fail(B2G_ERR_ARG, "null argument");
    if (!s->prepared) return fail(B2G_ERR_STATE, "task created before prepare_sim");
    if (s->model.fixed_base) return fail(B2G_ERR_UNSUPPORTED, "the terrain task needs a floating base");
    if (s->has_task) return fail(B2G_ERR_STATE, "a task already lives on this sim");
    const int nb = s->model.n_bodies, nd = s->model.n_dof;
    if (cfg->base_body < 0 || cfg->base_body >= nb || cfg->n_knee < 0 || cfg->n_knee > 8 || cfg->n_feet < 0 || cfg->n_feet > 4 ||
        cfg->n_term_extra < 0 || cfg->n_term_extra > 8 || cfg->n_hx < 1 || cfg->n_hx > 16 || cfg->n_hy < 1 || cfg->n_hy > 16)
        return fail(B2G_ERR_ARG, "bad body index counts / height grid");
    for (int k = 0; k < cfg->n_knee; k++) if (cfg->knee_bodies[k] < 0 || cfg->knee_bodies[k] >= nb) return fail(B2G_ERR_ARG, "bad knee body");
    for (int k = 0; k < cfg->n_feet; k++) if (cfg->feet_bodies[k] < 0 || cfg->feet_bodies[k] >= nb) return fail(B2G_ERR_ARG, "bad foot body");
    for (int k = 0; k < cfg->n_term_extra; k++) if (cfg->term_extra_bodies[k] < 0 || cfg->term_extra_bodies[k] >= nb) return fail(B2G_ERR_ARG, "bad termination body");
    if (cfg->custom_origins && (!hs_host || !origins_host || cfg->hs_rows < 2 || cfg->hs_cols < 2 || cfg->env_rows < 1 || cfg->env_cols < 1))
        return fail(B2G_ERR_ARG, "custom origins need height samples and terrain origins");
    s->tcfg = *cfg;
    s->seed = cfg->seed;
    s->task_kind = 3;
    s->n_cmd = 4;
    const bool arm = cfg->arm_chain >= 0;
    if (arm && (cfg->arm_chain >= s->model.n_chains || s->model.chain_len[cfg->arm_chain] != 6 || cfg->eef_body < 0 || cfg->eef_body >= nb ||
                cfg->jac_body < 0 || cfg->jac_body >= nb || cfg->n_ctrl_dof != s->model.chain_start[cfg->arm_chain]))
        return fail(B2G_ERR_ARG, "bad arm description (the arm must be the last chain, 6 DOF)");
    const int nctrl = cfg->n_ctrl_dof > 0 ? cfg->n_ctrl_dof : nd;
    const int nhp = cfg->n_hx * cfg->n_hy;
    const int no = 12 + 2 * nctrl + nhp + nd + (arm ? 10 : 0);
    int rc = alloc_task_buffers(s, no, nd, 2 * nctrl + 5 + (arm ? 6 : 0));
    if (rc != B2G_OK) return rc;
    return with_device(s, [&]() {
        const size_t n = s->n_envs;
        auto zalloc = [&](void** p, size_t bytes) -> cudaError_t {
            cudaError_t e = cudaMalloc(p, bytes ? bytes : 4);
            if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes ? bytes : 4);
            return e;
        };
        CUDA_TRY(zalloc((void**)&s->torques, sizeof(float) * n * nd));
        CUDA_TRY(zalloc((void**)&s->last_actions, sizeof(float) * n * nd));
        CUDA_TRY(zalloc((void**)&s->last_dof_vel, sizeof(float) * n * nd));
        CUDA_TRY(zalloc((void**)&s->feet_air_time, sizeof(float) * n * 4));
        CUDA_TRY(zalloc((void**)&s->episode_sums, sizeof(float) * n * 13));
        CUDA_TRY(zalloc((void**)&s->env_origins, sizeof(float) * n * 3));
        CUDA_TRY(zalloc((void**)&s->scratch9, sizeof(float) * n * 9));
        CUDA_TRY(zalloc((void**)&s->resetw, sizeof(float) * n));
        CUDA_TRY(zalloc((void**)&s->report, sizeof(float) * n * 13));
        CUDA_TRY(zalloc((void**)&s->cnorm, sizeof(float)));
        CUDA_TRY(zalloc((void**)&s->extras_part, sizeof(float) * 16 * (size_t)((n + kPostEnvs - 1) / kPostEnvs)));
        CUDA_TRY(zalloc((void**)&s->tickets, sizeof(unsigned) * 2));
        CUDA_TRY(zalloc((void**)&s->measured, sizeof(float) * n * nhp));
        CUDA_TRY(zalloc((void**)&s->noise_override, sizeof(float) * n * no));
        CUDA_TRY(zalloc((void**)&s->push_override, sizeof(float) * n * 2));
        CUDA_TRY(zalloc((void**)&s->extras, sizeof(float) * 16));
        CUDA_TRY(zalloc((void**)&s->terrain_levels, sizeof(long long) * n));
        CUDA_TRY(zalloc((void**)&s->terrain_types, sizeof(long long) * n));
        CUDA_TRY(zalloc((void**)&s->arm_mm, sizeof(float) * n * 36));
        CUDA_TRY(zalloc((void**)&s->arm_jac, sizeof(float) * n * 36));
        CUDA_TRY(zalloc((void**)&s->eef_state, sizeof(float) * n * 13));
        CUDA_TRY(zalloc((void**)&s->arm_commands, sizeof(float) * n * 3));
        if (hs_host && cfg->hs_rows > 0) {
            const size_t b = sizeof(int16_t) * (size_t)cfg->hs_rows * cfg->hs_cols;
            CUDA_TRY(cudaMalloc(&s->height_samples, b));
            CUDA_TRY(cudaMemcpy(s->height_samples, hs_host, b, cudaMemcpyHostToDevice));
        }
        if (origins_host && cfg->env_rows > 0) {
            const size_t b = sizeof(float) * (size_t)cfg->env_rows * cfg->env_cols * 3;
            CUDA_TRY(cudaMalloc(&s->terrain_origins, b));
            CUDA_TRY(cudaMemcpy(s->terrain_origins, origins_host, b, cudaMemcpyHostToDevice));
        }
        return (int)B2G_OK;
    });
}

int b2g_task_terrain_set_step(b2g_sim* s, int64_t step) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    s->common_step = step;
    if (s->auto_step && s->step_ctr)
        return with_device(s, [&]() {
            const long long v = step;
            CUDA_TRY(cudaMemcpy(s->step_ctr, &v, sizeof(v), cudaMemcpyHostToDevice));
            return (int)B2G_OK;
        });
    return B2G_OK;
}

int b2g_task_terrain_device_step(b2g_sim* s, int enable) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    if (!s->has_task || s->task_kind != 3) return fail(B2G_ERR_STATE, "no rough-terrain task created");
    return with_device(s, [&]() {
        if (enable) {
            if (!s->step_ctr) CUDA_TRY(cudaMalloc(&s->step_ctr, sizeof(long long)));
            const long long v = s->common_step + 1;      // the value the next step uses
            CUDA_TRY(cudaMemcpy(s->step_ctr, &v, sizeof(v), cudaMemcpyHostToDevice));
            s->common_step = v;
        } else if (s->auto_step) {
            s->common_step -= 1;                         // back to "last value used"; the caller sets the next one itself
        }
        s->auto_step = enable ? 1 : 0;
        return (int)B2G_OK;
    });
}

int b2g_task_terrain_set_init_done(b2g_sim* s, int v) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    s->init_done = v;
    return B2G_OK;
}

int b2g_task_tensor(b2g_sim* s, int kind, b2g_tensor_desc* d) {
    if (!s || !d) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_task) return fail(B2G_ERR_STATE, "no task created");
    const int64_t n = s->n_envs;
    d->device_id = s->device;
    d->dtype = 0;
    d->shape[0] = n; d->shape[1] = d->shape[2] = d->shape[3] = 1;
    d->ndim = 1;
    if (kind >= B2G_TT_TORQUES && s->task_kind != 3) return fail(B2G_ERR_ARG, "tensor kind %d belongs to the terrain task", kind);
    switch (kind) {
        case B2G_TT_OBS: d->data = s->obs; d->ndim = 2; d->shape[1] = s->num_obs; break;
        case B2G_TT_OBS_CLAMPED: d->data = s->obs_clamped; d->ndim = 2; d->shape[1] = s->num_obs; break;
        case B2G_TT_REW: d->data = s->rew; break;
        case B2G_TT_RESET: d->data = s->reset; d->dtype = 2; break;
        case B2G_TT_PROGRESS: d->data = s->progress; d->dtype = 2; break;
        case B2G_TT_TIMEOUT: d->data = s->timeout; d->dtype = 2; break;
        case B2G_TT_COMMANDS: d->data = s->commands; d->ndim = 2; d->shape[1] = s->n_cmd; break;
        case B2G_TT_ACTIONS: d->data = s->actions; d->ndim = 2; d->shape[1] = s->num_act; break;
        case B2G_TT_RAND_OVERRIDE: d->data = s->rand_override; d->ndim = 2; d->shape[1] = s->n_draws; break;
        case B2G_TT_TORQUES: d->data = s->torques; d->ndim = 2; d->shape[1] = s->num_act; break;
        case B2G_TT_LAST_ACTIONS: d->data = s->last_actions; d->ndim = 2; d->shape[1] = s->num_act; break;
        case B2G_TT_LAST_DOF_VEL: d->data = s->last_dof_vel; d->ndim = 2; d->shape[1] = s->num_act; break;
        case B2G_TT_FEET_AIR_TIME: d->data = s->feet_air_time; d->ndim = 2; d->shape[1] = 4; break;
        case B2G_TT_EPISODE_SUMS: d->data = s->episode_sums; d->ndim = 2; d->shape[0] = 13; d->shape[1] = n; break;
        case B2G_TT_ENV_ORIGINS: d->data = s->env_origins; d->ndim = 2; d->shape[1] = 3; break;
        case B2G_TT_TERRAIN_LEVELS: d->data = s->terrain_levels; d->dtype = 2; break;
        case B2G_TT_TERRAIN_TYPES: d->data = s->terrain_types; d->dtype = 2; break;
        case B2G_TT_NOISE_OVERRIDE: d->data = s->noise_override; d->ndim = 2; d->shape[1] = s->num_obs; break;
        case B2G_TT_PUSH_OVERRIDE: d->data = s->push_override; d->ndim = 2; d->shape[1] = 2; break;
        case B2G_TT_EXTRAS: d->data = s->extras; d->shape[0] = 16; break;
        case B2G_TT_MEASURED_HEIGHTS: d->data = s->measured; d->ndim = 2; d->shape[1] = s->tcfg.n_hx * s->tcfg.n_hy; break;
        case B2G_TT_ARM_MM: d->data = s->arm_mm; d->ndim = 3; d->shape[1] = 6; d->shape[2] = 6; break;
        case B2G_TT_ARM_JAC: d->data = s->arm_jac; d->ndim = 3; d->shape[1] = 6; d->shape[2] = 6; break;
        case B2G_TT_EEF_STATE: d->data = s->eef_state; d->ndim = 2; d->shape[1] = 13; break;
        case B2G_TT_ARM_COMMANDS: d->data = s->arm_commands; d->ndim = 2; d->shape[1] = 3; break;
        default: return fail(B2G_ERR_ARG, "unknown task tensor kind %d", kind);
    }
    return B2G_OK;
}

int b2g_task_set_rand_override(b2g_sim* s, int use) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    s->use_rand_override = use;
    return B2G_OK;
}

int b2g_task_anymal_reset_all(b2g_sim* s, void* stream) {
    if (!s) return fail(B2G_ERR_ARG, "null sim");
    if (!s->has_task || s->task_kind != 1) return fail(B2G_ERR_STATE, "no flat locomotion task created");
    return with_device(s, [&]() {
        const SimArgs A = make_args(s);
        const TaskArgs T = make_task_args(s, nullptr);
        const int grid = grid_size(s);
        if (s->v.lanes == 4) k_anymal_reset_all<4, 3><<<grid, kBlock, 0, (cudaStream_t)stream>>>(A, T);
        else if (s->v.seg) k_anymal_reset_all<8, 3><<<grid, kBlock, 0, (cudaStream_t)stream>>>(A, T);
        else k_anymal_reset_all<8, 6><<<grid, kBlock, 0, (cudaStream_t)stream>>>(A, T);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return (int)B2G_OK;
    });
}

int b2g_task_anymal_step(b2g_sim* s, const float* actions_dev, void* stream) {
    if (!s || !actions_dev) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_task) return fail(B2G_ERR_STATE, "no task created");
    return with_device(s, [&]() { return launch_anymal_step(s, actions_dev, (cudaStream_t)stream); });
}

int b2g_task_anymal_post_only(b2g_sim* s, const float* actions_dev, void* stream) {
    if (!s || !actions_dev) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_task) return fail(B2G_ERR_STATE, "no task created");
    return with_device(s, [&]() { return launch_anymal_step(s, actions_dev, (cudaStream_t)stream, 1); });
}

int b2g_task_host_layout(const b2g_sim* s, int64_t* offsets, int64_t* total_bytes) {
    if (!s || !offsets || !total_bytes) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_task) return fail(B2G_ERR_STATE, "no task created");
    for (int i = 0; i < 4; i++) offsets[i] = (int64_t)s->out_off[i];
    *total_bytes = (int64_t)s->out_total;
    return B2G_OK;
}

int b2g_task_anymal_step_host(b2g_sim* s, const float* actions_host, float* obs_host, float* rew_host, int64_t* reset_host,
                              int64_t* timeout_host, void* stream) {
    if (!s || !actions_host) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_task) return fail(B2G_ERR_STATE, "no task created");
    return with_device(s, [&]() {
        cudaStream_t st = (cudaStream_t)stream;
        const size_t n = s->n_envs, nd = s->num_act;
        // page-locked actions are read by the kernel in place (each thread loads its own few values once, over PCIe / C2C);
        // pageable ones are staged with a copy
        const float* actions_dev = s->actions_in;
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, actions_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
            actions_dev = static_cast<const float*>(pa.devicePointer);
        } else {
            cudaGetLastError();
            CUDA_TRY(cudaMemcpyAsync(s->actions_in, actions_host, sizeof(float) * n * nd, cudaMemcpyHostToDevice, st));
        }
        const unsigned char* base = reinterpret_cast<const unsigned char*>(obs_host);
        const bool packed = obs_host && rew_host && reset_host && timeout_host &&
                            reinterpret_cast<const unsigned char*>(rew_host) == base + s->out_off[1] &&
                            reinterpret_cast<const unsigned char*>(reset_host) == base + s->out_off[2] &&
                            reinterpret_cast<const unsigned char*>(timeout_host) == base + s->out_off[3];
        // packed page-locked outputs (b2g_task_host_layout): the SMs store the results into the caller's buffer and publish a
        // sequence number; no copy command, no stream synchronisation
        HostMirror H;
        static const bool mirror_on = !(getenv("B2G_HOST_MIRROR") && getenv("B2G_HOST_MIRROR")[0] == '0');   // =0: copy-engine path (A/B timing)
        if (mirror_on && packed && cudaPointerGetAttributes(&pa, obs_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
            (reinterpret_cast<size_t>(pa.devicePointer) & 15) == 0) {
            H.dst = static_cast<unsigned char*>(pa.devicePointer);
            H.src = s->out_arena;
            for (int i = 0; i < 4; i++) H.off[i] = s->out_off[i];
            H.num_obs = s->num_obs;
            H.done_ctr = s->done_ctr;
            H.flag = s->host_flag_dev;
            H.seq = ++s->host_seq;
            if (H.seq == 0) H.seq = ++s->host_seq;     // 0 is the initial value of the word
        } else {
            cudaGetLastError();
        }
        const bool fused_tail = H.dst && (s->task_kind == 1 || s->task_kind == 3);     // flat tasks: tail of k_anymal_step; rough terrain: tail of k_terrain_post
        const int rc = launch_anymal_step(s, actions_dev, st, 0, fused_tail ? &H : nullptr);
        if (rc != B2G_OK) return rc;
        if (H.dst) {
            if (!fused_tail) {
                const size_t n16 = (s->out_off[3] + sizeof(long long) * n + 15) / 16;     // the arena is padded to 256 B
                const int blocks = (int)((n16 + 255) / 256);
                k_mirror_host<<<blocks < 2 * s->n_sm ? blocks : 2 * s->n_sm, 256, 0, st>>>(H, n16);
                s->launches++;
                CUDA_TRY(cudaGetLastError());
            }
            volatile unsigned* flag = s->host_flag;
            for (unsigned spin = 1; *flag != H.seq; spin++) {
                if ((spin & 0xFFFFu) == 0) {      // a failed launch / device fault never publishes: ask the stream now and then
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q == cudaSuccess) {
                        if (*flag == H.seq) break;
                        CUDA_TRY(cudaStreamSynchronize(st));
                        if (*flag != H.seq) return fail(B2G_ERR_CUDA, "step_host: the stream drained without publishing the results");
                        break;
                    }
                    if (q != cudaErrorNotReady) return fail(B2G_ERR_CUDA, "step_host: %s", cudaGetErrorString(q));
                }
            }
            __sync_synchronize();
            return (int)B2G_OK;
        }
        if (packed) {   // pageable packed buffers: one copy
            CUDA_TRY(cudaMemcpyAsync(obs_host, s->out_arena, s->out_off[3] + sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        } else {
            if (obs_host) CUDA_TRY(cudaMemcpyAsync(obs_host, s->obs_clamped, sizeof(float) * n * s->num_obs, cudaMemcpyDeviceToHost, st));
            if (rew_host) CUDA_TRY(cudaMemcpyAsync(rew_host, s->rew, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
            if (reset_host) CUDA_TRY(cudaMemcpyAsync(reset_host, s->reset, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
            if (timeout_host) CUDA_TRY(cudaMemcpyAsync(timeout_host, s->timeout, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        return (int)B2G_OK;
    });
}

int b2g_task_step(b2g_sim* s, const float* actions_dev, void* stream) { return b2g_task_anymal_step(s, actions_dev, stream); }
int b2g_task_post_only(b2g_sim* s, const float* actions_dev, void* stream) { return b2g_task_anymal_post_only(s, actions_dev, stream); }
int b2g_task_osc_probe(b2g_sim* s, const float* actions_dev, void* stream) {
    if (!s || !actions_dev) return fail(B2G_ERR_ARG, "null argument");
    if (!s->has_task || s->task_kind != 3 || s->tcfg.arm_chain < 0) return fail(B2G_ERR_STATE, "no hound+arm task created");
    return with_device(s, [&]() { return launch_anymal_step(s, actions_dev, (cudaStream_t)stream, 2); });
}
int b2g_task_step_host(b2g_sim* s, const float* a, float* o, float* r, int64_t* rs, int64_t* to, void* stream) {
    return b2g_task_anymal_step_host(s, a, o, r, rs, to, stream);
}

int64_t b2g_sim_launch_count(const b2g_sim* s) { return s ? s->launches : 0; }

int b2g_sim_contact_stats(b2g_sim* s, int64_t* out, int reset) {
    if (!s || !out) return fail(B2G_ERR_ARG, "null argument");
    if (!s->d_stats) return fail(B2G_ERR_STATE, "contact statistics are switched off (B2G_CONTACT_STATS=0)");
    return with_device(s, [&]() -> int {
        unsigned long long h[4];
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(h, s->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 4; i++) out[i] = (int64_t)h[i];
        if (reset) CUDA_TRY(cudaMemset(s->d_stats, 0, sizeof(h)));
        return B2G_OK;
    });
}

// ---- DLPack export (dlpack v0.8 ABI, declared locally: plain C structs) ----
typedef struct { int32_t device_type; int32_t device_id; } B2gDLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } B2gDLDataType;
typedef struct {
    void* data; B2gDLDevice device; int32_t ndim; B2gDLDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset;
} B2gDLTensor;
typedef struct B2gDLManagedTensor {
    B2gDLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct B2gDLManagedTensor*);
} B2gDLManagedTensor;

static void b2g_dl_deleter(B2gDLManagedTensor* m) {
    if (!m) return;
    free(m->dl_tensor.shape);   // the data stays owned by the sim
    free(m);
}

// Wrap a tensor description as a DLManagedTensor (non-owning view; valid until b2g_sim_destroy).
int b2g_dlpack_from_desc(const b2g_tensor_desc* d, void** out_managed) {
    if (!d || !out_managed || !d->data) return fail(B2G_ERR_ARG, "null argument");
    B2gDLManagedTensor* m = (B2gDLManagedTensor*)calloc(1, sizeof(B2gDLManagedTensor));
    int64_t* shape = (int64_t*)calloc(4, sizeof(int64_t));
    if (!m || !shape) { free(m); free(shape); return fail(B2G_ERR_ARG, "out of host memory"); }
    for (int i = 0; i < d->ndim; i++) shape[i] = d->shape[i];
    m->dl_tensor.data = d->data;
    m->dl_tensor.device.device_type = 2;   // kDLCUDA
    m->dl_tensor.device.device_id = d->device_id;
    m->dl_tensor.ndim = d->ndim;
    switch (d->dtype) {
        case 0: m->dl_tensor.dtype = B2gDLDataType{2, 32, 1}; break;   // float32
        case 1: m->dl_tensor.dtype = B2gDLDataType{0, 32, 1}; break;   // int32
        case 2: m->dl_tensor.dtype = B2gDLDataType{0, 64, 1}; break;   // int64
        case 3: m->dl_tensor.dtype = B2gDLDataType{1, 8, 1}; break;    // uint8
        case 4: m->dl_tensor.dtype = B2gDLDataType{0, 16, 1}; break;   // int16
        default: free(m); free(shape); return fail(B2G_ERR_ARG, "bad dtype");
    }
    m->dl_tensor.shape = shape;
    m->dl_tensor.strides = nullptr;
    m->dl_tensor.byte_offset = 0;
    m->manager_ctx = nullptr;
    m->deleter = b2g_dl_deleter;
    *out_managed = m;
    return B2G_OK;
}

}  // extern "C"
