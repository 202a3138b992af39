// One physics sub-step for one environment, executed by a group of LANES cooperating threads (one
// thread per serial chain hanging off the root: a leg, the arm).  This is what replaces PhysX behind
// `gym.simulate` (reference: isaacgymenvs/tasks/base/vec_task.py:379-382, solver settings
// cfg/task/Anymal.yaml:81-100, drives tasks/anymal.py:199-203).
//
// Algorithm (DESIGN.md "dynamics"): Featherstone articulated-body algorithm in one common frame (world
// axes at the root origin, so no spatial transforms inside the recursion); implicit PD drives folded
// (and one-sided joint-limit springs) folded into the joint-space diagonal; hard contact solved at the velocity
// level by projected relaxation (Gauss-Seidel along a chain, Jacobi across chains) with impulses propagated through
// the articulated inertias; position iterations with penetration bias, positions integrated, then velocity
// iterations without bias; semi-implicit Euler.
//
// Parallel mapping: each lane owns one chain (kinematics, ABA recursion, its contact candidates); the
// root's 6x6 articulated inertia and bias are summed across the lanes with warp shuffles, the root
// solve is replicated, and in every contact sweep the lanes' root-level impulse biases are summed the same way.
#pragma once

#include "b2g_dev.h"
#include "b2g_math.cuh"

namespace b2g {

constexpr int MAXC = B2G_MAX_CONTACTS_PER_CHAIN;
// per-link loops of the recursions: fully unrolled for the 3-link legs (link state in registers: 41 vs 49 us per Anymal step),
// rolled for the long-chain variant (NL = 6: the unrolled kernel is 360 KB of SASS and spends 45 % of its stall cycles waiting for
// instructions; rolled, with the link state in local memory, it is 1/3 of the code and 13 % faster -- measured)
#if defined(B2G_ROLL_LINKS)
#define B2G_LINK_UNROLL _Pragma("unroll 1")
#elif defined(B2G_UNROLL_LINKS)
#define B2G_LINK_UNROLL _Pragma("unroll")
#else
#define B2G_LINK_UNROLL _Pragma("unroll (NL <= 3 ? NL : 1)")
#endif

// per-thread scratch for the (dynamically indexed) contact slots
enum ContactField {
    CF_RX, CF_RY, CF_RZ, CF_NX, CF_NY, CF_NZ, CF_T1X, CF_T1Y, CF_T1Z, CF_T2X, CF_T2Y, CF_T2Z,
    CF_GAP, CF_JC, CF_BODY, CF_ANN, CF_ANT1, CF_ANT2, CF_AT1T1, CF_AT1T2, CF_AT2T2, CF_ATISO, CF_LN, CF_L1, CF_L2, CF_COUNT
};

#if defined(B2G_HOST_EMU)
// ---- host lane emulator (tests only): LANES host threads in lock-step through a spin barrier ----
struct EmuGroup {
    int lanes;
    volatile int count;
    volatile int sense;
    float fslot[8];
    int islot[8];
};
struct EmuCtx {
    EmuGroup* g;
    int lane;
    int local_sense;
};
extern thread_local EmuCtx emu_ctx;
extern long long emu_hfc_tests, emu_hfc_skips;   // coarse-bound statistics (tests assert the early-out is not vacuous)
inline void emu_barrier() {
    EmuGroup* g = emu_ctx.g;
    if (g->lanes == 1) return;
    int s = emu_ctx.local_sense ^= 1;
    if (__atomic_add_fetch(&g->count, 1, __ATOMIC_ACQ_REL) == g->lanes) {
        g->count = 0;
        __atomic_store_n(&g->sense, s, __ATOMIC_RELEASE);
    } else {
        while (__atomic_load_n(&g->sense, __ATOMIC_ACQUIRE) != s) {
        }
    }
}
template <int LANES>
struct Grp {
    static float bcast(float x, int src) {
        if (LANES == 1) return x;
        emu_ctx.g->fslot[emu_ctx.lane] = x;
        emu_barrier();
        float r = emu_ctx.g->fslot[src];
        emu_barrier();
        return r;
    }
    static float sum(float x) {
        if (LANES == 1) return x;
        // same butterfly order as the device xor-shuffle reduction
        for (int o = LANES / 2; o > 0; o >>= 1) {
            emu_ctx.g->fslot[emu_ctx.lane] = x;
            emu_barrier();
            float other = emu_ctx.g->fslot[emu_ctx.lane ^ o];
            emu_barrier();
            x += other;
        }
        return x;
    }
    static int warp_max(int x) {
        if (LANES == 1) return x;
        emu_ctx.g->islot[emu_ctx.lane] = x;
        emu_barrier();
        int r = x;
        for (int i = 0; i < LANES; i++) r = emu_ctx.g->islot[i] > r ? emu_ctx.g->islot[i] : r;
        emu_barrier();
        return r;
    }
    static bool warp_any(bool p) { return warp_max(p ? 1 : 0) != 0; }
    static void sync() { emu_barrier(); }
};
#else
// Warp layout of a lane group.  LANES <= 4: consecutive lanes.  LANES == 8 (the long-chain variant, 4 environments per warp):
// chains 0-3 of the four environments sit in lanes 0-15 (lane = 4 e + c), chains 4-7 in lanes 16-31 chain-major
// (lane = 16 + 4 (c - 4) + e).  The per-thread link state of that variant lives in local memory, whose lines interleave the 32
// lanes word by word: with a hound + arm (chains 0-3 = legs, 4 = arm, 5-7 unused) the legs fill two 32-byte sectors and the four
// arms share one, instead of every environment's 8 lanes half-filling a sector of its own.  The reduction pairs the same chains in
// the same order as a plain xor butterfly (c ^ 4, c ^ 2, c ^ 1), so sums are bit-identical to the consecutive layout.
#if defined(B2G_GRP8_CONSECUTIVE)      // A/B switch: the plain consecutive layout for the 8-lane variant as well
constexpr int kSplit8 = 0;
#else
constexpr int kSplit8 = 8;
#endif
template <int LANES>
struct Grp {
    static __device__ __forceinline__ int phys_lane(int chain) {      // physical lane of chain `chain` of the caller's environment
        const int wl = threadIdx.x & 31;
        if (LANES != kSplit8) return (wl / LANES) * LANES + chain;
        const int e = wl < 16 ? wl >> 2 : wl & 3;
        return chain < 4 ? 4 * e + chain : 16 + 4 * (chain - 4) + e;
    }
    static __device__ __forceinline__ float bcast(float x, int src) {
        if (LANES == 1) return x;
        return __shfl_sync(0xffffffffu, x, phys_lane(src));
    }
    static __device__ __forceinline__ float sum(float x) {
        if (LANES == kSplit8) {
            const int wl = threadIdx.x & 31, up = wl >> 4;
            x += __shfl_sync(0xffffffffu, x, up ? 4 * (wl & 3) + ((wl - 16) >> 2) : 16 + 4 * (wl & 3) + (wl >> 2));   // c ^ 4
            x += __shfl_xor_sync(0xffffffffu, x, up ? 8 : 2);                                                            // c ^ 2
            x += __shfl_xor_sync(0xffffffffu, x, up ? 4 : 1);                                                            // c ^ 1
            return x;
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    }
    static __device__ __forceinline__ int warp_max(int x) { return __reduce_max_sync(0xffffffffu, x); }
    static __device__ __forceinline__ bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
};
#endif

template <int LANES>
B2G_HD B2G_INL SV grp_bcast(SV s, int src) {
    return SV{V3{Grp<LANES>::bcast(s.w.x, src), Grp<LANES>::bcast(s.w.y, src), Grp<LANES>::bcast(s.w.z, src)},
              V3{Grp<LANES>::bcast(s.v.x, src), Grp<LANES>::bcast(s.v.y, src), Grp<LANES>::bcast(s.v.z, src)}};
}
template <int LANES>
B2G_HD B2G_INL SI grp_bcast(const SI& s, int src) {
    SI o;
    o.A.xx = Grp<LANES>::bcast(s.A.xx, src); o.A.xy = Grp<LANES>::bcast(s.A.xy, src); o.A.xz = Grp<LANES>::bcast(s.A.xz, src);
    o.A.yy = Grp<LANES>::bcast(s.A.yy, src); o.A.yz = Grp<LANES>::bcast(s.A.yz, src); o.A.zz = Grp<LANES>::bcast(s.A.zz, src);
    o.C.xx = Grp<LANES>::bcast(s.C.xx, src); o.C.xy = Grp<LANES>::bcast(s.C.xy, src); o.C.xz = Grp<LANES>::bcast(s.C.xz, src);
    o.C.yy = Grp<LANES>::bcast(s.C.yy, src); o.C.yz = Grp<LANES>::bcast(s.C.yz, src); o.C.zz = Grp<LANES>::bcast(s.C.zz, src);
#pragma unroll
    for (int k = 0; k < 9; k++) o.B.m[k] = Grp<LANES>::bcast(s.B.m[k], src);
    return o;
}
template <int LANES>
B2G_HD B2G_INL SV grp_sum(SV s) {
    return SV{V3{Grp<LANES>::sum(s.w.x), Grp<LANES>::sum(s.w.y), Grp<LANES>::sum(s.w.z)},
              V3{Grp<LANES>::sum(s.v.x), Grp<LANES>::sum(s.v.y), Grp<LANES>::sum(s.v.z)}};
}
template <int LANES>
B2G_HD B2G_INL SI grp_sum(const SI& s) {
    SI o;
    o.A.xx = Grp<LANES>::sum(s.A.xx); o.A.xy = Grp<LANES>::sum(s.A.xy); o.A.xz = Grp<LANES>::sum(s.A.xz);
    o.A.yy = Grp<LANES>::sum(s.A.yy); o.A.yz = Grp<LANES>::sum(s.A.yz); o.A.zz = Grp<LANES>::sum(s.A.zz);
    o.C.xx = Grp<LANES>::sum(s.C.xx); o.C.xy = Grp<LANES>::sum(s.C.xy); o.C.xz = Grp<LANES>::sum(s.C.xz);
    o.C.yy = Grp<LANES>::sum(s.C.yy); o.C.yz = Grp<LANES>::sum(s.C.yz); o.C.zz = Grp<LANES>::sum(s.C.zz);
#pragma unroll
    for (int k = 0; k < 9; k++) o.B.m[k] = Grp<LANES>::sum(s.B.m[k]);
    return o;
}

// contact statistics of one sub-step (DevParams::stats): warp-aggregated, at most four reductions without return value per warp
template <int LANES>
B2G_HD B2G_INL void contact_stats_add(unsigned long long* stats, int lane, bool live, int ncon, int ndrop) {
    const float gdrop = Grp<LANES>::sum((float)ndrop);      // dropped candidates of the whole environment
    const int act = live ? ncon : 0, drop = live ? ndrop : 0;
    const int envdrop = (live && lane == 0 && gdrop > 0.0f) ? 1 : 0, envs = (live && lane == 0) ? 1 : 0;
#if defined(B2G_HOST_EMU)
    if (act) __atomic_add_fetch(stats + 0, (unsigned long long)act, __ATOMIC_RELAXED);
    if (drop) __atomic_add_fetch(stats + 1, (unsigned long long)drop, __ATOMIC_RELAXED);
    if (envdrop) __atomic_add_fetch(stats + 2, 1ull, __ATOMIC_RELAXED);
    if (envs) __atomic_add_fetch(stats + 3, 1ull, __ATOMIC_RELAXED);
#else
    const unsigned full = 0xffffffffu;
    const int wa = __reduce_add_sync(full, act), wd = __reduce_add_sync(full, drop), we = __reduce_add_sync(full, envdrop), wn = __reduce_add_sync(full, envs);
    if ((threadIdx.x & 31) == 0) {
        if (wa) atomicAdd(stats + 0, (unsigned long long)wa);
        if (wd) atomicAdd(stats + 1, (unsigned long long)wd);
        if (we) atomicAdd(stats + 2, (unsigned long long)we);
        if (wn) atomicAdd(stats + 3, (unsigned long long)wn);
    }
#endif
}

// state one lane keeps in registers across the sub-steps of a simulate() call
template <int NL>
struct LaneState {
    V3 rp;                  // root position (world)
    float qx, qy, qz, qw;   // root orientation
    V3 rv, rw;              // root linear / angular velocity (world)
    float q[NL], qd[NL];    // this lane's chain
    float tgt[NL], act[NL]; // drive targets / efforts
    float frc[NL];          // DOF force output
};

// per-environment physical randomisation (tensorised domain randomisation, reference vec_task.py:610-840 /
// cfg/task/Anymal.yaml:104-170): shape friction, and scale factors on every link mass (+ inertia), drive stiffness and damping
struct EnvDr {
    float mu = 1.0f, mass = 1.0f, kp = 1.0f, kd = 1.0f;
    bool live = true;      // false for the padding threads of the last block (they shadow the last environment): not counted in the statistics
    const float* link = nullptr;   // this environment's (nd + 1, B2G_LINK_SCALE_COLS) per-link rows (B2G_T_LINK_SCALE) or null
    // combined scales / limit offsets of link l (0 = root, 1 + d = child of DOF d)
    B2G_HD B2G_INL float mass_of(int l) const { return link ? mass * link[B2G_LINK_SCALE_COLS * l] : mass; }
    B2G_HD B2G_INL float kp_of(int l) const { return link ? kp * link[B2G_LINK_SCALE_COLS * l + 1] : kp; }
    B2G_HD B2G_INL float kd_of(int l) const { return link ? kd * link[B2G_LINK_SCALE_COLS * l + 2] : kd; }
    B2G_HD B2G_INL float lower_of(int l) const { return link ? link[B2G_LINK_SCALE_COLS * l + 3] : 0.0f; }
    B2G_HD B2G_INL float upper_of(int l) const { return link ? link[B2G_LINK_SCALE_COLS * l + 4] : 0.0f; }
};
B2G_HD B2G_INL EnvDr load_env_dr(const float* friction, const float* env_scale, int env, bool live = true, const float* link_scale = nullptr,
                                 int n_links = 0) {
    EnvDr d;
    d.live = live;
    if (link_scale) d.link = link_scale + (size_t)env * n_links * B2G_LINK_SCALE_COLS;
    if (friction) d.mu = friction[env];
    if (env_scale) {
        d.mass = env_scale[(size_t)env * 4 + 0];
        d.kp = env_scale[(size_t)env * 4 + 1];
        d.kd = env_scale[(size_t)env * 4 + 2];
    }
    return d;
}

// per-link state of the recursions.  Members are grouped in 16-byte units so that the shared-memory variant moves them with
// 128-bit accesses; 11 units per link (one of padding) keeps the eight lanes of an environment, which walk their chains side
// by side, on different bank quads.
struct alignas(16) LinkData {
    M3 Rl;             // link rotation (world axes)
    V3 pl;             // link origin relative to the root origin
    SV S, cb;          // joint motion subspace, velocity-product term
    SV vl, U;          // link velocity, I^A S
    float tau, dext;   // joint-space force, joint-space diagonal (armature + implicit drive / limit terms)
    float Dinv, u;
    float pad[4];
};
// Rolled link loops (NL > 3) index the link state dynamically, so as per-thread arrays it lives in local memory (1056 B per
// thread in lane-interleaved lines, more than L1 holds for a resident wave: 28-33 % L1 misses measured on UsefulHound).
// -DB2G_LINKS_SHARED moves it to shared memory instead (one LinkData per DOF of the environment, slot = DOF index).  Measured on
// B200: the stall class halves but the extra instructions cost more -- UsefulHound 372 us per step against
// 350 us for the aligned per-thread arrays -- so the local-memory form is the default and the shared form an experiment switch.
static_assert(sizeof(LinkData) == 44 * sizeof(float), "LinkData is 11 x 16 bytes");
template <int NL>
B2G_HD constexpr bool link_loops_rolled() {
#if defined(B2G_ROLL_LINKS)
    return true;
#elif defined(B2G_UNROLL_LINKS)
    return false;
#else
    return NL > 3;
#endif
}
#if defined(B2G_LINKS_SHARED)
constexpr bool kLinksShared = true;
#else
constexpr bool kLinksShared = false;
#endif
template <int NL>
B2G_HD constexpr bool links_in_shared() {
#if defined(B2G_LINKS_SHARED)
    return link_loops_rolled<NL>();
#else
    return false;
#endif
}
#if defined(B2G_HOST_EMU)
inline
#else
__host__ __device__ inline
#endif
size_t link_store_floats(int n_envs_per_block, int n_dof) { return (size_t)n_envs_per_block * n_dof * (sizeof(LinkData) / sizeof(float)); }

// Segment variant (LANES = 8, NL = 3): a chain longer than three links is cut into pieces that own a lane each (DevModel::seg_*).
// The recursions run in two dependency stages with the piece boundary passed by shuffles; a distal piece reaches the links of its
// parent piece (joint axis S, U = I^A S, 1 / D) through shared memory when it propagates a contact impulse to the root.
template <int LANES, int NL>
B2G_HD constexpr bool is_segmented() { return LANES == 8 && NL == 3; }
constexpr int kSelfJc = 64;                           // CF_JC of a self-collision slot = link index + kSelfJc
constexpr int kAncLinkFloats = 16;                    // S (6) | U (6) | 1/D | 3 pad: 16-byte rows
constexpr int kAncFloats = 3 * kAncLinkFloats;        // a piece with a child is always full

#if defined(B2G_HOST_EMU)
#define B2G_NOINLINE
#else
#define B2G_NOINLINE __noinline__
#endif

// per-thread contact scratch: field f of slot s
struct ScratchStrided {
    float* base;   // points at this thread's column
    int stride;    // number of threads sharing the buffer
    float* links = nullptr;   // this environment's slice of the block's shared link store (links_in_shared variants only)
    float* anc = nullptr;     // this environment's ancestor store (segment variant only): kAncFloats per piece that has a child
    B2G_HD B2G_INL float& at(int slot, int f) { return base[(slot * CF_COUNT + f) * stride]; }
};

// slot with the largest gap (the shallowest contact) among the first `nslots`; out of line: only reached when a lane's slots are full
B2G_HD B2G_NOINLINE int shallowest_slot(const float* base, int stride, int nslots, float* gap_out) {
    int w = 0;
    float gw = base[(0 * CF_COUNT + CF_GAP) * stride];
    for (int q = 1; q < nslots; q++) {
        const float g = base[(q * CF_COUNT + CF_GAP) * stride];
        if (g > gw) { gw = g; w = q; }
    }
    *gap_out = gw;
    return w;
}

B2G_HD B2G_INL void ground_sample(const DevParams& P, float x, float y, float& hgt, V3& n) {
    float gx = (x - P.hf_ox) / P.hf_hs, gy = (y - P.hf_oy) / P.hf_hs;
    gx = fminf(fmaxf(gx, 0.0f), (float)(P.hf_rows - 1));
    gy = fminf(fmaxf(gy, 0.0f), (float)(P.hf_cols - 1));
    int i = (int)gx, j = (int)gy;
    i = i > P.hf_rows - 2 ? P.hf_rows - 2 : i;
    j = j > P.hf_cols - 2 ? P.hf_cols - 2 : j;
    float fx = gx - (float)i, fy = gy - (float)j;
    const int16_t* s = P.hf + (size_t)i * P.hf_cols + j;
    float h00 = P.hf_vs * (float)s[0], h01 = P.hf_vs * (float)s[1];
    float h10 = P.hf_vs * (float)s[P.hf_cols], h11 = P.hf_vs * (float)s[P.hf_cols + 1];
    float dzdx, dzdy;
    if (fx >= fy) { dzdx = h10 - h00; dzdy = h11 - h10; }
    else { dzdx = h11 - h01; dzdy = h01 - h00; }
    hgt = h00 + dzdx * fx + dzdy * fy;
    float nx = -dzdx / P.hf_hs, ny = -dzdy / P.hf_hs;
    float inv = 1.0f / sqrtf(nx * nx + ny * ny + 1.0f);
    n = V3{nx * inv, ny * inv, inv};
}

// The warps of a block run the same instruction stream, each through its own instruction-cache lines once it has drifted away from
// the others (data-dependent contact loops).  A block barrier at a few points of the sub-step keeps them within one cache window.
B2G_HD B2G_INL void block_align(const DevParams& P, int bit) {
#if defined(__CUDA_ARCH__)
    if (P.block_align & bit) __syncthreads();
#else
    (void)P; (void)bit;
#endif
}

// One sub-step.  `len` = links in this lane's chain (0 for an idle lane), `d0` = first DOF of the chain.
// `bf` = this environment's body-force accumulator (n_bodies*3 floats, shared by the group), written on
// the last sub-step only.
// FULL = true promises len == NL on every lane.  PROBE = true turns the call into the forward-dynamics probe of the parity tests: efforts st.act are
// applied raw (no drives), the function returns after the ABA with joint accelerations in st.frc and the
// root's spatial acceleration (angular, linear) in st.rw / st.rv.
// SC = false compiles the self-collision code out (the fused flat-terrain step of the quadrupeds, whose tasks never enable it).
template <int LANES, int NL, bool FIXED, bool HF, bool PROBE = false, bool FULL = false, bool SC = true>
B2G_HD B2G_INL void substep(const DevModel* __restrict__ M, const DevParams& P, int lane, int len_in, int d0,
                            LaneState<NL>& st, const EnvDr dr, bool last, ScratchStrided sc, float* bf) {
    // FULL: every lane's chain has exactly NL links (the quadrupeds) -> the `j < len` predicates fold away
    const int len = FULL ? NL : len_in;
    constexpr bool SEG = is_segmented<LANES, NL>();
    constexpr int NSTG = SEG ? 2 : 1;                          // dependency stages of the recursions: proximal pieces, then distal ones
    const int par = SEG ? M->seg_par[lane] : -1;               // lane of the piece this lane's piece hangs off
    const int child = SEG ? M->seg_child[lane] : -1;           // lane of the piece hanging off this one
    const bool distal = SEG && par >= 0;
    const int up_lane = distal ? par : lane, down_lane = child >= 0 ? child : lane;      // shuffle sources (self = no-op)
    const float h = P.h;
    const V3 grav = V3{P.g[0], P.g[1], P.g[2]};
    block_align(P, 1);

    // ---------------- kinematics down the chain ----------------
    const M3 R0 = quat_to_m3(st.qx, st.qy, st.qz, st.qw);
    const SV v0 = FIXED ? sv0() : SV{st.rw, st.rv};
    // per-link state of this lane's chain: registers when the link loops are unrolled, this lane's slice of the block's
    // shared-memory link store (sc.links) when they are rolled
    constexpr bool LSH = links_in_shared<NL>();
    LinkData Lreg[LSH ? 1 : NL];
    LinkData* const L = LSH ? reinterpret_cast<LinkData*>(sc.links) + d0 : Lreg;   // sc.links = this environment's store, slot = DOF index
    {
        M3 Rp = R0;
        V3 pp = V3{0, 0, 0};
        SV vp = v0;
#pragma unroll 1
        for (int stg = 0; stg < NSTG; stg++) {
        if (SEG && stg == 1) {      // distal pieces start from the frame and velocity of their parent's last link
            M3 Rq;
#pragma unroll
            for (int k = 0; k < 9; k++) Rq.m[k] = Grp<LANES>::bcast(Rp.m[k], up_lane);
            const V3 pq = V3{Grp<LANES>::bcast(pp.x, up_lane), Grp<LANES>::bcast(pp.y, up_lane), Grp<LANES>::bcast(pp.z, up_lane)};
            const SV vq = grp_bcast<LANES>(vp, up_lane);
            Rp = Rq; pp = pq; vp = vq;
        }
        if (SEG && distal != (stg == 1)) continue;
B2G_LINK_UNROLL
        for (int j = 0; j < NL; j++) {
            if (j < len) {
                const DevDof& D = M->dof[d0 + j];
                M3 jr;
#pragma unroll
                for (int k = 0; k < 9; k++) jr.m[k] = D.jrot[k];
                const V3 ax = V3{D.axis[0], D.axis[1], D.axis[2]};
                const M3 RJ = mul(Rp, jr);
                const V3 pj = pp + mul(Rp, V3{D.jpos[0], D.jpos[1], D.jpos[2]});
                const V3 axw = mul(RJ, ax);
                if (D.type == B2G_JOINT_REVOLUTE) {
                    L[j].Rl = mul(RJ, axis_angle_m3(ax, st.q[j]));
                    L[j].pl = pj;
                    L[j].S = SV{axw, cross(pj, axw)};
                } else {
                    L[j].Rl = RJ;
                    L[j].pl = pj + axw * st.q[j];
                    L[j].S = SV{V3{0, 0, 0}, axw};
                }
                const SV vj = L[j].S * st.qd[j];
                L[j].vl = vp + vj;
                L[j].cb = crm(L[j].vl, vj);
                // drive
                const float kp = D.kp * dr.kp_of(1 + d0 + j), kd = D.kd * dr.kd_of(1 + d0 + j);
                float t = 0.0f, de = D.armature;
                if (PROBE) {
                    t = st.act[j];
                } else if (D.drive_mode == B2G_DOF_MODE_POS) {
                    // implicit PD drive; saturated at the effort limit (PhysX clamps the drive force to maxForce): when the torque estimated at
                    // the end-of-step position exceeds it, the drive is a constant torque of that size and contributes no implicit terms
                    const float te = kp * (st.tgt[j] - st.q[j] - h * st.qd[j]) - kd * st.qd[j];
                    if (D.effort > 0.0f && fabsf(te) > D.effort) {
                        t = copysignf(D.effort, te);
                    } else {
                        t = kp * (st.tgt[j] - st.q[j]) - (kd + h * kp) * st.qd[j];
                        de += h * kd + h * h * kp;
                    }
                } else if (D.drive_mode == B2G_DOF_MODE_VEL) {
                    const float te = kd * (st.tgt[j] - st.qd[j]);
                    if (D.effort > 0.0f && fabsf(te) > D.effort) {
                        t = copysignf(D.effort, te);
                    } else {
                        t = te;
                        de += h * kd;
                    }
                } else if (D.drive_mode == B2G_DOF_MODE_EFFORT) {
                    t = st.act[j];
                    if (D.effort > 0.0f) t = fminf(fmaxf(t, -D.effort), D.effort);
                }
                // joint limit: one-sided implicit spring-damper, folded in exactly like the implicit drive
                {
                    const float lo = D.lower + dr.lower_of(1 + d0 + j), hi = D.upper + dr.upper_of(1 + d0 + j), qp = st.q[j] + h * st.qd[j];
                    float ref = 0.0f;
                    bool on = false;
                    if (lo > -1e30f && (st.q[j] < lo || qp < lo)) { ref = lo; on = true; }
                    else if (hi < 1e30f && (st.q[j] > hi || qp > hi)) { ref = hi; on = true; }
                    if (on && !PROBE) {
                        t += P.limit_kp * (ref - st.q[j]) - (P.limit_kd + h * P.limit_kp) * st.qd[j];
                        de += h * P.limit_kd + h * h * P.limit_kp;
                    }
                }
                L[j].tau = t;
                L[j].dext = de;
                Rp = L[j].Rl; pp = L[j].pl; vp = L[j].vl;
            } else if (!link_loops_rolled<NL>()) {      // unused link slots are never read; in registers they are defined anyway (free once unrolled)
                L[j].Rl = R0; L[j].pl = V3{0, 0, 0}; L[j].S = sv0(); L[j].cb = sv0(); L[j].vl = sv0(); L[j].tau = 0; L[j].dext = 1.0f;
            }
        }
        }
    }

    // ---------------- ABA backward: articulated inertias and bias forces ----------------
    SI IAc;
    SV pAc = sv0();
    {
        IAc.A = S3{0, 0, 0, 0, 0, 0}; IAc.C = S3{0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 9; k++) IAc.B.m[k] = 0;
#pragma unroll 1
        for (int stg = NSTG - 1; stg >= 0; stg--) {
        if (SEG && stg == 0) {      // a distal piece hands its articulated inertia and bias force to its parent's last link
            const SI Iq = grp_bcast<LANES>(IAc, down_lane);
            const SV pq = grp_bcast<LANES>(pAc, down_lane);
            if (child >= 0) { IAc = Iq; pAc = pq; }
        }
        if (SEG && distal != (stg == 1)) continue;
B2G_LINK_UNROLL
        for (int j = NL - 1; j >= 0; j--) {
            if (j < len) {
                const DevDof& D = M->dof[d0 + j];
                const V3 cw = L[j].pl + mul(L[j].Rl, V3{D.com[0], D.com[1], D.com[2]});
                const float ms = dr.mass_of(1 + d0 + j);
                const S3 iw = rotate_sym(L[j].Rl, S3{D.inertia[0] * ms, D.inertia[1] * ms, D.inertia[2] * ms, D.inertia[3] * ms, D.inertia[4] * ms, D.inertia[5] * ms});
                SI I = rigid_inertia(D.mass * ms, cw, iw);
                SV pA = crf(L[j].vl, mul(I, L[j].vl));
                if (j + 1 < len || (SEG && child >= 0)) { I += IAc; pA += pAc; }
                L[j].U = mul(I, L[j].S);
                const float Dj = dot(L[j].S, L[j].U) + L[j].dext;
                L[j].Dinv = 1.0f / Dj;
                L[j].u = L[j].tau - dot(L[j].S, pA);
                IAc = rank1_sub(I, L[j].U, L[j].Dinv);
                pAc = pA + mul(IAc, L[j].cb) + L[j].U * (L[j].u * L[j].Dinv);
            } else if (!link_loops_rolled<NL>()) {
                L[j].U = sv0(); L[j].Dinv = 0.0f; L[j].u = 0.0f;
            }
        }
        }
    }

    // ---------------- root: sum the chains, solve ----------------
    P6 inv;
    SV a0 = sv0();   // relative root acceleration a' = a - a_g
    if (!FIXED) {
        if (SEG && distal) {      // already inside its parent's terms
            pAc = sv0();
            IAc.A = S3{0, 0, 0, 0, 0, 0}; IAc.C = S3{0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 9; k++) IAc.B.m[k] = 0;
        }
        SI IA0 = grp_sum<LANES>(IAc);
        SV pA0 = grp_sum<LANES>(pAc);
        const V3 cw = mul(R0, V3{M->root_com[0], M->root_com[1], M->root_com[2]});
        const float ms0 = dr.mass_of(0);
        const S3 iw = rotate_sym(R0, S3{M->root_inertia[0] * ms0, M->root_inertia[1] * ms0, M->root_inertia[2] * ms0, M->root_inertia[3] * ms0,
                                        M->root_inertia[4] * ms0, M->root_inertia[5] * ms0});
        SI I0 = rigid_inertia(M->root_mass * ms0, cw, iw);
        pA0 += crf(v0, mul(I0, v0));
        IA0 += I0;
        bool ok;
        inv = spd_inverse6(pack6(IA0), ok);
        a0 = -mul(inv, pA0);
    } else {
#pragma unroll
        for (int k = 0; k < 21; k++) inv.a[k] = 0.0f;
    }

    // ---------------- ABA forward: accelerations, free velocities ----------------
    float qdn[NL];
    SV v0n = sv0();
    {
        SV a = FIXED ? SV{V3{0, 0, 0}, -grav} : a0;
#pragma unroll 1
        for (int stg = 0; stg < NSTG; stg++) {
        if (SEG && stg == 1) a = grp_bcast<LANES>(a, up_lane);      // acceleration of the parent's last link
        if (SEG && distal != (stg == 1)) continue;
B2G_LINK_UNROLL
        for (int j = 0; j < NL; j++) {
            if (j < len) {
                const SV ap = a + L[j].cb;
                const float qdd = (L[j].u - dot(L[j].U, ap)) * L[j].Dinv;
                a = ap + L[j].S * qdd;
                qdn[j] = st.qd[j] + h * qdd;
                if (PROBE) st.frc[j] = qdd;
            } else {
                qdn[j] = 0.0f;
            }
        }
        }
        if (!FIXED) {
            // root origin is a body-fixed point: classical acceleration = spatial + w x v
            v0n.w = st.rw + a0.w * h;
            v0n.v = st.rv + (a0.v + grav + cross(st.rw, st.rv)) * h;
        }
    }

    if (PROBE) {
        st.rw = a0.w;
        st.rv = FIXED ? V3{0, 0, 0} : a0.v + grav;
        return;
    }

    // impulse helpers -----------------------------------------------------------------
    // backward half: impulse F on link `jc` of this chain (-1 = root): returns the root-level bias Pb and the per-joint
    // terms ud[] the forward half needs
    // Segment variant: pieces with a child publish S, U, 1/D of their links; a distal piece continues its impulse recursion through them.
    // vbase = velocity of the link the piece hangs off (the root's for proximal pieces, where v0n is used instead).
    SV vbase = v0n;
    const float* anc = nullptr;
    if (SEG) {
        Grp<LANES>::sync();
        const int slot = M->seg_store[lane];
        if (slot >= 0) {
#pragma unroll
            for (int j = 0; j < NL; j++) {
                float* a = sc.anc + slot * kAncFloats + j * kAncLinkFloats;
                a[0] = L[j].S.w.x; a[1] = L[j].S.w.y; a[2] = L[j].S.w.z; a[3] = L[j].S.v.x; a[4] = L[j].S.v.y; a[5] = L[j].S.v.z;
                a[6] = L[j].U.w.x; a[7] = L[j].U.w.y; a[8] = L[j].U.w.z; a[9] = L[j].U.v.x; a[10] = L[j].U.v.y; a[11] = L[j].U.v.z;
                a[12] = L[j].Dinv;
            }
        }
        SV vend = v0n;      // free velocity of this piece's last link
#pragma unroll
        for (int j = 0; j < NL; j++)
            if (j < len) vend += L[j].S * qdn[j];
        const SV vq = grp_bcast<LANES>(vend, up_lane);
        if (distal) { vbase = vq; anc = sc.anc + M->seg_store[par] * kAncFloats; }
        Grp<LANES>::sync();
    }
    auto anc_S = [&](int i) -> SV { const float* a = anc + i * kAncLinkFloats; return SV{V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}}; };
    auto anc_U = [&](int i) -> SV { const float* a = anc + i * kAncLinkFloats; return SV{V3{a[6], a[7], a[8]}, V3{a[9], a[10], a[11]}}; };
    auto anc_Dinv = [&](int i) -> float { return anc[i * kAncLinkFloats + 12]; };
    auto push_up = [&](int jc, SV F, float* ud, float* uda) -> SV {
        SV Pb = -F;
#pragma unroll
        for (int i = NL - 1; i >= 0; i--) {
            if (i <= jc) {
                ud[i] = -dot(L[i].S, Pb);
                Pb += L[i].U * (ud[i] * L[i].Dinv);
            } else {
                ud[i] = 0.0f;
            }
        }
        if (SEG && distal) {      // on through the parent piece
#pragma unroll
            for (int i = 2; i >= 0; i--) {
                uda[i] = -dot(anc_S(i), Pb);
                Pb += anc_U(i) * (uda[i] * anc_Dinv(i));
            }
        }
        return Pb;
    };
    // forward half through the parent piece's links (distal pieces)
    auto run_anc = [&](SV& a, const float* uda) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float dq = (uda[i] - dot(anc_U(i), a)) * anc_Dinv(i);
            a += anc_S(i) * dq;
        }
    };

    // ---------------- contact candidates ----------------
    if (!PROBE) block_align(P, 2);
    int ncon = 0, ndrop = 0;
    const bool ground = HF || (P.has_ground != 0);
    // A candidate inside the contact offset takes a free slot; when the lane's slots are full it is counted as dropped -- but the DEEPEST
    // candidates are the ones kept: it replaces the slot with the largest gap if it penetrates further than that one (PhysX's contact
    // reduction keeps the deepest points too).  A candidate left without a slot while it sinks would otherwise be picked up centimetres
    // deep and pushed out at max_depenetration_velocity.
    auto claim_slot = [&](float gap) -> int {
        if (ncon < P.max_contacts) return ncon++;
        ndrop++;
        float gw;
        const int w = shallowest_slot(sc.base, sc.stride, P.max_contacts, &gw);
        return gap < gw ? w : -1;
    };
    auto test_candidate = [&](int i, const M3& R, V3 p, int jc) {
        const float cx = M->cp[i][0], cy = M->cp[i][1], cz = M->cp[i][2], cr = M->cp[i][3];
        const V3 rc = p + mul(R, V3{cx, cy, cz});
        float gh = 0.0f;
        V3 n = V3{0, 0, 1};
        if (HF) ground_sample(P, st.rp.x + rc.x, st.rp.y + rc.y, gh, n);
        const float gap = (st.rp.z + rc.z - gh) * n.z - cr;
        const int s = gap < P.contact_offset ? claim_slot(gap) : -1;      // -1 with the slots full: this candidate (or the one it replaces) is dropped, counted, never silent
        if (s >= 0) {
            const V3 r = rc - n * cr;
            V3 t1 = V3{1.0f - n.x * n.x, -n.x * n.y, -n.x * n.z};
            t1 = t1 * (1.0f / sqrtf(dot(t1, t1)));
            const V3 t2 = cross(n, t1);
            sc.at(s, CF_RX) = r.x; sc.at(s, CF_RY) = r.y; sc.at(s, CF_RZ) = r.z;
            sc.at(s, CF_NX) = n.x; sc.at(s, CF_NY) = n.y; sc.at(s, CF_NZ) = n.z;
            sc.at(s, CF_T1X) = t1.x; sc.at(s, CF_T1Y) = t1.y; sc.at(s, CF_T1Z) = t1.z;
            sc.at(s, CF_T2X) = t2.x; sc.at(s, CF_T2Y) = t2.y; sc.at(s, CF_T2Z) = t2.z;
            sc.at(s, CF_GAP) = gap;
            sc.at(s, CF_JC) = (float)jc;
            sc.at(s, CF_BODY) = (float)M->cp_body[i];
            sc.at(s, CF_LN) = 0.0f; sc.at(s, CF_L1) = 0.0f; sc.at(s, CF_L2) = 0.0f;
        }
    };
    // Self-collision (DevParams::self_collide): the candidate sphere against the BASE's bounding box (self_box_c / self_box_h: the box, in the
    // root frame, around the candidates of API body 0 and their radii -- bodies fixed to the base, e.g. an arm's mount, are left out).  Signed distance = the largest of the three slab distances (exact over
    // a face, a lower bound next to an edge), normal = that face's outward normal.  The slot is marked by kSelfJc added to its link index:
    // the impulse acts on the link and, with the opposite sign, on the root; velocities at the contact are relative to the root.
    auto test_self = [&](int i, const M3& R, V3 p, int jc) {
        const float cx = M->cp[i][0], cy = M->cp[i][1], cz = M->cp[i][2], cr = M->cp[i][3];
        const V3 rc = p + mul(R, V3{cx, cy, cz});
        const V3 q = mulT(R0, rc) - V3{M->self_box_c[0], M->self_box_c[1], M->self_box_c[2]};
        const float dx = fabsf(q.x) - M->self_box_h[0] - cr, dy = fabsf(q.y) - M->self_box_h[1] - cr, dz = fabsf(q.z) - M->self_box_h[2] - cr;
        int best = 0;
        float gap = dx;
        if (dy > gap) { gap = dy; best = 1; }
        if (dz > gap) { gap = dz; best = 2; }
        const int s = gap < P.contact_offset ? claim_slot(gap) : -1;
        if (s >= 0) {
            const float sgn = (best == 0 ? q.x : best == 1 ? q.y : q.z) < 0.0f ? -1.0f : 1.0f;
            const V3 n = V3{R0.m[best], R0.m[3 + best], R0.m[6 + best]} * sgn;
            const V3 r = rc - n * cr;
            V3 t1 = V3{1.0f - n.x * n.x, -n.x * n.y, -n.x * n.z};
            t1 = t1 * (1.0f / sqrtf(dot(t1, t1)));
            const V3 t2 = cross(n, t1);
            sc.at(s, CF_RX) = r.x; sc.at(s, CF_RY) = r.y; sc.at(s, CF_RZ) = r.z;
            sc.at(s, CF_NX) = n.x; sc.at(s, CF_NY) = n.y; sc.at(s, CF_NZ) = n.z;
            sc.at(s, CF_T1X) = t1.x; sc.at(s, CF_T1Y) = t1.y; sc.at(s, CF_T1Z) = t1.z;
            sc.at(s, CF_T2X) = t2.x; sc.at(s, CF_T2Y) = t2.y; sc.at(s, CF_T2Z) = t2.z;
            sc.at(s, CF_GAP) = gap;
            sc.at(s, CF_JC) = (float)(jc + kSelfJc);
            sc.at(s, CF_BODY) = (float)M->cp_body[i];
            sc.at(s, CF_LN) = 0.0f; sc.at(s, CF_L1) = 0.0f; sc.at(s, CF_L2) = 0.0f;
        }
    };
    // plane ground: a link whose candidate bounding box clears the contact offset cannot touch (skips most loops)
    // heightfield: the same test against a coarse conservative bound of the terrain under the box (max height and min
    // normal z of the dilated block that holds the box centre, b2g_host_pack.h::build_hf_coarse): every candidate of the
    // link has gap = (z - gh) nz - r >= (zlow - Hmax) nzmin - r (1 - nzmin), r <= the smallest half extent
    auto may_touch = [&](const M3& R, V3 p, const float* c, const float* hx) -> bool {
        if (HF && !P.hfc) return true;
        const float zc = st.rp.z + p.z + R.m[6] * c[0] + R.m[7] * c[1] + R.m[8] * c[2];
        const float ext = fabsf(R.m[6]) * hx[0] + fabsf(R.m[7]) * hx[1] + fabsf(R.m[8]) * hx[2];
        if (!HF) return zc - ext < P.contact_offset;
        const float xc = st.rp.x + p.x + R.m[0] * c[0] + R.m[1] * c[1] + R.m[2] * c[2];
        const float yc = st.rp.y + p.y + R.m[3] * c[0] + R.m[4] * c[1] + R.m[5] * c[2];
        const float inv_hs = 1.0f / P.hf_hs;
        const float gx = fminf(fmaxf((xc - P.hf_ox) * inv_hs, 0.0f), (float)(P.hf_rows - 1));
        const float gy = fminf(fmaxf((yc - P.hf_oy) * inv_hs, 0.0f), (float)(P.hf_cols - 1));
        const int ci = ((int)gx) >> B2G_HFC_SHIFT, cj = ((int)gy) >> B2G_HFC_SHIFT;
        const float* b = P.hfc + 2 * ((size_t)ci * P.hfc_cols + cj);
        const float hmax = b[0], nzmin = b[1];
        const float d = zc - ext - hmax;
        const float rmin = fminf(hx[0], fminf(hx[1], hx[2]));
        const bool skip = d > 0.0f && d * nzmin - rmin * (1.0f - nzmin) >= P.contact_offset;
#if defined(B2G_HOST_EMU)
        __atomic_add_fetch(&emu_hfc_tests, 1, __ATOMIC_RELAXED);
        if (skip) __atomic_add_fetch(&emu_hfc_skips, 1, __ATOMIC_RELAXED);
#endif
        return !skip;
    };
    const bool self_col = SC && P.self_collide != 0;
    // a link whose candidate bounding sphere stays clear of the base box's bounding sphere by two contact offsets cannot produce a self
    // contact (the slab distance used above is at least 1 / sqrt(3) of the Euclidean one): skips the per-candidate loop, never a result
    auto may_self = [&](const M3& R, V3 p, const float* c, const float* hx) -> bool {
        if (M->self_box_h[0] < 0.0f) return false;      // the base has no candidates of its own: no box
        const V3 d = p + mul(R, V3{c[0], c[1], c[2]}) - mul(R0, V3{M->self_box_c[0], M->self_box_c[1], M->self_box_c[2]});
        const float rl = sqrtf(hx[0] * hx[0] + hx[1] * hx[1] + hx[2] * hx[2]);
        const float rb = sqrtf(M->self_box_h[0] * M->self_box_h[0] + M->self_box_h[1] * M->self_box_h[1] + M->self_box_h[2] * M->self_box_h[2]);
        const float reach = rl + rb + 2.0f * P.contact_offset;
        return dot(d, d) <= reach * reach;
    };
    if (ground || self_col) {
B2G_LINK_UNROLL
        for (int j = NL - 1; j >= 0; j--) {
            if (j < len) {
                const DevDof& D = M->dof[d0 + j];
                const int c0 = D.cp_start, cn = D.cp_count;
                if (ground && cn > 0 && may_touch(L[j].Rl, L[j].pl, D.cp_c, D.cp_h))
                    for (int i = c0; i < c0 + cn; i++) test_candidate(i, L[j].Rl, L[j].pl, j);
                if (self_col && (j >= 1 || distal) && cn > 0 && may_self(L[j].Rl, L[j].pl, D.cp_c, D.cp_h))      // not the link that hangs off the root directly (PhysX: no parent-child collision)
                    for (int i = c0; i < c0 + cn; i++) test_self(i, L[j].Rl, L[j].pl, j);
            }
        }
        if (!FIXED && ground) {
            const int c0 = M->root_cp_start[lane], cn = M->root_cp_count[lane];
            if (cn > 0 && may_touch(R0, V3{0, 0, 0}, M->root_cp_c, M->root_cp_h))
                for (int i = c0; i < c0 + cn; i++) test_candidate(i, R0, V3{0, 0, 0}, -1);
        }
    }

    // Segment variant: the contacts of a chain stay Gauss-Seidel ordered, distal piece first (the order of the whole-chain kernels):
    // a piece with a child starts its slots `off` steps late
    int off = 0;
    if (SEG) {
        const float nc = Grp<LANES>::bcast((float)ncon, down_lane);
        off = child >= 0 ? (int)nc : 0;
    }
    const int maxc = Grp<LANES>::warp_max(ncon);
    const int maxs = SEG ? Grp<LANES>::warp_max(off + ncon) : maxc;
    if (P.stats) contact_stats_add<LANES>(P.stats, lane, dr.live, ncon, ndrop);

    // local Delassus block of every active contact: response of the contact point to unit impulses along n, t1, t2
    // (one code instance for all links: the contact's link index is a run-time value, the chain loops are predicated)
    for (int s = 0; s < maxc; s++) {
        if (s < ncon) {
            const V3 r = V3{sc.at(s, CF_RX), sc.at(s, CF_RY), sc.at(s, CF_RZ)};
            const V3 dirs[3] = {V3{sc.at(s, CF_NX), sc.at(s, CF_NY), sc.at(s, CF_NZ)}, V3{sc.at(s, CF_T1X), sc.at(s, CF_T1Y), sc.at(s, CF_T1Z)},
                                V3{sc.at(s, CF_T2X), sc.at(s, CF_T2Y), sc.at(s, CF_T2Z)}};
            const int jcf = (int)sc.at(s, CF_JC);
            const bool selfc = SC && jcf >= kSelfJc / 2;
            const int jc = selfc ? jcf - kSelfJc : jcf;
            float A[9];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                float ud[NL], uda[3];
                const SV F = SV{cross(r, dirs[b]), dirs[b]};
                SV Pb = push_up(jc, F, ud, uda);
                if (selfc && !FIXED) Pb += F;      // the reaction -F acts on the root
                SV a = FIXED ? sv0() : -mul(inv, Pb);
                const SV a_root = a;
                if (SEG && distal) run_anc(a, uda);
#pragma unroll
                for (int k = 0; k < NL; k++) {
                    if (k <= jc) {
                        const float dq = (ud[k] - dot(L[k].U, a)) * L[k].Dinv;
                        a += L[k].S * dq;
                    }
                }
                if (selfc) a = a - a_root;         // response of the velocity RELATIVE to the root
                const V3 pv = a.v + cross(a.w, r);
#pragma unroll
                for (int c = 0; c < 3; c++) A[c * 3 + b] = dot(pv, dirs[c]);
            }
            // diagonal entries are kept as reciprocals: the sweeps multiply instead of dividing
            sc.at(s, CF_ANN) = 1.0f / A[0]; sc.at(s, CF_ANT1) = A[3]; sc.at(s, CF_ANT2) = A[6];
            sc.at(s, CF_AT1T1) = 1.0f / A[4]; sc.at(s, CF_AT1T2) = A[7]; sc.at(s, CF_AT2T2) = 1.0f / A[8];
            sc.at(s, CF_ATISO) = 1.0f / (fmaxf(A[4], A[8]) + fabsf(A[7]));      // scalar step length of the sliding branch: >= the largest eigenvalue of the tangential block
        }
    }

    // ---------------- projected relaxation: Gauss-Seidel along a lane, Jacobi across lanes ----------------
    block_align(P, 4);
    // Contacts of different chains couple only through the (heavy) root, so slot s of every lane is updated from the
    // same velocities; the root-level biases are summed with one shuffle butterfly and applied once.
    SV v0pos = v0n;
    float qdpos[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) qdpos[j] = qdn[j];
    const int nit = P.npos + P.nvel;
    const float mu = 0.5f * (P.mu_ground + dr.mu);   // PhysX default combine mode: average
    const float inv_h = 1.0f / h;

    for (int it = 0; it < nit; it++) {
        if (it == P.npos) {
            v0pos = v0n;
#pragma unroll
            for (int j = 0; j < NL; j++) qdpos[j] = qdn[j];
        }
        const bool with_bias = it < P.npos;
        for (int step = 0; step < maxs; step++) {
            SV Pb = sv0();
            float ud[NL], uda[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < NL; k++) ud[k] = 0.0f;
            const int s = SEG ? step - off : step;
            if (s >= 0 && s < ncon) {
                const V3 r = V3{sc.at(s, CF_RX), sc.at(s, CF_RY), sc.at(s, CF_RZ)};
                const V3 n = V3{sc.at(s, CF_NX), sc.at(s, CF_NY), sc.at(s, CF_NZ)};
                const V3 t1 = V3{sc.at(s, CF_T1X), sc.at(s, CF_T1Y), sc.at(s, CF_T1Z)};
                const V3 t2 = V3{sc.at(s, CF_T2X), sc.at(s, CF_T2Y), sc.at(s, CF_T2Z)};
                const int jcf = (int)sc.at(s, CF_JC);
                const bool selfc = SC && jcf >= kSelfJc / 2;
                const int jc = selfc ? jcf - kSelfJc : jcf;
                SV lv = distal ? vbase : v0n;
                if (selfc) lv = distal ? vbase - v0n : sv0();      // relative to the root
#pragma unroll
                for (int k = 0; k < NL; k++)
                    if (k <= jc) lv += L[k].S * qdn[k];
                const V3 pv = lv.v + cross(lv.w, r);
                float vn = dot(pv, n), vt1 = dot(pv, t1), vt2 = dot(pv, t2);
                float tgt = -sc.at(s, CF_GAP) * inv_h;
                tgt = fminf(tgt, P.max_depen);
                if (!with_bias) tgt = fminf(tgt, 0.0f);
                const float l0 = sc.at(s, CF_LN), l1o = sc.at(s, CF_L1), l2o = sc.at(s, CF_L2);
                float ln = fmaxf(l0 - (vn - tgt) * sc.at(s, CF_ANN), 0.0f);
                const float dn = ln - l0;
                vt1 += sc.at(s, CF_ANT1) * dn;
                vt2 += sc.at(s, CF_ANT2) * dn;
                float l1 = l1o - vt1 * sc.at(s, CF_AT1T1);
                const float vt2s = vt2 + sc.at(s, CF_AT1T2) * (l1 - l1o);
                float l2 = l2o - vt2s * sc.at(s, CF_AT2T2);
                const float lim_t = mu * ln;
                float mag = sqrtf(l1 * l1 + l2 * l2);
                if (mag > lim_t) {
                    // The sticking impulse leaves the cone: the contact slides.  Scaling the sticking impulse back would keep ITS direction
                    // (A_tt^-1 v_t), which is not opposite to the sliding velocity when the tangential Delassus block is anisotropic (a box
                    // corner, a foot on a leg).  A proximal step with a SCALAR step length followed by the radial projection has the Coulomb
                    // law as its fixed point: magnitude mu * lambda_n, opposite to the tangential velocity (block-on-a-slope known answer).
                    const float ia = sc.at(s, CF_ATISO);
                    l1 = l1o - vt1 * ia;
                    l2 = l2o - vt2 * ia;
                    mag = sqrtf(l1 * l1 + l2 * l2);
                    if (mag > lim_t) {
                        const float scl = (mag > 0.0f) ? lim_t / mag : 0.0f;
                        l1 *= scl; l2 *= scl;
                    }
                }
                sc.at(s, CF_LN) = ln; sc.at(s, CF_L1) = l1; sc.at(s, CF_L2) = l2;
                const V3 dir = n * dn + t1 * (l1 - l1o) + t2 * (l2 - l2o);
                Pb = push_up(jc, SV{cross(r, dir), dir}, ud, uda);
                if (selfc && !FIXED) Pb += SV{cross(r, dir), dir};
            }
            if (SEG) {      // the joint terms a distal piece found for its parent's links
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const float g = Grp<LANES>::bcast(uda[i], down_lane);
                    if (child >= 0) ud[i] += g;
                }
            }
            // all lanes' impulses reach the root together, then run down every chain
            SV a = sv0();
            if (!FIXED) {
                const SV Psum = (LANES > 1) ? grp_sum<LANES>(Pb) : Pb;
                a = -mul(inv, Psum);
                v0n += a;
            }
            auto run_own = [&]() {
#pragma unroll
                for (int k = 0; k < NL; k++) {
                    if (k < len) {
                        const float dq = (ud[k] - dot(L[k].U, a)) * L[k].Dinv;
                        a += L[k].S * dq;
                        qdn[k] += dq;
                    }
                }
            };
            if (!SEG) {
                run_own();
            } else {      // proximal pieces first; the velocity change of their last link is where the distal piece starts
                if (!distal) run_own();
                const SV aq = grp_bcast<LANES>(a, up_lane);
                if (distal) {
                    vbase += aq;
                    a = aq;
                    run_own();
                }
            }
        }
    }
    if (P.npos >= nit) {   // no velocity iterations: positions integrate with the final velocity
        v0pos = v0n;
#pragma unroll
        for (int j = 0; j < NL; j++) qdpos[j] = qdn[j];
    }

    // root velocity limits (PhysX clamps a body's linear / angular velocity: AssetOptions.max_linear_velocity / max_angular_velocity):
    // inactive in any sane state -- no arithmetic happens then -- and the guard that keeps a diverging controller from reaching inf / nan
    if (!FIXED) {
        auto clamp_norm = [](V3& v, float lim) {
            const float n2 = dot(v, v);
            if (lim > 0.0f && n2 > lim * lim) v = v * (lim / sqrtf(n2));
        };
        clamp_norm(v0n.v, P.max_lin_vel); clamp_norm(v0pos.v, P.max_lin_vel);
        clamp_norm(v0n.w, P.max_ang_vel); clamp_norm(v0pos.w, P.max_ang_vel);
    }

    // ---------------- integrate ----------------
#pragma unroll
    for (int j = 0; j < NL; j++) {
        if (j < len) {
            const DevDof& D = M->dof[d0 + j];
            float qv = qdn[j], qp = qdpos[j];
            if (D.vel_limit > 0.0f) {
                qv = fminf(fmaxf(qv, -D.vel_limit), D.vel_limit);
                qp = fminf(fmaxf(qp, -D.vel_limit), D.vel_limit);
            }
            st.q[j] += h * qp;
            st.qd[j] = qv;
            float f = 0.0f;
            if (D.drive_mode == B2G_DOF_MODE_POS) f = D.kp * dr.kp_of(1 + d0 + j) * (st.tgt[j] - st.q[j]) - D.kd * dr.kd_of(1 + d0 + j) * qv;
            else if (D.drive_mode == B2G_DOF_MODE_VEL) f = D.kd * dr.kd_of(1 + d0 + j) * (st.tgt[j] - qv);
            else if (D.drive_mode == B2G_DOF_MODE_EFFORT) f = st.act[j];
            if (D.effort > 0.0f) f = fminf(fmaxf(f, -D.effort), D.effort);
            st.frc[j] = f;
        }
    }
    if (!FIXED) {
        st.rp += v0pos.v * h;
        const V3 w = v0pos.w;
        const float wn = sqrtf(dot(w, w)), ang = wn * h;
        float dx = 0, dy = 0, dz = 0, dw = 1.0f;
        if (ang > 1e-12f) {
            float sh, ch;
#if defined(B2G_HOST_EMU)
            sh = sinf(0.5f * ang); ch = cosf(0.5f * ang);
#else
            sincosf(0.5f * ang, &sh, &ch);
#endif
            const float k = sh / wn;
            dx = w.x * k; dy = w.y * k; dz = w.z * k; dw = ch;
        }
        const float x2 = st.qx, y2 = st.qy, z2 = st.qz, w2 = st.qw;
        float nx = dw * x2 + dx * w2 + dy * z2 - dz * y2;
        float ny = dw * y2 - dx * z2 + dy * w2 + dz * x2;
        float nz = dw * z2 + dx * y2 - dy * x2 + dz * w2;
        float nw = dw * w2 - dx * x2 - dy * y2 - dz * z2;
        const float nn = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
        st.qx = nx * nn; st.qy = ny * nn; st.qz = nz * nn; st.qw = nw * nn;
        st.rv = v0n.v;
        st.rw = v0n.w;
    }

    // ---------------- net contact force per API body (last sub-step only) ----------------
    if (last) {
        const int nb3 = M->n_bodies * 3;
        for (int i = lane; i < nb3; i += LANES) bf[i] = 0.0f;
        Grp<LANES>::sync();
        const float ih = 1.0f / h;
        for (int owner = 0; owner < LANES; owner++) {
            if (lane == owner) {
                for (int s = 0; s < ncon; s++) {
                    const float ln = sc.at(s, CF_LN) * ih, l1 = sc.at(s, CF_L1) * ih, l2 = sc.at(s, CF_L2) * ih;
                    const int b = (int)sc.at(s, CF_BODY);
                    const float fx = sc.at(s, CF_NX) * ln + sc.at(s, CF_T1X) * l1 + sc.at(s, CF_T2X) * l2;
                    const float fy = sc.at(s, CF_NY) * ln + sc.at(s, CF_T1Y) * l1 + sc.at(s, CF_T2Y) * l2;
                    const float fz = sc.at(s, CF_NZ) * ln + sc.at(s, CF_T1Z) * l1 + sc.at(s, CF_T2Z) * l2;
                    bf[b * 3 + 0] += fx; bf[b * 3 + 1] += fy; bf[b * 3 + 2] += fz;
                    if (SC && (int)sc.at(s, CF_JC) >= kSelfJc / 2) { bf[0] -= fx; bf[1] -= fy; bf[2] -= fz; }      // self contact: the reaction on the root body (API body 0)
                }
            }
            Grp<LANES>::sync();
        }
    }
}

}  // namespace b2g
