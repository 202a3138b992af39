// b2g_policy.cu -- fused actor-critic MLP forward for the rollout ("next"): the network rl_games
// builds from cfg/train/AnymalPPO.yaml:10-33 (shared trunk, units [256,128,64], ELU, mu head + value head, input
// normalisation cfg/train/AnymalPPO.yaml:45) evaluated for every environment once per VecTask.step
// (isaacgymenvs/utils/rlgames_utils.py:242-262 is the env side of that loop).
//
// sm_100a only.  One CTA = one tile of 128 observation rows:
//   * all weights (bf16, UMMA canonical K-major no-swizzle layout, packed once by k_pack_layer) are brought into shared memory
//     with four 1-D bulk async copies (cp.async.bulk, the TMA engine) completing on an mbarrier and stay resident for every
//     tile the CTA processes;
//   * each layer is a chain of tcgen05.mma (M=128, N=layer width, K=16 per instruction) issued by ONE thread, A and B from
//     shared memory, fp32 accumulators in tensor memory (TMEM); completion is signalled by tcgen05.commit on an mbarrier;
//   * the epilogue reads the accumulators with tcgen05.ld (one TMEM lane = one row; the row's columns are split between two
//     threads, warps w and w+4), adds the bias, applies ELU, converts to bf16 and writes the row straight into the A-operand
//     layout of the next layer in shared memory, so activations never leave the SM; the head writes mu (n_actions) and
//     value to global memory.  The kernel is bound by this elementwise epilogue (K is tiny), not by the tensor cores.
// Observation normalisation ((x-mean)/sqrt(var+eps), clamp) is fused into the first operand load.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <new>

#include "b200gym.h"

namespace b2g {
int fail_msg(int code, const char* msg);   // b200gym.cu: sets b2g_last_error()
}

namespace {

constexpr int kTileM = 128;       // rows per tile = UMMA M = TMEM lanes
constexpr int kThreads = 256;     // 8 warps: warp w reads TMEM lanes 32(w%4)..+31; warps w and w+4 split a row's columns
constexpr int kHeadN = 16;        // mu (n_actions) + value, padded to the minimum UMMA N for M=128
constexpr int kTmemCols = 512;

struct PolicyDims {
    int k0, k0p;          // observation width, padded to a multiple of 16
    int n1, n2, n3;       // hidden widths (multiples of 16, <= 256)
    int n_act;            // mu head width (n_act + 1 <= 16)
    // byte offsets into dynamic shared memory
    int off_w1, off_w2, off_w3, off_wh, off_a, off_bias, off_norm, off_bar, total;
    int w_bytes[4];
};

#define CUDA_TRY_P(expr)                                                                   \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            char b_[384];                                                                  \
            snprintf(b_, sizeof(b_), "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return b2g::fail_msg(B2G_ERR_CUDA, b_);                                        \
        }                                                                                  \
    } while (0)

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a descriptor bug must surface as a trap (launch failure), never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
// accumulator-ready wait of the epilogue warps: ONE lane per warp polls (with a short back-off), the others park at the warp
// barrier -- 256 threads spinning on try_wait starve the single thread that is issuing copies and MMAs next to them
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        uint32_t done = 0;
        for (uint32_t spin = 0; spin < (1u << 22); spin++) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
            if (done) break;
            __nanosleep(40);
        }
        if (!done) __trap();
    }
    __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: 8x(16 B) core matrices; LBO = byte distance between the two core
// matrices of one K=16 slice, SBO = byte distance between consecutive 8-row groups; version 1 (Blackwell), base offset 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 (16) accumulator columns of this thread's TMEM lane -> registers.  The load and its wait are ONE asm statement so that
// neither the compiler nor the assembler can place a use of r[] between them.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// ELU, branch-free: max(x, exp(min(x,0)) - 1)  (x > 0: max(x, 0) = x; x <= 0: e^x - 1 >= x).  ex2.approx.ftz is a single MUFU
// (2^-22 relative); the non-ftz __expf expands to a denormal range check around it.
__device__ __forceinline__ float elu(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 0.f) * 1.4426950408889634f));
    return fmaxf(x, e - 1.f);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// one layer's MMAs: D[128 x n] (TMEM, column tcol) = A[128 x k] (smem a_addr) . W[n x k]^T (smem w_addr); single thread
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, int tcol, uint32_t a_addr, uint32_t w_addr, int n, int k, uint32_t bar) {
    const uint32_t a_lbo = kTileM * 16, b_lbo = (uint32_t)n * 16, idesc = instr_desc(n);
    for (int k16 = 0; k16 < k / 16; k16++) {
        uint64_t ad = smem_desc(a_addr + (uint32_t)k16 * 2u * a_lbo, a_lbo, 128);
        uint64_t bd = smem_desc(w_addr + (uint32_t)k16 * 2u * b_lbo, b_lbo, 128);
        umma_bf16(tmem_base + (uint32_t)tcol, ad, bd, idesc, k16 > 0 ? 1u : 0u);
    }
    umma_commit(bar);
}

// 16 accumulator columns -> +bias -> ELU -> bf16 -> two 16-byte k-slabs of the next layer's A operand
__device__ __forceinline__ void store_slabs16(const uint32_t* r, const float* bias, uint8_t* a_row, int c0) {
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 b = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
        v[4 * q + 0] = elu(__uint_as_float(r[4 * q + 0]) + b.x);
        v[4 * q + 1] = elu(__uint_as_float(r[4 * q + 1]) + b.y);
        v[4 * q + 2] = elu(__uint_as_float(r[4 * q + 2]) + b.z);
        v[4 * q + 3] = elu(__uint_as_float(r[4 * q + 3]) + b.w);
    }
    *reinterpret_cast<uint4*>(a_row + (size_t)(c0 / 8) * (kTileM * 16)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(a_row + (size_t)(c0 / 8 + 1) * (kTileM * 16)) =
        make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}

// hidden-layer epilogue: thread = row; acc -> +bias -> ELU -> bf16 -> next layer's A operand (k-slab s at s*2048 + row*16),
// 32 columns per tcgen05.ld, a 16-column tail if n % 32
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_row, int tcol, int n, const float* bias, uint8_t* a_smem, int row, int half) {
    uint8_t* a_row = a_smem + row * 16;
    const int n32 = n / 32;
    uint32_t r[32];
    for (int c = half; c < n32; c += 2) {        // the two threads of a row take alternate 32-column chunks
        tmem_ld32(tmem_row + (uint32_t)(tcol + c * 32), r);
        store_slabs16(r, bias, a_row, c * 32);
        store_slabs16(r + 16, bias, a_row, c * 32 + 16);
    }
    if ((n % 32) && (n32 & 1) == half) {
        tmem_ld16(tmem_row + (uint32_t)(tcol + n32 * 32), r);
        store_slabs16(r, bias, a_row, n32 * 32);
    }
}

// observation tile: 128 rows x k0 floats, contiguous in global memory.  Vector path (k0 % 4 == 0): thread t owns float4
// number j*128+t, j < kObsVec -- coalesced, all loads issued before any use, so the next tile's rows can be fetched under the
// current tile's epilogues.
constexpr int kObsVec = 8;         // <= 8 float4 per thread (256 threads): k0 <= 64 on the vector path

__device__ __forceinline__ void obs_fetch(const float* __restrict__ obs, int row_base, int n_rows, int k0, int tid, float4* pre) {
    const int per_row = k0 >> 2, total = kTileM * per_row;
    const float inv = 1.0f / (float)per_row;     // f < 2^13, per_row <= 16: (f + 0.5) * inv truncates to f / per_row exactly
#pragma unroll
    for (int j = 0; j < kObsVec; j++) {
        const int f = j * kThreads + tid;
        const int r = __float2int_rz(((float)f + 0.5f) * inv);
        pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < total && row_base + r < n_rows) pre[j] = __ldg(reinterpret_cast<const float4*>(obs + (size_t)row_base * k0) + f);
    }
}

__device__ __forceinline__ void obs_commit(const float4* pre, const float* norm_s, float clip, int k0, int tid, uint8_t* a_smem) {
    const int per_row = k0 >> 2, total = kTileM * per_row;
    const float inv = 1.0f / (float)per_row;
#pragma unroll
    for (int j = 0; j < kObsVec; j++) {
        const int f = j * kThreads + tid;
        if (f < total) {
            const int r = __float2int_rz(((float)f + 0.5f) * inv), k = (f - r * per_row) * 4;
            const float4 m = *reinterpret_cast<const float4*>(norm_s + k);
            const float4 rs = *reinterpret_cast<const float4*>(norm_s + k0 + k);
            const float x0 = fminf(fmaxf((pre[j].x - m.x) * rs.x, -clip), clip), x1 = fminf(fmaxf((pre[j].y - m.y) * rs.y, -clip), clip);
            const float x2 = fminf(fmaxf((pre[j].z - m.z) * rs.z, -clip), clip), x3 = fminf(fmaxf((pre[j].w - m.w) * rs.w, -clip), clip);
            *reinterpret_cast<uint2*>(a_smem + (size_t)(k >> 3) * (kTileM * 16) + r * 16 + (k & 7) * 2) =
                make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// fp32 row-major W[rows x cols] (nn.Linear layout: out x in) -> bf16 canonical K-major operand with np rows, at row row0
__global__ void k_pack_layer(const float* __restrict__ W, const float* __restrict__ b, int rows, int cols, int row0, int np,
                             __nv_bfloat16* __restrict__ Wp, float* __restrict__ bp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < rows * cols) {
        const int n = idx / cols, k = idx % cols;
        Wp[(size_t)(k / 8) * ((size_t)np * 8) + (size_t)(row0 + n) * 8 + (k % 8)] = __float2bfloat16_rn(W[idx]);
    }
    if (idx < rows) bp[row0 + idx] = b[idx];
}

__global__ void k_pack_norm(const float* __restrict__ mean, const float* __restrict__ var, float eps, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i] = mean ? mean[i] : 0.f;
        out[n + i] = var ? rsqrtf(var[i] + eps) : 1.f;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
k_policy_forward(PolicyDims d, const uint8_t* __restrict__ wpack, const float* __restrict__ bias_g, const float* __restrict__ norm_g,
                 float clip, const float* __restrict__ obs, int n_rows, float* __restrict__ mu, float* __restrict__ value, int obs_aligned) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTileM - 1), half = tid >> 7;
    float* bias_s = reinterpret_cast<float*>(smem + d.off_bias);
    float* norm_s = reinterpret_cast<float*>(smem + d.off_norm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + d.off_bar);      // [0..3] weights of layer l landed, [4] MMA chain done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const uint32_t bar_mma = smem_u32(&bars[4]);
    uint8_t* a_smem = smem + d.off_a;
    const uint32_t a_addr = smem_u32(a_smem);
    const int offs[4] = {d.off_w1, d.off_w2, d.off_w3, d.off_wh};
    const bool vec = obs_aligned && (d.k0 & 3) == 0 && kTileM * (d.k0 >> 2) <= kObsVec * kThreads;
    const int n_tiles = (n_rows + kTileM - 1) / kTileM;

    // first tile's observations: in flight while the weights stream in
    float4 pre[kObsVec];
    if (vec && (int)blockIdx.x < n_tiles) obs_fetch(obs, blockIdx.x * kTileM, n_rows, d.k0, tid, pre);

    if (tid == 0) {
        for (int l = 0; l < 5; l++) mbar_init(smem_u32(&bars[l]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
        // weights: global (already in operand layout) -> shared, one bulk copy and one barrier per layer, first layer first
        size_t g = 0;
        for (int l = 0; l < 4; l++) {
            mbar_expect_tx(smem_u32(&bars[l]), (uint32_t)d.w_bytes[l]);
            bulk_g2s(smem_u32(smem + offs[l]), wpack + g, (uint32_t)d.w_bytes[l], smem_u32(&bars[l]));
            g += (size_t)d.w_bytes[l];
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int n_bias = d.n1 + d.n2 + d.n3 + kHeadN;
    for (int i = tid; i < n_bias; i += kThreads) bias_s[i] = bias_g[i];
    for (int i = tid; i < 2 * d.k0; i += kThreads) norm_s[i] = norm_g[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int tcol[4] = {0, d.n1, d.n1 + d.n2, d.n1 + d.n2 + d.n3};
    const int width[4] = {d.n1, d.n2, d.n3, kHeadN};
    const int depth[4] = {d.k0p, d.n1, d.n2, d.n3};
    const float* bias_l[4] = {bias_s, bias_s + d.n1, bias_s + d.n1 + d.n2, bias_s + d.n1 + d.n2 + d.n3};
    uint32_t phase = 0;
    bool first = true;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row_base = tile * kTileM;
        // ---- observations -> normalise -> bf16 A operand ----
        if (vec) {
            obs_commit(pre, norm_s, clip, d.k0, tid, a_smem);
            if (d.k0p != d.k0)      // zero the K padding (the buffer holds the previous tile's activations)
                for (int idx = tid; idx < kTileM * (d.k0p - d.k0); idx += kThreads) {
                    const int r = idx / (d.k0p - d.k0), k = d.k0 + idx % (d.k0p - d.k0);
                    *reinterpret_cast<__nv_bfloat16*>(a_smem + (size_t)(k >> 3) * (kTileM * 16) + r * 16 + (k & 7) * 2) = __float2bfloat16_rn(0.f);
                }
            // next tile's rows: fetched under this tile's four layers
            if (tile + (int)gridDim.x < n_tiles) obs_fetch(obs, (tile + gridDim.x) * kTileM, n_rows, d.k0, tid, pre);
        } else {
            for (int idx = tid; idx < kTileM * d.k0p; idx += kThreads) {
                const int r = idx / d.k0p, k = idx - r * d.k0p;
                float x = 0.f;
                if (k < d.k0 && row_base + r < n_rows) {
                    x = (obs[(size_t)(row_base + r) * d.k0 + k] - norm_s[k]) * norm_s[d.k0 + k];
                    x = fminf(fmaxf(x, -clip), clip);
                }
                *reinterpret_cast<__nv_bfloat16*>(a_smem + (size_t)(k >> 3) * (kTileM * 16) + r * 16 + (k & 7) * 2) = __float2bfloat16_rn(x);
            }
        }
#pragma unroll
        for (int l = 0; l < 4; l++) {
            // the A operand was written through the generic proxy: make it visible to the tensor core's async proxy
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                if (first) mbar_wait(smem_u32(&bars[l]), 0);
                tc_fence_after();
                issue_layer(tmem_base, tcol[l], a_addr, smem_u32(smem + offs[l]), width[l], depth[l], bar_mma);
            }
            __syncwarp();      // lanes 1-31 of the issuing warp park here instead of spinning on the barrier next to lane 0
            mbar_wait_warp(bar_mma, phase);
            phase ^= 1;
            tc_fence_after();
            if (l < 3) {
                epilogue_hidden(tmem_row, tcol[l], width[l], bias_l[l], a_smem, row, half);
            } else if (half == 0) {
                // heads: mu (n_act columns) and value (column n_act)
                uint32_t r[16];
                tmem_ld16(tmem_row + (uint32_t)tcol[3], r);
                const int grow = row_base + row;
                if (grow < n_rows) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float o = __uint_as_float(r[i]) + bias_l[3][i];
                        if (i < d.n_act) mu[(size_t)grow * d.n_act + i] = o;
                        else if (i == d.n_act) value[grow] = o;
                    }
                }
            }
        }
        first = false;
        tc_fence_before();
        __syncthreads();    // the A buffer and the TMEM columns are free for the next tile
    }
    // a CTA without a tile still has the weight copies in flight: they must land before the CTA exits
    if (tid == 0 && first)
        for (int l = 0; l < 4; l++) mbar_wait(smem_u32(&bars[l]), 0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------
// wide variant: networks whose weights do not fit in shared memory next to one activation tile (the rough-terrain policies,
// units [512,256,128], 188 / 204 observations: cfg/train/AnymalTerrainPPO.yaml, UsefulHoundPPO.yaml).  Same tile, same MMA /
// epilogue scheme; what changes is where the operands live:
//   * the weights STREAM: one chunk = one K=16 slice of at most 256 output rows (<= 8 KB, two bulk copies), brought into a ring
//     of kStages shared-memory stages by the TMA engine; full[s] counts the bytes in, empty[s] is armed by tcgen05.commit when the
//     MMA that read stage s has retired.  One elected thread issues both the copies and the MMAs; the copies run kStages-1 chunks
//     ahead in a fixed per-tile order (a table built by the host), across layer boundaries -- the next layer's first slices arrive
//     under the current layer's epilogue;
//   * two activation buffers ping-pong: Y = observations, then layer-2 output; X = layer-1 output, then layer-3 output;
//   * a first layer wider than 256 is produced in ONE accumulator (all 512 tensor-memory columns) but handed to layer 2 in two
//     halves through the 64 KB X buffer: epilogue of columns 0..n1/2 -> layer-2 MMAs over the first n1/2 of K -> epilogue of the
//     other half -> layer-2 MMAs over the rest (accumulating).  Layer 2's accumulator reuses the columns the first half freed.
struct WideDims {
    int k0, k0p, n1, n2, n3, n_act, head_n;
    int parts1;               // 1 or 2: column halves of layer 1
    int off_x, off_y, off_ring, off_bias, off_norm, off_bar, total;
    int stages, stage_bytes, n_chunks;
};

constexpr int kMaxStages = 12;


// The ring is driven by ONE thread, so its bookkeeping is kept to adds and compares: no division, no 64-bit arithmetic (a
// first version that derived stage and chunk from running counters with / and % spent ~0.5 us per chunk in that scalar code).
// Chunk order of a tile = the order k_policy_forward_wide consumes them: layer 1 (K slices outer, column halves inner), layer 2,
// layer 3, heads.  wpack holds the four matrices back to back in the canonical operand layout (k_pack_layer).
struct WideRing {
    uint32_t ring_addr, full0, empty0;     // shared addresses: ring base, full[0], empty[0] (consecutive 8-byte barriers)
    uint32_t stages, stage_bytes;
    const uint8_t* wpack;
    // per segment (layer): chunks, byte offset of the matrix, bytes per 8-column slab of the whole matrix, bytes per slab of a chunk
    uint32_t seg_cnt[4], seg_base[4], seg_pitch[4], seg_bytes[4];
    uint32_t parts1;
    // loader cursor
    uint32_t tiles_left, ld_seg, ld_idx, ld_stage, ld_lap;
    // consumer cursor
    uint32_t cs_stage, cs_lap;
    __device__ __forceinline__ void init(const WideDims& d, uint32_t my_tiles) {
        const uint32_t h1 = d.n1 / d.parts1;
        parts1 = d.parts1;
        seg_cnt[0] = (d.k0p / 16) * d.parts1; seg_base[0] = 0; seg_pitch[0] = d.n1 * 16; seg_bytes[0] = h1 * 16;
        seg_cnt[1] = d.n1 / 16; seg_base[1] = d.n1 * d.k0p * 2; seg_pitch[1] = d.n2 * 16; seg_bytes[1] = d.n2 * 16;
        seg_cnt[2] = d.n2 / 16; seg_base[2] = seg_base[1] + d.n2 * d.n1 * 2; seg_pitch[2] = d.n3 * 16; seg_bytes[2] = d.n3 * 16;
        seg_cnt[3] = d.n3 / 16; seg_base[3] = seg_base[2] + d.n3 * d.n2 * 2; seg_pitch[3] = d.head_n * 16; seg_bytes[3] = d.head_n * 16;
        tiles_left = my_tiles; ld_seg = 0; ld_idx = 0; ld_stage = 0; ld_lap = 0; cs_stage = 0; cs_lap = 0;
    }
    __device__ __forceinline__ void load_next() {
        if (tiles_left == 0) return;
        if (ld_lap > 0) mbar_wait(empty0 + 8u * ld_stage, (ld_lap - 1) & 1);
        uint32_t cnt, base, pitch, bytes;
        // explicit selection keeps the four-entry tables in registers
        if (ld_seg == 0) { cnt = seg_cnt[0]; base = seg_base[0]; pitch = seg_pitch[0]; bytes = seg_bytes[0]; }
        else if (ld_seg == 1) { cnt = seg_cnt[1]; base = seg_base[1]; pitch = seg_pitch[1]; bytes = seg_bytes[1]; }
        else if (ld_seg == 2) { cnt = seg_cnt[2]; base = seg_base[2]; pitch = seg_pitch[2]; bytes = seg_bytes[2]; }
        else { cnt = seg_cnt[3]; base = seg_base[3]; pitch = seg_pitch[3]; bytes = seg_bytes[3]; }
        uint32_t k16 = ld_idx, r0 = 0;
        if (ld_seg == 0 && parts1 == 2) { k16 = ld_idx >> 1; r0 = (ld_idx & 1) * bytes; }
        const uint32_t off0 = base + 2u * k16 * pitch + r0;
        const uint32_t dst = ring_addr + ld_stage * stage_bytes, bar = full0 + 8u * ld_stage;
        mbar_expect_tx(bar, 2u * bytes);
        bulk_g2s(dst, wpack + off0, bytes, bar);
        bulk_g2s(dst + bytes, wpack + off0 + pitch, bytes, bar);
        if (++ld_idx == cnt) {
            ld_idx = 0;
            if (++ld_seg == 4) { ld_seg = 0; tiles_left--; }
        }
        if (++ld_stage == stages) { ld_stage = 0; ld_lap++; }
    }
    // one MMA: D[128 x n] (+)= A[128 x 16] (shared, k-slab pair at a_slice) . chunk^T
    __device__ __forceinline__ void consume(uint32_t tmem_d, uint32_t a_slice, int n, uint32_t accumulate) {
        mbar_wait(full0 + 8u * cs_stage, cs_lap & 1);
        tc_fence_after();
        const uint64_t ad = smem_desc(a_slice, kTileM * 16, 128);
        const uint64_t bd = smem_desc(ring_addr + cs_stage * stage_bytes, (uint32_t)n * 16, 128);
        umma_bf16(tmem_d, ad, bd, instr_desc(n), accumulate);
        umma_commit(empty0 + 8u * cs_stage);
        if (++cs_stage == stages) { cs_stage = 0; cs_lap++; }
    }
};

// observation tile -> normalise -> bf16 A operand, any width (k0 % 4 == 0 and a 16-byte aligned base take float4 loads).
// Four 4-column groups per thread and pass, all loads issued before the first use.
__device__ __forceinline__ void obs_stage_wide(const float* __restrict__ obs, int row_base, int n_rows, int k0, int k0p, const float* norm_s,
                                               float clip, int tid, uint8_t* a_smem, bool vec) {
    const int q = k0p >> 2;                         // 4-column groups per (padded) row
    const float inv = 1.0f / (float)q;              // idx < 2^13, q <= 64: (idx + 0.5) * inv truncates to idx / q exactly
    const int total = kTileM * q;
    for (int base = 0; base < total; base += 4 * kThreads) {
        float x[4][4];
        int rr[4], kk[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base + u * kThreads + tid;
            const int r = __float2int_rz(((float)idx + 0.5f) * inv), k = (idx - r * q) * 4;
            rr[u] = r; kk[u] = k;
#pragma unroll
            for (int j = 0; j < 4; j++) x[u][j] = 0.f;
            if (idx < total && row_base + r < n_rows) {
                const float* src = obs + (size_t)(row_base + r) * k0 + k;
                if (vec && k + 3 < k0) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                    x[u][0] = v.x; x[u][1] = v.y; x[u][2] = v.z; x[u][3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (k + j < k0) x[u][j] = __ldg(src + j);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int idx = base + u * kThreads + tid, r = rr[u], k = kk[u];
            if (idx < total) {
                float y[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    y[j] = (k + j < k0 && row_base + r < n_rows) ? fminf(fmaxf((x[u][j] - norm_s[k + j]) * norm_s[k0 + k + j], -clip), clip) : 0.f;
                *reinterpret_cast<uint2*>(a_smem + (size_t)(k >> 3) * (kTileM * 16) + r * 16 + (k & 7) * 2) =
                    make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads + 32, 1)
k_policy_forward_wide(WideDims d, const uint8_t* __restrict__ wpack, const float* __restrict__ bias_g,
                      const float* __restrict__ norm_g, float clip, const float* __restrict__ obs, int n_rows, float* __restrict__ mu,
                      float* __restrict__ value, int obs_aligned) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTileM - 1), half = tid >> 7;
    float* bias_s = reinterpret_cast<float*>(smem + d.off_bias);
    float* norm_s = reinterpret_cast<float*>(smem + d.off_norm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + d.off_bar);      // full[kMaxStages], empty[kMaxStages], mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
    const uint32_t bar_mma = smem_u32(&bars[2 * kMaxStages]);
    uint8_t* x_smem = smem + d.off_x;
    uint8_t* y_smem = smem + d.off_y;
    const uint32_t x_addr = smem_u32(x_smem), y_addr = smem_u32(y_smem);
    const int n_tiles = (n_rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const bool vec = obs_aligned && (d.k0 & 3) == 0;

    WideRing ring;
    ring.ring_addr = smem_u32(smem + d.off_ring);
    ring.full0 = smem_u32(&bars[0]);
    ring.empty0 = smem_u32(&bars[kMaxStages]);
    ring.stages = (uint32_t)d.stages; ring.stage_bytes = (uint32_t)d.stage_bytes;
    ring.wpack = wpack;
    ring.init(d, (uint32_t)my_tiles);

    if (tid == 0) {
        for (int l = 0; l < 2 * kMaxStages + 1; l++) mbar_init(smem_u32(&bars[l]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int n_bias = d.n1 + d.n2 + d.n3 + d.head_n;
    for (int i = tid; i < n_bias; i += kThreads) bias_s[i] = bias_g[i];
    for (int i = tid; i < 2 * d.k0; i += kThreads) norm_s[i] = norm_g[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kThreads / 32) {
        // producer warp: lane 0 streams every chunk of every tile of this CTA through the ring, paced only by the empty barriers
        if ((tid & 31) == 0)
            while (ring.tiles_left) ring.load_next();
        __syncwarp();
        __syncthreads();      // teardown barrier (the epilogue warps synchronise among themselves on named barrier 1 until then)
        return;
    }
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float* bias1 = bias_s;
    const float* bias2 = bias_s + d.n1;
    const float* bias3 = bias_s + d.n1 + d.n2;
    const float* biash = bias_s + d.n1 + d.n2 + d.n3;
    const int h1 = d.n1 / d.parts1;       // columns of layer 1 handed to layer 2 at a time
    uint32_t phase = 0;
    // every phase below: [A operand complete] -> barrier -> thread 0 issues the MMAs of the phase -> all wait for the accumulator
    auto phase_begin = [&]() {
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    auto phase_wait = [&]() {
        mbar_wait_warp(bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
    };

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row_base = tile * kTileM;
        obs_stage_wide(obs, row_base, n_rows, d.k0, d.k0p, norm_s, clip, tid, y_smem, vec);
        // ---- layer 1: A = Y (observations), D = columns [0, n1) ----
        phase_begin();
        if (tid == 0) {
            tc_fence_after();
            for (int k16 = 0; k16 < d.k0p / 16; k16++)
                for (int p = 0; p < d.parts1; p++)
                    ring.consume(tmem_base + (uint32_t)(p * h1), y_addr + (uint32_t)k16 * 2u * (kTileM * 16), h1, k16 > 0 ? 1u : 0u);
            umma_commit(bar_mma);
        }
        __syncwarp();      // lanes 1-31 of the issuing warp park here instead of spinning on the barrier next to lane 0
        phase_wait();
        // ---- layer 2 in parts1 passes: epilogue of one column half of layer 1 -> X, then the MMAs over that part of K ----
        for (int p = 0; p < d.parts1; p++) {
            epilogue_hidden(tmem_row, p * h1, h1, bias1 + p * h1, x_smem, row, half);
            phase_begin();
            if (tid == 0) {
                tc_fence_after();
                // the second pass accumulates; the accumulator lives in columns [0, n2): freed by the first half's epilogue
                for (int k16 = 0; k16 < h1 / 16; k16++)
                    ring.consume(tmem_base, x_addr + (uint32_t)k16 * 2u * (kTileM * 16), d.n2, (p > 0 || k16 > 0) ? 1u : 0u);
                umma_commit(bar_mma);
            }
            __syncwarp();
            phase_wait();
        }
        // ---- layer 3: A = Y (layer-2 output), D = columns [256, 256 + n3) ----
        epilogue_hidden(tmem_row, 0, d.n2, bias2, y_smem, row, half);
        phase_begin();
        if (tid == 0) {
            tc_fence_after();
            for (int k16 = 0; k16 < d.n2 / 16; k16++)
                ring.consume(tmem_base + 256u, y_addr + (uint32_t)k16 * 2u * (kTileM * 16), d.n3, k16 > 0 ? 1u : 0u);
            umma_commit(bar_mma);
        }
        __syncwarp();      // lanes 1-31 of the issuing warp park here instead of spinning on the barrier next to lane 0
        phase_wait();
        // ---- heads: A = X (layer-3 output), D = columns [0, head_n) ----
        epilogue_hidden(tmem_row, 256, d.n3, bias3, x_smem, row, half);
        phase_begin();
        if (tid == 0) {
            tc_fence_after();
            for (int k16 = 0; k16 < d.n3 / 16; k16++)
                ring.consume(tmem_base, x_addr + (uint32_t)k16 * 2u * (kTileM * 16), d.head_n, k16 > 0 ? 1u : 0u);
            umma_commit(bar_mma);
        }
        __syncwarp();      // lanes 1-31 of the issuing warp park here instead of spinning on the barrier next to lane 0
        phase_wait();
        if (half == 0) {
            const int grow = row_base + row;
            for (int c = 0; c < d.head_n; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_row + (uint32_t)c, r);
                if (grow < n_rows) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float o = __uint_as_float(r[i]) + biash[c + i];
                        if (c + i < d.n_act) mu[(size_t)grow * d.n_act + c + i] = o;
                        else if (c + i == d.n_act) value[grow] = o;
                    }
                }
            }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 256;" ::: "memory");    // both activation buffers and every accumulator column are free for the next tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct b2g_policy {
    int device = 0;
    bool wide = false;            // streamed-weights kernel (k_policy_forward_wide)
    int head_n = kHeadN;          // mu + value columns, padded to a multiple of 16
    WideDims wd{};
    PolicyDims d{};
    uint8_t* wpack = nullptr;     // the four packed weight matrices, back to back
    float* bias = nullptr;        // n1 + n2 + n3 + 16
    float* norm = nullptr;        // mean[k0], rstd[k0]
    float clip = 5.0f;
    int n_sm = 148;
    int64_t launches = 0;
};

extern "C" {

int b2g_policy_create(int device, int n_obs, const int* units, int n_actions, b2g_policy** out) {
    if (!out || !units) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_create: null argument");
    *out = nullptr;
    if (n_obs < 1 || n_actions < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_create: n_obs and n_actions must be positive");
    for (int i = 0; i < 3; i++) {
        const int lim = i == 0 ? 512 : 256;
        if (units[i] < 16 || units[i] > lim || units[i] % 16 || (units[i] > 256 && units[i] % 32))
            return b2g::fail_msg(B2G_ERR_UNSUPPORTED,
                                 "b2g_policy_create: three hidden layers, widths multiples of 16: first in [16,512] (multiple of 32 above 256), others in [16,256]");
    }
    if (n_actions + 1 > 32) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_policy_create: n_actions + 1 must be <= 32");
    if (n_obs > 256) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_policy_create: n_obs must be <= 256");
    CUDA_TRY_P(cudaSetDevice(device));
    b2g_policy* p = new (std::nothrow) b2g_policy();
    if (!p) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_create: out of memory");
    p->device = device;
    p->head_n = round_up(n_actions + 1, 16);
    PolicyDims& d = p->d;
    d.k0 = n_obs;
    d.k0p = round_up(n_obs, 16);
    d.n1 = units[0];
    d.n2 = units[1];
    d.n3 = units[2];
    d.n_act = n_actions;
    d.w_bytes[0] = d.n1 * d.k0p * 2;
    d.w_bytes[1] = d.n2 * d.n1 * 2;
    d.w_bytes[2] = d.n3 * d.n2 * 2;
    d.w_bytes[3] = p->head_n * d.n3 * 2;
    int off = 0;
    d.off_w1 = off; off += round_up(d.w_bytes[0], 128);
    d.off_w2 = off; off += round_up(d.w_bytes[1], 128);
    d.off_w3 = off; off += round_up(d.w_bytes[2], 128);
    d.off_wh = off; off += round_up(d.w_bytes[3], 128);
    int kmax = d.k0p;
    kmax = d.n1 > kmax ? d.n1 : kmax;
    kmax = d.n2 > kmax ? d.n2 : kmax;
    kmax = d.n3 > kmax ? d.n3 : kmax;
    d.off_a = off; off += kTileM * kmax * 2;
    d.off_bias = off; off += round_up((d.n1 + d.n2 + d.n3 + p->head_n) * 4, 16);
    d.off_norm = off; off += round_up(2 * d.k0 * 4, 16);
    d.off_bar = off; off += 64;
    d.total = off;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device);
    // resident-weight kernel when everything fits (weights + one activation tile in shared memory, all accumulators side by side in
    // tensor memory, 16-column head); otherwise the weights stream (k_policy_forward_wide)
    p->wide = d.total > max_smem || d.n1 > 256 || p->head_n != kHeadN || d.n1 + d.n2 + d.n3 + kHeadN > kTmemCols;
    if (p->wide) {
        WideDims& w = p->wd;
        w.k0 = d.k0; w.k0p = d.k0p; w.n1 = d.n1; w.n2 = d.n2; w.n3 = d.n3; w.n_act = d.n_act; w.head_n = p->head_n;
        w.parts1 = d.n1 > 256 ? 2 : 1;
        const int h1 = d.n1 / w.parts1;
        auto mx = [](int a, int b) { return a > b ? a : b; };
        w.stage_bytes = 2 * 16 * mx(mx(h1, d.n2), mx(d.n3, p->head_n));
        int o = 0;
        w.off_x = o; o += kTileM * 2 * mx(h1, d.n3);
        w.off_y = o; o += kTileM * 2 * mx(d.k0p, d.n2);
        w.off_bias = o; o += round_up((d.n1 + d.n2 + d.n3 + p->head_n) * 4, 16);
        w.off_norm = o; o += round_up(2 * d.k0 * 4, 16);
        w.off_bar = o; o += 256;
        o = round_up(o, 128);
        w.off_ring = o;
        w.stages = (max_smem - o) / w.stage_bytes;
        if (w.stages > kMaxStages) w.stages = kMaxStages;
        if (w.stages < 3) {
            delete p;
            return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_policy_create: activation tiles leave no room for the weight ring in shared memory");
        }
        w.total = o + w.stages * w.stage_bytes;
        w.n_chunks = (d.k0p / 16) * w.parts1 + d.n1 / 16 + d.n2 / 16 + d.n3 / 16;
    }
    const size_t wtot = (size_t)d.w_bytes[0] + d.w_bytes[1] + d.w_bytes[2] + d.w_bytes[3];
    cudaError_t e = cudaMalloc(&p->wpack, wtot);
    if (e == cudaSuccess) e = cudaMalloc(&p->bias, (size_t)(d.n1 + d.n2 + d.n3 + p->head_n) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&p->norm, (size_t)2 * d.k0 * 4);
    if (e == cudaSuccess) e = cudaMemset(p->wpack, 0, wtot);
    if (e == cudaSuccess) e = cudaMemset(p->bias, 0, (size_t)(d.n1 + d.n2 + d.n3 + p->head_n) * 4);
    if (e == cudaSuccess && !p->wide) e = cudaFuncSetAttribute(k_policy_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, d.total);
    if (e == cudaSuccess && p->wide) e = cudaFuncSetAttribute(k_policy_forward_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, p->wd.total);
    if (e == cudaSuccess) {
        k_pack_norm<<<(d.k0 + 127) / 128, 128>>>(nullptr, nullptr, 0.f, d.k0, p->norm);   // identity normalisation
        e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) {
        cudaFree(p->wpack); cudaFree(p->bias); cudaFree(p->norm);
        delete p;
        char b[256];
        snprintf(b, sizeof(b), "b2g_policy_create: %s", cudaGetErrorString(e));
        return b2g::fail_msg(B2G_ERR_CUDA, b);
    }
    *out = p;
    return B2G_OK;
}

void b2g_policy_destroy(b2g_policy* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->wpack);
    cudaFree(p->bias);
    cudaFree(p->norm);
    delete p;
}

int b2g_policy_set_layer(b2g_policy* p, int layer, const float* W_dev, const float* b_dev, void* stream) {
    if (!p || !W_dev || !b_dev) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_set_layer: null argument");
    const PolicyDims& d = p->d;
    int rows, cols, row0 = 0, np, slot;
    size_t woff = 0, boff = 0;
    switch (layer) {
        case B2G_POLICY_HIDDEN0: rows = d.n1; cols = d.k0; np = d.n1; slot = 0; break;
        case B2G_POLICY_HIDDEN1: rows = d.n2; cols = d.n1; np = d.n2; slot = 1; break;
        case B2G_POLICY_HIDDEN2: rows = d.n3; cols = d.n2; np = d.n3; slot = 2; break;
        case B2G_POLICY_MU: rows = d.n_act; cols = d.n3; np = p->head_n; slot = 3; break;
        case B2G_POLICY_VALUE: rows = 1; cols = d.n3; np = p->head_n; slot = 3; row0 = d.n_act; break;
        default: return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_set_layer: layer must be one of B2G_POLICY_*");
    }
    const int nb[4] = {d.n1, d.n2, d.n3, p->head_n};
    for (int l = 0; l < slot; l++) {
        woff += (size_t)d.w_bytes[l];
        boff += (size_t)nb[l];
    }
    CUDA_TRY_P(cudaSetDevice(p->device));
    k_pack_layer<<<(rows * cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        W_dev, b_dev, rows, cols, row0, np, reinterpret_cast<__nv_bfloat16*>(p->wpack + woff), p->bias + boff);
    CUDA_TRY_P(cudaGetLastError());
    p->launches++;
    return B2G_OK;
}

int b2g_policy_set_obs_norm(b2g_policy* p, const float* mean_dev, const float* var_dev, float eps, float clip, void* stream) {
    if (!p) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_set_obs_norm: null policy");
    if (!(clip > 0.f)) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_set_obs_norm: clip must be positive");
    CUDA_TRY_P(cudaSetDevice(p->device));
    k_pack_norm<<<(p->d.k0 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean_dev, var_dev, eps, p->d.k0, p->norm);
    CUDA_TRY_P(cudaGetLastError());
    p->clip = clip;
    p->launches++;
    return B2G_OK;
}

int b2g_policy_forward(b2g_policy* p, const float* obs_dev, int n_rows, float* mu_dev, float* value_dev, void* stream) {
    if (!p || !obs_dev || !mu_dev || !value_dev) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_forward: null argument");
    if (n_rows < 0) return b2g::fail_msg(B2G_ERR_ARG, "b2g_policy_forward: n_rows < 0");
    if (n_rows == 0) return B2G_OK;
    CUDA_TRY_P(cudaSetDevice(p->device));
    const int n_tiles = (n_rows + kTileM - 1) / kTileM;
    const int grid = n_tiles < p->n_sm ? n_tiles : p->n_sm;
    const int aligned = (reinterpret_cast<uintptr_t>(obs_dev) & 15) == 0 ? 1 : 0;
    if (p->wide)
        k_policy_forward_wide<<<grid, kThreads + 32, p->wd.total, (cudaStream_t)stream>>>(p->wd, p->wpack, p->bias, p->norm, p->clip, obs_dev,
                                                                                   n_rows, mu_dev, value_dev, aligned);
    else
        k_policy_forward<<<grid, kThreads, p->d.total, (cudaStream_t)stream>>>(p->d, p->wpack, p->bias, p->norm, p->clip, obs_dev, n_rows, mu_dev,
                                                                              value_dev, aligned);
    CUDA_TRY_P(cudaGetLastError());
    p->launches++;
    return B2G_OK;
}

int64_t b2g_policy_launch_count(const b2g_policy* p) { return p ? p->launches : 0; }

}  // extern "C"
