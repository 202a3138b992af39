// libb200gym.so -- learner-side kernels of the PPO update (entry points in include/b200gym.h):
//
//   b2g_ppo_head        the whole loss head of one minibatch in ONE launch (+ a one-block finalize): clipped surrogate, clipped
//                       value loss, bound loss, entropy, KL, and their gradients with respect to the network outputs (mu, value)
//                       and log_std in closed form.  Replaces ~40 element-wise torch kernels of the forward and their autograd
//                       mirrors; the minibatch rows of the rollout buffers are gathered through the index vector in the kernel.
//                       Loss as rl_games' a2c_continuous computes it (the fork states the same structure in-tree for its AMP agent,
//                       learning/common_agent.py:312-400): a_loss + 0.5 critic_coef c_loss - entropy_coef entropy + bounds_coef b_loss.
//   b2g_adam_clip_step  global-norm gradient clipping (nn.utils.clip_grad_norm_, common_agent.py:372-381) + Adam on ONE flat parameter
//                       vector: two launches (block partial sums of g^2; norm, clip coefficient and the update), learning rate and step
//                       count in device memory so the pair replays from a CUDA graph while the adaptive-KL schedule changes the rate.
//
// Everything is deterministic: block partial sums are added in a fixed order, no floating-point atomics.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b200gym.h"

namespace b2g {
int fail_msg(int code, const char* msg);
}

namespace {

constexpr int kHeadBlock = 256;
constexpr int kMaxAct = B2G_PPO_MAX_ACTIONS;
constexpr int kHeadCols = 4 + kMaxAct;      // a_loss, c_loss, b_loss, kl, d/d log_std[k]

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// one thread per minibatch row
__global__ void __launch_bounds__(kHeadBlock) k_ppo_head(b2g_ppo_head_args a) {
    __shared__ float red[kHeadBlock / 32][kHeadCols];
    const int i = blockIdx.x * kHeadBlock + threadIdx.x;
    const int A = a.n_actions;
    float acc[kHeadCols];
#pragma unroll
    for (int k = 0; k < kHeadCols; k++) acc[k] = 0.0f;
    if (i < a.n_rows) {
        const int64_t r = a.index ? a.index[i] : (int64_t)i;      // row of the rollout buffers this minibatch row came from
        const float inv_b = 1.0f / (float)a.n_rows;
        const float* mu = a.mu + (size_t)i * A;
        const float* act = a.actions + (size_t)r * A;
        const float* omu = a.old_mu + (size_t)r * A;
        float nlp = 0.5f * (float)A * 1.8378770664093453f, bl = 0.0f, kl = 0.0f;
        float z[kMaxAct], isg[kMaxAct], gb[kMaxAct];
#pragma unroll
        for (int k = 0; k < kMaxAct; k++) {
            z[k] = 0.0f; isg[k] = 0.0f; gb[k] = 0.0f;
            if (k < A) {
                const float ls = a.log_std[k], m = mu[k];
                isg[k] = expf(-ls);
                z[k] = (act[k] - m) * isg[k];
                nlp += 0.5f * z[k] * z[k] + ls;
                const float hi = fmaxf(m - a.mu_bound, 0.0f), lo = fmaxf(-a.mu_bound - m, 0.0f);
                bl += hi * hi + lo * lo;
                gb[k] = 2.0f * (hi - lo);
                const float dm = m - omu[k];
                kl += dm * dm * 0.5f * isg[k] * isg[k];
            }
        }
        // clipped surrogate (torch.max of two equal values splits its gradient in halves: inside the clip range both halves are -adv)
        const float adv = a.advantages[r];
        const float ratio = expf(a.old_neglogp[r] - nlp);
        const float s1 = -adv * ratio, s2 = -adv * fminf(fmaxf(ratio, 1.0f - a.e_clip), 1.0f + a.e_clip);
        const float a_loss = fmaxf(s1, s2);
        const float g_nlp = (s1 >= s2 ? adv * ratio : 0.0f) * inv_b;      // d a_loss / d nlp = (-adv)(-ratio)
        // clipped value loss
        const float v = a.value[i], fv = a.old_values[r], fr = a.returns[r];
        const float dv = v - fv;
        const float vc = fv + fminf(fmaxf(dv, -a.e_clip), a.e_clip);
        const float l1 = (v - fr) * (v - fr), l2 = (vc - fr) * (vc - fr);
        const float c_loss = fmaxf(l1, l2);
        const bool inside = dv >= -a.e_clip && dv <= a.e_clip;
        float g_v;
        if (l1 > l2) g_v = 2.0f * (v - fr);
        else if (l2 > l1) g_v = inside ? 2.0f * (vc - fr) : 0.0f;
        else g_v = (v - fr) + (inside ? (vc - fr) : 0.0f);                 // tie: half of each branch
        a.grad_value[i] = 0.5f * a.critic_coef * g_v * inv_b;
        float* gm = a.grad_mu + (size_t)i * A;
#pragma unroll
        for (int k = 0; k < kMaxAct; k++) {
            if (k < A) {
                gm[k] = g_nlp * (-z[k] * isg[k]) + a.bounds_loss_coef * gb[k] * inv_b;
                acc[4 + k] = g_nlp * (1.0f - z[k] * z[k]);
            }
        }
        acc[0] = a_loss * inv_b; acc[1] = c_loss * inv_b; acc[2] = bl * inv_b; acc[3] = kl * inv_b;
    }
    // block sums in a fixed order: lanes by butterfly, warps serially
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kHeadCols; k++) {
        if (k < 4 + A) {
            const float s = warp_sum(acc[k]);
            if ((threadIdx.x & 31) == 0) red[w][k] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < 4 + A) {
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < kHeadBlock / 32; q++) s += red[q][threadIdx.x];
        a.partial[(size_t)blockIdx.x * kHeadCols + threadIdx.x] = s;
    }
}

// one block: out[0] = loss, out[1..4] = a_loss, c_loss, b_loss, kl, out[5] = entropy; grad_log_std[k]
__global__ void __launch_bounds__(kHeadBlock) k_ppo_head_finalize(b2g_ppo_head_args a, int n_blocks) {
    __shared__ float tot[kHeadCols];
    const int A = a.n_actions;
    // one warp per column (columns warp, warp + n_warps, ...): lane l adds blocks l, l + 32, ... and the lanes are reduced by the fixed
    // shuffle tree -- deterministic, one barrier instead of two per column
    for (int k = threadIdx.x >> 5; k < 4 + A; k += kHeadBlock / 32) {
        float s = 0.0f;
        for (int b = threadIdx.x & 31; b < n_blocks; b += 32) s += a.partial[(size_t)b * kHeadCols + k];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) tot[k] = s;
    }
    __shared__ float ls[kMaxAct];      // one parallel read of log_std: a serial loop over global loads costs a memory latency per action
    if (threadIdx.x < A) ls[threadIdx.x] = a.log_std[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        float ent = 0.0f;
        for (int k = 0; k < A; k++) ent += ls[k] + 0.5f + 0.9189385332046727f;
        a.out[1] = tot[0]; a.out[2] = tot[1]; a.out[3] = tot[2]; a.out[4] = tot[3]; a.out[5] = ent;
        a.out[0] = tot[0] + 0.5f * a.critic_coef * tot[1] - a.entropy_coef * ent + a.bounds_loss_coef * tot[2];
    }
    if (threadIdx.x < A) a.grad_log_std[threadIdx.x] = tot[4 + threadIdx.x] - a.entropy_coef;
}

// ---- hidden layers of the actor-critic MLP: bias + ELU in one pass, and its backward fused with the bias gradient --------------------
// forward: h = elu(z + b) in place over the GEMM output z (rows x cols, row-major)
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, size_t i4, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&lo); u.y = *reinterpret_cast<unsigned*>(&hi);
    reinterpret_cast<uint2*>(dst)[i4] = u;
}

__global__ void __launch_bounds__(256) k_bias_elu(float* __restrict__ z, const float* __restrict__ b, size_t n4, int cols4, __nv_bfloat16* __restrict__ h16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(z)[i];
        const float4 bb = reinterpret_cast<const float4*>(b)[i % cols4];
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        v.x = v.x > 0.0f ? v.x : expm1f(v.x); v.y = v.y > 0.0f ? v.y : expm1f(v.y);
        v.z = v.z > 0.0f ? v.z : expm1f(v.z); v.w = v.w > 0.0f ? v.w : expm1f(v.w);
        reinterpret_cast<float4*>(z)[i] = v;
        if (h16) store_bf16x4(h16, i, v);      // bf16 copy for the next layer's weight-gradient GEMM
    }
}

// backward: dz = dh * elu'(z) with elu'(z) = 1 (h > 0) or h + 1 (h <= 0) from the stored OUTPUT h; column sums of dz (= the bias
// gradient) per block of kColRows rows into partial[block][col].  Block = (cols / 4) column threads x row lanes.
constexpr int kColRows = 64;       // 512 blocks for a 32768-row minibatch: the pass is a pure stream (14 B per element), it wants every SM loaded several times over
__global__ void __launch_bounds__(256) k_elu_bwd_colsum(const float* __restrict__ dh, const float* __restrict__ h, float* __restrict__ dz,
                                                        float* __restrict__ partial, int rows, int cols, __nv_bfloat16* __restrict__ dz16) {
    extern __shared__ float4 red4[];      // (row lanes, cols / 4)
    const int cols4 = cols >> 2;
    const int c4 = threadIdx.x % cols4, lane = threadIdx.x / cols4, lanes = blockDim.x / cols4;
    const int r0 = blockIdx.x * kColRows, r1 = min(r0 + kColRows, rows);
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    auto one = [&](const float4 g, const float4 o, size_t i) {
        float4 d;
        d.x = g.x * (o.x > 0.0f ? 1.0f : o.x + 1.0f); d.y = g.y * (o.y > 0.0f ? 1.0f : o.y + 1.0f);
        d.z = g.z * (o.z > 0.0f ? 1.0f : o.z + 1.0f); d.w = g.w * (o.w > 0.0f ? 1.0f : o.w + 1.0f);
        if (dz) reinterpret_cast<float4*>(dz)[i] = d;      // null: the caller only needs the bf16 copy (first layer: no dX)
        if (dz16) store_bf16x4(dz16, i, d);
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
    };
    if (lane < lanes) {
        int r = r0 + lane;
        // four rows per trip: all eight loads are issued before the first use (rows added in the same order as a one-row loop)
        for (; r + 3 * lanes < r1; r += 4 * lanes) {
            const size_t i0 = (size_t)r * cols4 + c4, st = (size_t)lanes * cols4;
            const float4 g0 = *(reinterpret_cast<const float4*>(dh) + i0), o0 = *(reinterpret_cast<const float4*>(h) + i0);
            const float4 g1 = *(reinterpret_cast<const float4*>(dh) + i0 + st), o1 = *(reinterpret_cast<const float4*>(h) + i0 + st);
            const float4 g2 = *(reinterpret_cast<const float4*>(dh) + i0 + 2 * st), o2 = *(reinterpret_cast<const float4*>(h) + i0 + 2 * st);
            const float4 g3 = *(reinterpret_cast<const float4*>(dh) + i0 + 3 * st), o3 = *(reinterpret_cast<const float4*>(h) + i0 + 3 * st);
            one(g0, o0, i0); one(g1, o1, i0 + st); one(g2, o2, i0 + 2 * st); one(g3, o3, i0 + 3 * st);
        }
        for (; r < r1; r += lanes) {
            const size_t i = (size_t)r * cols4 + c4;
            one(*(reinterpret_cast<const float4*>(dh) + i), *(reinterpret_cast<const float4*>(h) + i), i);
        }
        red4[lane * cols4 + c4] = acc;
    }
    __syncthreads();
    if (lane == 0) {      // row lanes added in order: deterministic
        for (int q = 1; q < lanes; q++) {
            const float4 t = red4[q * cols4 + c4];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        reinterpret_cast<float4*>(partial)[(size_t)blockIdx.x * cols4 + c4] = acc;
    }
}

// second stage of the column sums: 32 columns x 32 row lanes per block; lane q adds partial rows q, q + 32, ... (four independent running
// sums, so four loads are in flight per thread -- the stage is latency-bound: a few thousand threads reading an L2-resident table) and the
// lanes are added in order (deterministic).  `out` may be scattered: column c goes to out_of(c).
constexpr int kFinLanes = 32;
template <class OutOf>
__device__ __forceinline__ void colsum_finalize_body(const float* __restrict__ partial, int n_blocks, int cols, OutOf out_of) {
    __shared__ float red[kFinLanes][33];
    const int tx = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.0f;
    if (c < cols) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int b = q;
        for (; b + 3 * kFinLanes < n_blocks; b += 4 * kFinLanes) {
            const float v0 = partial[(size_t)b * cols + c], v1 = partial[(size_t)(b + kFinLanes) * cols + c];
            const float v2 = partial[(size_t)(b + 2 * kFinLanes) * cols + c], v3 = partial[(size_t)(b + 3 * kFinLanes) * cols + c];
            s0 += v0; s1 += v1; s2 += v2; s3 += v3;
        }
        for (; b < n_blocks; b += kFinLanes) s0 += partial[(size_t)b * cols + c];
        s = (s0 + s1) + (s2 + s3);
    }
    red[q][tx] = s;
    __syncthreads();
    if (q == 0 && c < cols) {
        for (int k = 1; k < kFinLanes; k++) s += red[k][tx];
        *out_of(c) = s;
    }
}
__global__ void __launch_bounds__(32 * kFinLanes) k_colsum_finalize(const float* __restrict__ partial, int n_blocks, int cols, float* __restrict__ out) {
    colsum_finalize_body(partial, n_blocks, cols, [=](int c) { return out + c; });
}
// heads: entry e = o (H + 1) + c of dWcat goes to the four parameter gradients (d W_mu | d b_mu | d W_v | d b_v)
__global__ void __launch_bounds__(32 * kFinLanes) k_heads_finalize_scatter(const float* __restrict__ partial, int n_blocks, int H, int A, float* __restrict__ dw_mu,
                                                                          float* __restrict__ db_mu, float* __restrict__ dw_v, float* __restrict__ db_v) {
    colsum_finalize_body(partial, n_blocks, (A + 1) * (H + 1), [=](int e) {
        const int o = e / (H + 1), c = e - o * (H + 1);
        return o < A ? (c < H ? dw_mu + o * H + c : db_mu + o) : (c < H ? dw_v + c : db_v);
    });
}

// rows idx[r] of src (cols % 4 == 0) -> dst (float32) and / or dst16 (bf16): the minibatch gather of the update
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ src, const long long* __restrict__ idx, int rows, int cols4, float* __restrict__ dst,
                                                     __nv_bfloat16* __restrict__ dst16) {
    const size_t n4 = (size_t)rows * cols4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols4), c4 = (int)(i - (size_t)r * cols4);
        const float4 v = reinterpret_cast<const float4*>(src)[(size_t)idx[r] * cols4 + c4];
        if (dst) reinterpret_cast<float4*>(dst)[i] = v;
        if (dst16) store_bf16x4(dst16, i, v);
    }
}

// ---- rollout bookkeeping of the learner: what rl_games' play_steps does between two env steps, as a handful of launches --------------
// running mean / variance (rl_games RunningMeanStd, float64 state): per-column batch sums in double, then the parallel-variance merge
// rows per block: at most ~256 blocks, at least 64 rows each (8192 observation rows -> 128 blocks; a fixed 512 left 16 blocks on 148 SMs)
static inline int stat_rows_per_block(int rows) {
    const int r = (rows + 255) / 256;
    return r < 64 ? 64 : r;
}
static inline int stat_blocks(int rows) {
    const int rpb = stat_rows_per_block(rows);
    return (rows + rpb - 1) / rpb;
}
__global__ void __launch_bounds__(256) k_stat_partial(const float* __restrict__ x, int rows, int cols, double* __restrict__ partial, int rows_per_block) {
    extern __shared__ double redd[];      // (lanes, cols, 2)
    const int lanes = cols <= 256 ? 256 / cols : 1;
    const int c = threadIdx.x % cols, lane = threadIdx.x / cols;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    for (int cc = c; cc < cols; cc += 256) {      // cols > 256: every thread walks several columns
        double s = 0.0, q = 0.0;
        if (lane < lanes)
            for (int r = r0 + lane; r < r1; r += lanes) {
                const double v = (double)x[(size_t)r * cols + cc];
                s += v; q += v * v;
            }
        if (cols <= 256) {
            if (lane < lanes) { redd[(lane * cols + c) * 2] = s; redd[(lane * cols + c) * 2 + 1] = q; }
            __syncthreads();
            if (lane == 0) {
                for (int l = 1; l < lanes; l++) { s += redd[(l * cols + c) * 2]; q += redd[(l * cols + c) * 2 + 1]; }
                partial[((size_t)blockIdx.x * cols + c) * 2] = s;
                partial[((size_t)blockIdx.x * cols + c) * 2 + 1] = q;
            }
        } else {
            partial[((size_t)blockIdx.x * cols + cc) * 2] = s;
            partial[((size_t)blockIdx.x * cols + cc) * 2 + 1] = q;
        }
    }
}

// one WARP per column: lane l adds the partial rows l, l + 32, ..., the lanes are combined by the fixed shuffle tree (deterministic), lane 0
// merges the batch moments into the running (mean, var, count) and writes the float32 copies of mean and 1 / sqrt(var + eps).  (One thread
// per column walking all partial rows cost 11-15 us per call -- 24 calls per rollout.)
__global__ void __launch_bounds__(256) k_stat_merge(const double* __restrict__ partial, int n_blocks, int rows, int cols, double* mean, double* var,
                                                    double* count, int bump_count, float* mean_f, float* inv_std_f, float eps) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= cols) return;      // warp-uniform
    double s = 0.0, q = 0.0;
    const double2* p2 = reinterpret_cast<const double2*>(partial);
    for (int b = lane; b < n_blocks; b += 32) {
        const double2 v = p2[(size_t)b * cols + c];
        s += v.x; q += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane) return;
    const double bc = (double)rows, bm = s / bc;
    double bv = q / bc - bm * bm;
    if (bv < 0.0) bv = 0.0;
    const double cnt = *count, tot = cnt + bc, delta = bm - mean[c];
    const double m = mean[c] + delta * bc / tot;
    const double v = (var[c] * cnt + bv * bc + delta * delta * cnt * bc / tot) / tot;
    mean[c] = m; var[c] = v;
    if (mean_f) mean_f[c] = (float)m;
    if (inv_std_f) inv_std_f[c] = (float)(1.0 / sqrt(v + (double)eps));
    // every column reads `count`; it is advanced by a separate one-thread launch (k_count_add) queued behind this kernel
    (void)bump_count;
}
__global__ void k_count_add(double* count, double n) { *count += n; }

// normalised observations of step t into the rollout buffer: clamp((x - mean) * inv_std, +-clip)
__global__ void __launch_bounds__(256) k_normalize_store(const float* __restrict__ x, const float* __restrict__ mean_f, const float* __restrict__ inv_std_f,
                                                         float* __restrict__ out, size_t n, int cols, float clip) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        out[i] = fminf(fmaxf((x[i] - mean_f[c]) * inv_std_f[c], -clip), clip);
    }
}

__device__ __forceinline__ void philox4(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* o) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)c0 * 0xD2511F53ull, p1 = (unsigned long long)c2 * 0xCD9E8D57ull;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// one thread per environment: a = mu + sigma * N(0, 1) (Philox4x32-10 + Box-Muller keyed by (seed, rollout counter, step, env)),
// neglogp, de-normalised value, the step's rows of the rollout buffers, and the clamped action the environment receives
__global__ void __launch_bounds__(256) k_rollout_sample(b2g_rollout_sample_args a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n_envs) return;
    const int A = a.n_actions;
    const unsigned long long epoch = (unsigned long long)*a.rollout_counter;
    float nlp = 0.5f * (float)A * 1.8378770664093453f;
    const size_t row = ((size_t)a.t * a.n_envs + e) * A;
    for (int k0 = 0; k0 < A; k0 += 4) {
        unsigned w[4];
        philox4((unsigned)e, (unsigned)a.t, (unsigned)(k0 >> 2), (unsigned)epoch, (unsigned)(a.seed & 0xffffffffull) ^ (unsigned)(epoch >> 32), (unsigned)(a.seed >> 32), w);
        float g[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {      // two Box-Muller pairs per Philox block
            const float u1 = ((float)(w[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(w[2 * h + 1] >> 8) * (1.0f / 16777216.0f);
            const float r = sqrtf(-2.0f * logf(u1));
            float sn, cs;
            sincosf(6.283185307179586f * u2, &sn, &cs);
            g[2 * h] = r * cs; g[2 * h + 1] = r * sn;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + j;
            if (k < A) {
                const float ls = a.log_std[k], m = a.mu[(size_t)e * A + k];
                const float act = m + expf(ls) * g[j];
                nlp += 0.5f * g[j] * g[j] + ls;
                a.b_actions[row + k] = act;
                a.b_mu[row + k] = m;
                a.env_actions[(size_t)e * A + k] = fminf(fmaxf(act, -a.action_clip), a.action_clip);
            }
        }
    }
    const size_t i = (size_t)a.t * a.n_envs + e;
    a.b_neglogp[i] = nlp;
    a.b_values[i] = a.value[e] * sqrtf((float)(*a.value_var) + a.value_eps) + (float)(*a.value_mean);
}
__global__ void k_counter_add(int64_t* c) { *c += 1; }

// one thread per environment after env.step: shaped reward (reward_shaper scale + value bootstrap on time-outs), done flags, raw
// episode statistics
__global__ void __launch_bounds__(256) k_rollout_post(b2g_rollout_post_args a) {
    __shared__ double red[3][8];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    if (e < a.n_envs) {
        const size_t i = (size_t)a.t * a.n_envs + e;
        const float rew = a.reward[e];
        const bool done = a.flag_bytes == 8 ? reinterpret_cast<const int64_t*>(a.done)[e] != 0 : reinterpret_cast<const unsigned char*>(a.done)[e] != 0;
        const bool tout = a.flag_bytes == 8 ? reinterpret_cast<const int64_t*>(a.time_out)[e] != 0 : reinterpret_cast<const unsigned char*>(a.time_out)[e] != 0;
        a.b_rewards[i] = a.reward_scale * rew + (tout ? a.gamma * a.b_values[i] : 0.0f);
        a.b_dones[i] = done ? 1.0f : 0.0f;
        const float er = a.ep_reward[e] + rew, el = a.ep_length[e] + 1.0f;
        if (done) { f0 = er; f1 = el; f2 = 1.0; }
        a.ep_reward[e] = done ? 0.0f : er;
        a.ep_length[e] = done ? 0.0f : el;
    }
    // block sums, then one atomic per block (logging statistics only)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f0 += __shfl_xor_sync(0xffffffffu, f0, o); f1 += __shfl_xor_sync(0xffffffffu, f1, o); f2 += __shfl_xor_sync(0xffffffffu, f2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = f0; red[1][threadIdx.x >> 5] = f1; red[2][threadIdx.x >> 5] = f2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int q = 0; q < 8; q++) t += red[threadIdx.x][q];
        if (t != 0.0) atomicAdd(a.finished + threadIdx.x, t);
    }
}

// generalised advantage estimation, one thread per environment walking the horizon backwards (rl_games discount_values)
__global__ void __launch_bounds__(256) k_gae(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
                                             const float* __restrict__ v_last, int T, int N, float gamma, float tau, float* __restrict__ adv,
                                             float* __restrict__ ret) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    float last = 0.0f, nv = v_last[e];
    for (int t = T - 1; t >= 0; t--) {
        const size_t i = (size_t)t * N + e;
        const float nonterm = 1.0f - done[i], v = val[i];
        const float delta = rew[i] + gamma * nv * nonterm - v;
        last = delta + gamma * tau * nonterm * last;
        adv[i] = last;
        ret[i] = last + v;
        nv = v;
    }
}

// after the return statistics are merged: normalised returns / values, and advantages normalised by their own mean and UNBIASED std
__global__ void __launch_bounds__(256) k_finish_batch(const float* __restrict__ ret, const float* __restrict__ val, const float* __restrict__ adv,
                                                      const double* __restrict__ vmean, const double* __restrict__ vvar, const double* __restrict__ adv_partial,
                                                      int n_adv_blocks, size_t n, float eps_v, float* __restrict__ f_ret, float* __restrict__ f_val,
                                                      float* __restrict__ f_adv) {
    __shared__ float am_s, as_s;
    if (threadIdx.x == 0) {
        double s = 0.0, q = 0.0;
        for (int b = 0; b < n_adv_blocks; b++) { s += adv_partial[2 * b]; q += adv_partial[2 * b + 1]; }
        const double m = s / (double)n;
        double var = (q - (double)n * m * m) / (double)(n > 1 ? n - 1 : 1);
        if (var < 0.0) var = 0.0;
        am_s = (float)m; as_s = (float)sqrt(var);
    }
    __syncthreads();
    const float vm = (float)*vmean, vis = rsqrtf((float)*vvar + eps_v), am = am_s, ais = 1.0f / (as_s + 1e-8f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        f_ret[i] = (ret[i] - vm) * vis;
        f_val[i] = (val[i] - vm) * vis;
        f_adv[i] = (adv[i] - am) * ais;
    }
}

// ---- backward of the two output heads (mu: A rows, value: 1 row of W, O = A + 1 outputs of the last hidden layer h, width H) in one
// pass over the minibatch: dh = dY W  (rows x H)  and  dWcat = dY^T [h | 1]  (O x (H + 1); the last column is the bias gradient).
// A K = 32768 GEMM with 13 output rows is a poor fit for the library (66 us measured, tools/gemm_probe.py); here one block stages
// kHeadRows rows of dY and h in shared memory, every thread owns a few entries of dWcat and half a row of dh, and block partial sums
// are added in a fixed order by k_colsum_finalize.
constexpr int kHeadRows = 64;
constexpr int kHeadGroups = 4;                       // row groups of a tile: 256 threads = 64 columns x 4 groups of 16 rows
constexpr int kHeadGroupRows = kHeadRows / kHeadGroups;
// Thread (c, g) owns column c (of a 64-wide column chunk) for the 16 rows of group g.  Both products then read ONE value per row that differs
// between the lanes of a warp -- h[r][c], consecutive c: conflict-free -- and the row's O gradients as broadcast 128-bit loads (the rows of
// dY and the rows of W are zero-padded to NQ quads, so no predicate sits in the inner loops):
//   dWcat[o][c] = sum_r dY[r][o] h[r][c]   accumulators acc[o] in registers, the four row groups added in order through shared memory;
//   dh[r][c]    = sum_o dY[r][o] W[o][c]   the column of W in registers, one coalesced store per row.
// The earlier form (one dWcat entry per thread, 64-long dot products out of shared memory, two 32-bit loads per FMA) spent 35 us per
// minibatch in the load/store unit; this one issues ~1/6 of the shared-memory instructions.
template <int NQ>
__global__ void __launch_bounds__(256) k_heads_backward(const float* __restrict__ h, const float* __restrict__ dmu, const float* __restrict__ dv,
                                                        const float* __restrict__ w_mu, const float* __restrict__ w_v, int rows, int H, int A,
                                                        float* __restrict__ dh, float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[];
    constexpr int OP = 4 * NQ;                  // padded number of outputs
    const int O = A + 1, H1 = H + 1;
    float* s_dy = sm;                           // (kHeadRows, OP)
    float* s_w = s_dy + kHeadRows * OP;         // (H, OP): W transposed, so that thread c reads its column as quads
    float* s_h = s_w + H * OP;                  // (kHeadRows, H)
    float* s_red = s_h + kHeadRows * H;         // (kHeadGroups - 1, OP, 64)
    const int tc = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int r0 = blockIdx.x * kHeadRows;
    const int nr = min(kHeadRows, rows - r0);
    if ((H & 3) == 0) {      // the tile is contiguous in h: 128-bit loads, four in flight per thread before the first store
        const int H4 = H >> 2, total4 = kHeadRows * H4;
        const float4* src = reinterpret_cast<const float4*>(h + (size_t)r0 * H);
        for (int base = 0; base < total4; base += 4 * 256) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 256 + threadIdx.x;
                v[u] = (i < total4 && i / H4 < nr) ? src[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 256 + threadIdx.x;
                if (i < total4) reinterpret_cast<float4*>(s_h)[i] = v[u];
            }
        }
    } else {
        for (int r = g; r < kHeadRows; r += kHeadGroups)
            for (int c = tc; c < H; c += 64) s_h[r * H + c] = r < nr ? h[(size_t)(r0 + r) * H + c] : 0.0f;
    }
    for (int i = threadIdx.x; i < kHeadRows * OP; i += blockDim.x) {
        const int r = i / OP, o = i - r * OP;
        s_dy[i] = (r < nr && o < O) ? (o < A ? dmu[(size_t)(r0 + r) * A + o] : dv[r0 + r]) : 0.0f;
    }
    for (int i = threadIdx.x; i < H * OP; i += blockDim.x) {
        const int c = i / OP, o = i - c * OP;
        s_w[i] = o < A ? w_mu[o * H + c] : (o == A ? w_v[c] : 0.0f);
    }
    __syncthreads();
    const int n_out = O * H1;
    float* part = partial + (size_t)blockIdx.x * n_out;
    const float4* dy4 = reinterpret_cast<const float4*>(s_dy) + (size_t)g * kHeadGroupRows * NQ;
    for (int cc = 0; cc < H; cc += 64) {
        const int c = cc + tc;
        const bool live = c < H;
        // ---- dWcat ----
        float acc[OP];
#pragma unroll
        for (int o = 0; o < OP; o++) acc[o] = 0.0f;
        if (live) {
#pragma unroll 4
            for (int r = 0; r < kHeadGroupRows; r++) {
                const float hv = s_h[(g * kHeadGroupRows + r) * H + c];
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const float4 d = dy4[r * NQ + q];
                    acc[4 * q + 0] += d.x * hv; acc[4 * q + 1] += d.y * hv; acc[4 * q + 2] += d.z * hv; acc[4 * q + 3] += d.w * hv;
                }
            }
        }
        if (cc) __syncthreads();      // the previous chunk's s_red has been read
        if (g > 0) {
#pragma unroll
            for (int o = 0; o < OP; o++) s_red[((g - 1) * OP + o) * 64 + tc] = acc[o];
        }
        __syncthreads();
        if (g == 0 && live) {         // groups added in order: deterministic
#pragma unroll
            for (int o = 0; o < OP; o++) {
                float t = acc[o];
#pragma unroll
                for (int k = 0; k < kHeadGroups - 1; k++) t += s_red[(k * OP + o) * 64 + tc];
                if (o < O) part[o * H1 + c] = t;
            }
        }
        // ---- dh ----
        if (live) {
            float4 wq[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) wq[q] = reinterpret_cast<const float4*>(s_w)[(size_t)c * NQ + q];
#pragma unroll 4
            for (int r = 0; r < kHeadGroupRows; r++) {
                const int rr = g * kHeadGroupRows + r;
                float t = 0.0f;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const float4 d = dy4[r * NQ + q];
                    t += d.x * wq[q].x; t += d.y * wq[q].y; t += d.z * wq[q].z; t += d.w * wq[q].w;
                }
                if (rr < nr) dh[(size_t)(r0 + rr) * H + c] = t;
            }
        }
    }
    // bias column of dWcat: column sums of dY over the tile, rows in order
    if (threadIdx.x < O) {
        float t = 0.0f;
        for (int r = 0; r < kHeadRows; r++) t += s_dy[r * OP + threadIdx.x];
        part[threadIdx.x * H1 + H] = t;
    }
}

// ---- forward of the two heads: mu = h W_mu^T + b_mu, value = h W_v^T + b_v in one pass over h (float32 FMA).  As library calls these are a
// 12-column GEMM and a matrix-vector product (8.6 + 7.6 us per minibatch); here a block stages 64 rows of h (row stride H + 1: the lanes of a
// warp read different rows) and W transposed in quads; thread (row, q) produces four outputs.
template <int NQ>
__global__ void __launch_bounds__(64 * NQ) k_heads_forward(const float* __restrict__ h, const float* __restrict__ w_mu, const float* __restrict__ b_mu,
                                                          const float* __restrict__ w_v, const float* __restrict__ b_v, int rows, int H, int A,
                                                          float* __restrict__ mu, float* __restrict__ value) {
    extern __shared__ __align__(16) float sm[];
    constexpr int OP = 4 * NQ, NT = 64 * NQ;
    float* s_w = sm;                      // (H, OP)
    float* s_h = s_w + H * OP;            // (kHeadRows, H + 1)
    const int H1 = H + 1;
    const int r0 = blockIdx.x * kHeadRows;
    const int nr = min(kHeadRows, rows - r0);
    for (int i = threadIdx.x; i < H * OP; i += NT) {
        const int c = i / OP, o = i - c * OP;
        s_w[i] = o < A ? w_mu[o * H + c] : (o == A ? w_v[c] : 0.0f);
    }
    if ((H & 3) == 0) {      // contiguous tile: 128-bit loads, four in flight per thread before the first store
        const int H4 = H >> 2, total4 = kHeadRows * H4;
        const float4* src = reinterpret_cast<const float4*>(h + (size_t)r0 * H);
        for (int base = 0; base < total4; base += 4 * NT) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * NT + threadIdx.x;
                v[u] = (i < total4 && i / H4 < nr) ? src[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * NT + threadIdx.x;
                if (i < total4) {
                    const int r = i / H4, c = (i - r * H4) * 4;
                    float* d = s_h + r * H1 + c;
                    d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < kHeadRows * H; i += NT) {
            const int r = i / H, c = i - r * H;
            s_h[r * H1 + c] = r < nr ? h[(size_t)r0 * H + i] : 0.0f;
        }
    }
    __syncthreads();
    const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 8
    for (int c = 0; c < H; c++) {
        const float hv = s_h[r * H1 + c];
        const float4 w = reinterpret_cast<const float4*>(s_w)[c * NQ + q];
        acc.x += hv * w.x; acc.y += hv * w.y; acc.z += hv * w.z; acc.w += hv * w.w;
    }
    if (r < nr) {
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = 4 * q + k;
            if (o < A) mu[(size_t)(r0 + r) * A + o] = v[k] + b_mu[o];
            else if (o == A) value[r0 + r] = v[k] + b_v[0];
        }
    }
}

// ---- pseudo-random permutation of 0..n-1 (the mini-epoch shuffle of the update, common_agent.py's dataset permutation): a keyed Feistel
// network on the smallest even-width power-of-two domain >= n is a bijection; values that land outside [0, n) are walked through the network
// again (cycle walking keeps it a bijection on [0, n)).  One thread per index, no sort: torch.randperm is five radix-sort passes (~70 us).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
constexpr int kPermRounds = 6;
__global__ void __launch_bounds__(256) k_random_permutation(int64_t* __restrict__ out, unsigned n, int half_bits, uint64_t seed, uint64_t counter) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key[kPermRounds];
    uint64_t z = seed * 0x9e3779b97f4a7c15ull + counter * 0xbf58476d1ce4e5b9ull + 0x94d049bb133111ebull;
#pragma unroll
    for (int r = 0; r < kPermRounds; r++) {      // splitmix64 stream
        z += 0x9e3779b97f4a7c15ull;
        uint64_t t = z;
        t = (t ^ (t >> 30)) * 0xbf58476d1ce4e5b9ull; t = (t ^ (t >> 27)) * 0x94d049bb133111ebull; t ^= t >> 31;
        key[r] = (uint32_t)t;
    }
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t x = i;
    do {
        uint32_t L = x >> half_bits, R = x & mask;
#pragma unroll
        for (int r = 0; r < kPermRounds; r++) {
            const uint32_t f = mix32(R ^ key[r]) & mask;
            const uint32_t t = L ^ f;
            L = R; R = t;
        }
        x = (L << half_bits) | R;
    } while (x >= n);
    out[i] = (int64_t)x;
}

constexpr int kAdamBlock = 256;

__global__ void __launch_bounds__(kAdamBlock) k_sq_partial(const float* __restrict__ g, int n, float scale, float* __restrict__ partial) {
    __shared__ float red[kAdamBlock / 32];
    float s = 0.0f;
    for (int i = blockIdx.x * kAdamBlock + threadIdx.x; i < n; i += gridDim.x * kAdamBlock) {
        const float x = g[i] * scale;
        s += x * x;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int q = 0; q < kAdamBlock / 32; q++) t += red[q];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kAdamBlock) k_adam_clip(b2g_adam_args a, int n_partials) {
    __shared__ float coef_s, bc1_s, bc2_s;
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int q = 0; q < n_partials; q++) t += a.partial[q];      // same order in every block: every block gets the same norm
        const float norm = sqrtf(t);
        coef_s = a.max_grad_norm > 0.0f ? fminf(1.0f, a.max_grad_norm / (norm + 1e-6f)) : 1.0f;
        const int64_t step = *a.step + 1;      // the first block to finish bumps the counter below; all read the old value here
        bc1_s = 1.0f - powf(a.beta1, (float)step);
        bc2_s = 1.0f - powf(a.beta2, (float)step);
        if (blockIdx.x == 0) { a.out_norm[0] = norm; a.out_norm[1] = coef_s; }
    }
    __syncthreads();
    const float coef = coef_s * a.grad_scale, lr = *a.lr, b1 = a.beta1, b2 = a.beta2;
    const float step_size = lr / bc1_s, inv_sqrt_bc2 = rsqrtf(bc2_s);
    for (int i = blockIdx.x * kAdamBlock + threadIdx.x; i < a.n; i += gridDim.x * kAdamBlock) {
        const float g = a.grad[i] * coef;
        const float m = b1 * a.exp_avg[i] + (1.0f - b1) * g;
        const float v = b2 * a.exp_avg_sq[i] + (1.0f - b2) * g * g;
        a.exp_avg[i] = m; a.exp_avg_sq[i] = v;
        a.param[i] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    }
}

// the step counter advances after every block has read it: a separate one-thread launch keeps the pair race-free and graph-replayable
__global__ void k_bump(int64_t* step) { *step += 1; }

}  // namespace

extern "C" {

int b2g_ppo_head(const b2g_ppo_head_args* a, void* stream) {
    if (!a || !a->mu || !a->value || !a->grad_mu || !a->grad_value || !a->partial || !a->out || !a->grad_log_std)
        return b2g::fail_msg(B2G_ERR_ARG, "b2g_ppo_head: null argument");
    if (a->n_actions < 1 || a->n_actions > kMaxAct) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_ppo_head: 1..B2G_PPO_MAX_ACTIONS actions");
    if (a->n_rows < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_ppo_head: empty minibatch");
    const int blocks = (a->n_rows + kHeadBlock - 1) / kHeadBlock;
    cudaStream_t st = (cudaStream_t)stream;
    k_ppo_head<<<blocks, kHeadBlock, 0, st>>>(*a);
    k_ppo_head_finalize<<<1, kHeadBlock, 0, st>>>(*a, blocks);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_ppo_head: launch failed");
}

int b2g_ppo_head_workspace_floats(int n_rows) { return ((n_rows + kHeadBlock - 1) / kHeadBlock) * kHeadCols; }

int b2g_mlp_bias_elu(float* z, const float* bias, int rows, int cols, void* h_bf16, void* stream) {
    if (!z || !bias || rows < 1 || cols < 4 || (cols & 3)) return b2g::fail_msg(B2G_ERR_ARG, "b2g_mlp_bias_elu: cols must be a positive multiple of 4");
    const size_t n4 = (size_t)rows * cols / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    k_bias_elu<<<blocks, 256, 0, (cudaStream_t)stream>>>(z, bias, n4, cols / 4, reinterpret_cast<__nv_bfloat16*>(h_bf16));
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_mlp_bias_elu: launch failed");
}

int b2g_mlp_elu_backward_workspace_floats(int rows, int cols) { return ((rows + kColRows - 1) / kColRows) * cols; }

int b2g_mlp_elu_backward(const float* dh, const float* h, float* dz, float* dbias, float* partial, int rows, int cols, void* dz_bf16, void* stream) {
    if (!dh || !h || (!dz && !dz_bf16) || !dbias || !partial || rows < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_mlp_elu_backward: null argument");
    if (cols < 4 || (cols & 3) || cols > 1024) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_mlp_elu_backward: cols must be a multiple of 4, at most 1024");
    const int cols4 = cols / 4;
    int lanes = 256 / cols4;
    if (lanes < 1) lanes = 1;
    const int threads = lanes * cols4;          // <= 256
    const int blocks = (rows + kColRows - 1) / kColRows;
    cudaStream_t st = (cudaStream_t)stream;
    k_elu_bwd_colsum<<<blocks, threads, sizeof(float4) * lanes * cols4, st>>>(dh, h, dz, partial, rows, cols, reinterpret_cast<__nv_bfloat16*>(dz_bf16));
    k_colsum_finalize<<<(cols + 31) / 32, 32 * kFinLanes, 0, st>>>(partial, blocks, cols, dbias);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_mlp_elu_backward: launch failed");
}

int b2g_stat_workspace_doubles(int rows, int cols) { return stat_blocks(rows) * cols * 2; }

int b2g_running_stat_update(const float* x, int rows, int cols, double* mean, double* var, double* count, double* partial, float* mean_f32,
                            float* inv_std_f32, float eps, void* stream) {
    if (!x || !mean || !var || !count || !partial || rows < 1 || cols < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_running_stat_update: bad argument");
    const int blocks = stat_blocks(rows);
    const int lanes = cols <= 256 ? 256 / cols : 1;
    cudaStream_t st = (cudaStream_t)stream;
    k_stat_partial<<<blocks, 256, cols <= 256 ? sizeof(double) * 2 * lanes * cols : 0, st>>>(x, rows, cols, partial, stat_rows_per_block(rows));
    k_stat_merge<<<(cols + 7) / 8, 256, 0, st>>>(partial, blocks, rows, cols, mean, var, count, 1, mean_f32, inv_std_f32, eps);
    k_count_add<<<1, 1, 0, st>>>(count, (double)rows);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_running_stat_update: launch failed");
}

int b2g_normalize_store(const float* x, const float* mean_f32, const float* inv_std_f32, float* out, int rows, int cols, float clip, void* stream) {
    if (!x || !mean_f32 || !inv_std_f32 || !out || rows < 1 || cols < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_normalize_store: bad argument");
    const size_t n = (size_t)rows * cols;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_normalize_store<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, mean_f32, inv_std_f32, out, n, cols, clip);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_normalize_store: launch failed");
}

int b2g_rollout_sample(const b2g_rollout_sample_args* a, void* stream) {
    if (!a || !a->mu || !a->value || !a->log_std || !a->b_actions || !a->b_mu || !a->b_neglogp || !a->b_values || !a->env_actions || !a->rollout_counter ||
        !a->value_mean || !a->value_var)
        return b2g::fail_msg(B2G_ERR_ARG, "b2g_rollout_sample: null argument");
    k_rollout_sample<<<(a->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_rollout_sample: launch failed");
}

int b2g_rollout_counter_advance(int64_t* counter, void* stream) {
    if (!counter) return b2g::fail_msg(B2G_ERR_ARG, "null counter");
    k_counter_add<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "launch failed");
}

int b2g_rollout_post(const b2g_rollout_post_args* a, void* stream) {
    if (!a || !a->reward || !a->done || !a->time_out || !a->b_values || !a->b_rewards || !a->b_dones || !a->ep_reward || !a->ep_length || !a->finished)
        return b2g::fail_msg(B2G_ERR_ARG, "b2g_rollout_post: null argument");
    if (a->flag_bytes != 1 && a->flag_bytes != 8) return b2g::fail_msg(B2G_ERR_ARG, "b2g_rollout_post: flags are 1-byte (bool / uint8) or 8-byte (int64)");
    k_rollout_post<<<(a->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_rollout_post: launch failed");
}

int b2g_gae_finish(const b2g_gae_args* a, void* stream) {
    if (!a || !a->rewards || !a->values || !a->dones || !a->v_last || !a->adv || !a->ret || !a->f_ret || !a->f_val || !a->f_adv || !a->value_mean ||
        !a->value_var || !a->value_count || !a->partial)
        return b2g::fail_msg(B2G_ERR_ARG, "b2g_gae_finish: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = a->horizon, N = a->n_envs;
    const int n = T * N;
    k_gae<<<(N + 255) / 256, 256, 0, st>>>(a->rewards, a->values, a->dones, a->v_last, T, N, a->gamma, a->tau, a->adv, a->ret);
    // running statistics of the returns (value normaliser), then the advantages' own moments
    const int blocks = stat_blocks(n), rpb = stat_rows_per_block(n);
    k_stat_partial<<<blocks, 256, sizeof(double) * 2 * 256, st>>>(a->ret, n, 1, a->partial, rpb);
    k_stat_merge<<<1, 32, 0, st>>>(a->partial, blocks, n, 1, a->value_mean, a->value_var, a->value_count, 1, nullptr, nullptr, 0.0f);
    k_count_add<<<1, 1, 0, st>>>(a->value_count, (double)n);
    k_stat_partial<<<blocks, 256, sizeof(double) * 2 * 256, st>>>(a->adv, n, 1, a->partial + 2 * blocks, rpb);
    const int fb = (n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8;
    k_finish_batch<<<fb, 256, 0, st>>>(a->ret, a->values, a->adv, a->value_mean, a->value_var, a->partial + 2 * blocks, blocks, (size_t)n, a->value_eps,
                                       a->f_ret, a->f_val, a->f_adv);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_gae_finish: launch failed");
}

int b2g_mlp_heads_forward(const float* h, const float* w_mu, const float* b_mu, const float* w_v, const float* b_v, int rows, int hidden, int n_actions,
                          float* mu, float* value, void* stream) {
    if (!h || !w_mu || !b_mu || !w_v || !b_v || !mu || !value || rows < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_mlp_heads_forward: null argument");
    if (hidden < 1 || hidden > 256 || n_actions < 1 || n_actions > kMaxAct) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_mlp_heads_forward: hidden <= 256, actions <= 24");
    const int O = n_actions + 1, NQ = (O + 3) / 4, OP = 4 * NQ;
    const size_t smem = sizeof(float) * ((size_t)hidden * OP + (size_t)kHeadRows * (hidden + 1));
    if (smem > 96 * 1024) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_mlp_heads_forward: tile does not fit shared memory");
    const int blocks = (rows + kHeadRows - 1) / kHeadRows;
    cudaStream_t st = (cudaStream_t)stream;
    static bool opted[8] = {false, false, false, false, false, false, false, false};
    auto go = [&](auto kern) {
        if (!opted[NQ]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); opted[NQ] = true; }
        kern<<<blocks, 64 * NQ, smem, st>>>(h, w_mu, b_mu, w_v, b_v, rows, hidden, n_actions, mu, value);
    };
    switch (NQ) {
        case 1: go(k_heads_forward<1>); break;
        case 2: go(k_heads_forward<2>); break;
        case 3: go(k_heads_forward<3>); break;
        case 4: go(k_heads_forward<4>); break;
        case 5: go(k_heads_forward<5>); break;
        case 6: go(k_heads_forward<6>); break;
        default: go(k_heads_forward<7>); break;
    }
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_mlp_heads_forward: launch failed");
}

int b2g_mlp_heads_backward_workspace_floats(int rows, int hidden, int n_actions) {
    return ((rows + kHeadRows - 1) / kHeadRows) * (n_actions + 1) * (hidden + 1);
}

static int heads_backward_launch(const float* h, const float* dmu, const float* dv, const float* w_mu, const float* w_v, int rows, int hidden, int n_actions,
                                 float* dh, float* partial, cudaStream_t st, int& blocks) {
    if (!h || !dmu || !dv || !w_mu || !w_v || !dh || !partial || rows < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_mlp_heads_backward: null argument");
    if (hidden < 1 || hidden > 256 || n_actions < 1 || n_actions > kMaxAct) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_mlp_heads_backward: hidden <= 256, actions <= 24");
    const int O = n_actions + 1, NQ = (O + 3) / 4, OP = 4 * NQ;
    const size_t smem = sizeof(float) * ((size_t)kHeadRows * OP + (size_t)hidden * OP + (size_t)kHeadRows * hidden + (size_t)(kHeadGroups - 1) * OP * 64);
    if (smem > 96 * 1024) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_mlp_heads_backward: tile does not fit shared memory");
    blocks = (rows + kHeadRows - 1) / kHeadRows;
    static bool opted[8] = {false, false, false, false, false, false, false, false};      // per instantiation; first use is the learner's warm-up, outside any capture
    auto go = [&](auto kern) {
        if (!opted[NQ < 7 ? NQ : 7]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); opted[NQ < 7 ? NQ : 7] = true; }
        kern<<<blocks, 256, smem, st>>>(h, dmu, dv, w_mu, w_v, rows, hidden, n_actions, dh, partial);
    };
    switch (NQ) {
        case 1: go(k_heads_backward<1>); break;
        case 2: go(k_heads_backward<2>); break;
        case 3: go(k_heads_backward<3>); break;
        case 4: go(k_heads_backward<4>); break;
        case 5: go(k_heads_backward<5>); break;
        case 6: go(k_heads_backward<6>); break;
        default: go(k_heads_backward<7>); break;
    }
    return B2G_OK;
}

int b2g_mlp_heads_backward(const float* h, const float* dmu, const float* dv, const float* w_mu, const float* w_v, int rows, int hidden, int n_actions,
                           float* dh, float* dw_cat, float* partial, void* stream) {
    if (!dw_cat) return b2g::fail_msg(B2G_ERR_ARG, "b2g_mlp_heads_backward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = 0;
    const int rc = heads_backward_launch(h, dmu, dv, w_mu, w_v, rows, hidden, n_actions, dh, partial, st, blocks);
    if (rc != B2G_OK) return rc;
    const int n_out = (n_actions + 1) * (hidden + 1);
    k_colsum_finalize<<<(n_out + 31) / 32, 32 * kFinLanes, 0, st>>>(partial, blocks, n_out, dw_cat);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_mlp_heads_backward: launch failed");
}

int b2g_mlp_heads_backward_scatter(const float* h, const float* dmu, const float* dv, const float* w_mu, const float* w_v, int rows, int hidden, int n_actions,
                                   float* dh, float* dw_mu, float* db_mu, float* dw_v, float* db_v, float* partial, void* stream) {
    if (!dw_mu || !db_mu || !dw_v || !db_v) return b2g::fail_msg(B2G_ERR_ARG, "b2g_mlp_heads_backward_scatter: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = 0;
    const int rc = heads_backward_launch(h, dmu, dv, w_mu, w_v, rows, hidden, n_actions, dh, partial, st, blocks);
    if (rc != B2G_OK) return rc;
    const int n_out = (n_actions + 1) * (hidden + 1);
    k_heads_finalize_scatter<<<(n_out + 31) / 32, 32 * kFinLanes, 0, st>>>(partial, blocks, hidden, n_actions, dw_mu, db_mu, dw_v, db_v);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_mlp_heads_backward_scatter: launch failed");
}

int b2g_random_permutation(int64_t* out, int n, uint64_t seed, uint64_t counter, void* stream) {
    if (!out || n < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_random_permutation: bad argument");
    int bits = 1;
    while ((1ll << bits) < (long long)n) bits++;
    bits += bits & 1;      // even width: two equal halves
    if (bits < 2) bits = 2;
    k_random_permutation<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, (unsigned)n, bits / 2, seed, counter);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_random_permutation: launch failed");
}

int b2g_gather_rows(const float* src, const int64_t* index, int rows, int cols, float* dst, void* dst_bf16, void* stream) {
    if (!src || !index || (!dst && !dst_bf16) || rows < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_gather_rows: null argument");
    if (cols < 4 || (cols & 3)) return b2g::fail_msg(B2G_ERR_UNSUPPORTED, "b2g_gather_rows: cols must be a multiple of 4");
    const size_t n4 = (size_t)rows * (cols >> 2);
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    k_gather_rows<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<const long long*>(index), rows, cols >> 2, dst,
                                                            reinterpret_cast<__nv_bfloat16*>(dst_bf16));
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_gather_rows: launch failed");
}

int b2g_adam_clip_step(const b2g_adam_args* a, void* stream) {
    if (!a || !a->param || !a->grad || !a->exp_avg || !a->exp_avg_sq || !a->lr || !a->step || !a->partial || !a->out_norm)
        return b2g::fail_msg(B2G_ERR_ARG, "b2g_adam_clip_step: null argument");
    if (a->n < 1) return b2g::fail_msg(B2G_ERR_ARG, "b2g_adam_clip_step: empty parameter vector");
    int blocks = (a->n + kAdamBlock - 1) / kAdamBlock;
    if (blocks > B2G_ADAM_MAX_PARTIALS) blocks = B2G_ADAM_MAX_PARTIALS;
    cudaStream_t st = (cudaStream_t)stream;
    k_sq_partial<<<blocks, kAdamBlock, 0, st>>>(a->grad, a->n, a->grad_scale, a->partial);
    k_adam_clip<<<blocks, kAdamBlock, 0, st>>>(*a, blocks);
    k_bump<<<1, 1, 0, st>>>(a->step);
    return cudaGetLastError() == cudaSuccess ? B2G_OK : b2g::fail_msg(B2G_ERR_CUDA, "b2g_adam_clip_step: launch failed");
}

}  // extern "C"
