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
    __shared__ float red[kHeadBlock / 32];
    __shared__ float tot[kHeadCols];
    const int A = a.n_actions;
    for (int k = 0; k < 4 + A; k++) {
        float s = 0.0f;
        for (int b = threadIdx.x; b < n_blocks; b += kHeadBlock) s += a.partial[(size_t)b * kHeadCols + k];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
            for (int q = 0; q < kHeadBlock / 32; q++) t += red[q];
            tot[k] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float ent = 0.0f;
        for (int k = 0; k < A; k++) ent += a.log_std[k] + 0.5f + 0.9189385332046727f;
        a.out[1] = tot[0]; a.out[2] = tot[1]; a.out[3] = tot[2]; a.out[4] = tot[3]; a.out[5] = ent;
        a.out[0] = tot[0] + 0.5f * a.critic_coef * tot[1] - a.entropy_coef * ent + a.bounds_loss_coef * tot[2];
    }
    if (threadIdx.x < A) a.grad_log_std[threadIdx.x] = tot[4 + threadIdx.x] - a.entropy_coef;
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
