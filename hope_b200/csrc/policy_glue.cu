// policy_glue.cu — the two batched, non-GEMM steps between the env and the policy network in the rollout loop
// (SURVEY.md §8 row f2), as sm_100a kernels instead of chains of elementwise PyTorch ops on (N, 120) float64 tensors:
//
//   hope_state_norm      StateNorm.state_norm (src/model/state_norm.py:25-46): running mean / std of `lidar` and `target`
//                        updated with a whole batch of N observations (two-pass moments per block, Chan merge), every observation
//                        normalised as (x - mean) / (std + 1e-8) and cast to float32 for the network, the action mask cast
//                        along — 3 launches, one read of the float64 observations per pass
//   hope_masked_sample   ActionMask.choose_action (src/model/action_mask.py:199-227): probabilities of the 42 discrete
//                        actions under the policy's Gaussian, clipped log-densities, times the action mask; one action per
//                        env drawn by inverse CDF from a counter-based Philox4x32-10 stream (seed, step, env)
//
// Stateless C entry points (device pointers + stream); the running statistics live in a caller-owned device array.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/hope_b200.h"

namespace hope_glue {

constexpr int NL = HOPE_N_LIDAR, NT = 5, NC = NL + NT, NA = HOPE_N_ACTION;
constexpr int NORM_THREADS = 128;
static_assert(NC <= NORM_THREADS, "one thread per normalised column");

struct Moments { double n, mean, m2; };

__device__ __forceinline__ Moments merge(Moments a, Moments b) {  // Chan et al.: moments of the union of two samples
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    const double n = a.n + b.n, d = b.mean - a.mean;
    Moments r;
    r.n = n;
    r.mean = a.mean + d * (b.n / n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / n);
    return r;
}

// column c of the (N, 125) observation matrix [lidar | target]
__device__ __forceinline__ double load_col(const double *lidar, const double *target, size_t row, int c) {
    return c < NL ? lidar[row * NL + c] : target[row * NT + (c - NL)];
}

// pass 1: block b reduces rows [b R, (b+1) R) column by column (thread = column: consecutive threads read consecutive doubles).
// Two sweeps over the block's rows (the second one hits L2): the column sum with four independent accumulators, then the
// squared deviations from the block mean — no division per element and no loop-carried dependence between the loads.
__global__ void __launch_bounds__(NORM_THREADS) k_norm_partial(const double *__restrict__ lidar, const double *__restrict__ target, int n, int rows_per_block,
                                                               Moments *__restrict__ partial) {
    const int c = threadIdx.x;
    if (c >= NC) return;
    const size_t lo = (size_t)blockIdx.x * rows_per_block, hi = min((size_t)n, lo + rows_per_block);
    const double *base = c < NL ? lidar + c : target + (c - NL);
    const size_t stride = c < NL ? NL : NT;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    size_t r = lo;
    for (; r + 4 <= hi; r += 4) {
        s0 += base[r * stride]; s1 += base[(r + 1) * stride]; s2 += base[(r + 2) * stride]; s3 += base[(r + 3) * stride];
    }
    for (; r < hi; ++r) s0 += base[r * stride];
    const double cnt = (double)(hi - lo), mean = ((s0 + s1) + (s2 + s3)) / fmax(cnt, 1.0);
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
    for (r = lo; r + 4 <= hi; r += 4) {
        const double d0 = base[r * stride] - mean, d1 = base[(r + 1) * stride] - mean, d2 = base[(r + 2) * stride] - mean, d3 = base[(r + 3) * stride] - mean;
        q0 += d0 * d0; q1 += d1 * d1; q2 += d2 * d2; q3 += d3 * d3;
    }
    for (; r < hi; ++r) { const double d = base[r * stride] - mean; q0 += d * d; }
    partial[(size_t)blockIdx.x * NC + c] = Moments{cnt, mean, (q0 + q1) + (q2 + q3)};
}

// pass 2: merge the block partials (MERGE_GROUPS interleaved chains per column, combined through shared memory), then into the
// running statistics stats[0][c] = mean, stats[1][c] = m2 (count kept by the host)
constexpr int MERGE_GROUPS = 8;
__global__ void __launch_bounds__(NORM_THREADS * MERGE_GROUPS) k_norm_merge(const Moments *__restrict__ partial, int n_blocks, double *__restrict__ stats,
                                                                            double count_before, int update, double *__restrict__ scale /* [2][NC]: mean, std + 1e-8 */) {
    __shared__ Moments part[MERGE_GROUPS][NORM_THREADS];
    const int c = threadIdx.x % NORM_THREADS, g = threadIdx.x / NORM_THREADS;
    Moments acc{0.0, 0.0, 0.0};
    if (update && c < NC)
        for (int b = g; b < n_blocks; b += MERGE_GROUPS) acc = merge(acc, partial[(size_t)b * NC + c]);
    part[g][c] = acc;
    __syncthreads();
    if (g != 0 || c >= NC) return;
    Moments run{count_before, stats[c], stats[NC + c]};
    if (update) {
        Moments batch = part[0][c];
        for (int k = 1; k < MERGE_GROUPS; ++k) batch = merge(batch, part[k][c]);
        run = merge(run, batch);
        stats[c] = run.mean; stats[NC + c] = run.m2;
    }
    scale[c] = run.mean;
    scale[NC + c] = sqrt(run.m2 / fmax(run.n, 1.0)) + 1e-8;  // state_norm.py:43-44
}

// pass 3: (x - mean) / (std + 1e-8) -> float32; the action mask is only cast.  Thread = column again (its mean and scale stay
// in registers, no index arithmetic per element), rows strided over the grid.
__global__ void __launch_bounds__(NORM_THREADS) k_norm_apply(const double *__restrict__ lidar, const double *__restrict__ target, const double *__restrict__ mask, int n,
                                                             const double *__restrict__ scale, float *__restrict__ out_lidar, float *__restrict__ out_target,
                                                             float *__restrict__ out_mask) {
    const int c = threadIdx.x;
    if (c < NC) {
        const double mean = scale[c], denom = scale[NC + c];
        const double *src = c < NL ? lidar + c : target + (c - NL);
        float *dst = c < NL ? out_lidar + c : out_target + (c - NL);
        const size_t stride = c < NL ? NL : NT;
        for (size_t r = blockIdx.x; r < (size_t)n; r += gridDim.x) dst[r * stride] = (float)((src[r * stride] - mean) / denom);
    }
    if (out_mask && c < NA)
        for (size_t r = blockIdx.x; r < (size_t)n; r += gridDim.x) out_mask[r * NA + c] = (float)mask[r * NA + c];
}

// Philox4x32-10 (Salmon et al. 2011): counter (env, step), key (seed) -> 4 x 32 random bits
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// one thread per env
__global__ void __launch_bounds__(128) k_masked_sample(int n, const float *__restrict__ mean, const double *__restrict__ log_std, const double *__restrict__ mask,
                                                       const double *__restrict__ actions /* [42][2], policy scale */, uint64_t seed, uint64_t step,
                                                       double *__restrict__ action_out, int32_t *__restrict__ index_out, double *__restrict__ u_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double m[2], sd[2], lognorm[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        m[d] = fmin(fmax((double)mean[2 * i + d], -1.0), 1.0);  // ppo_agent.py:141 clamps the mean to the action range
        sd[d] = exp(log_std[d]);
        lognorm[d] = log(sqrt(2.0 * 3.141592653589793) * sd[d]);  // action_mask.py:215
    }
    double e[NA], total = 0.0;
    for (int j = 0; j < NA; ++j) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const double z = (actions[2 * j + d] - m[d]) / sd[d];
            const double lp = -0.5 * z * z - lognorm[d];
            s += fmin(fmax(lp, -10.0), 10.0);                       // np.clip(log_probabilities, -10, 10) per dimension, then the sum (:216)
        }
        e[j] = exp(s) * mask[(size_t)i * NA + j];                   // :224
        total += e[j];
    }
    uint32_t ctr[4] = {(uint32_t)i, (uint32_t)step, (uint32_t)(step >> 32), 0x484f5045u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = ((double)(((uint64_t)ctr[0] << 21) ^ (ctr[1] >> 11))) * (1.0 / 9007199254740992.0);  // 53 bits, [0, 1)
    const double target = u * total;
    int pick = NA - 1;
    double acc = 0.0;
    for (int j = 0; j < NA; ++j) {  // inverse CDF: first j whose cumulative weight exceeds u * total (np.random.choice's searchsorted)
        acc += e[j];
        if (acc > target) { pick = j; break; }
    }
    while (pick > 0 && e[pick] == 0.0) --pick;  // u * total rounding up to the full sum must not land on a masked action
    action_out[2 * i] = actions[2 * pick]; action_out[2 * i + 1] = actions[2 * pick + 1];
    if (index_out) index_out[i] = pick;
    if (u_out) u_out[i] = u;
}

}  // namespace hope_glue

extern "C" {

int hope_state_norm_scratch_bytes(int n) {
    const int blocks = (n + 63) / 64 > 592 ? 592 : (n + 63) / 64;
    return (int)(sizeof(hope_glue::Moments) * hope_glue::NC * (blocks > 0 ? blocks : 1) + sizeof(double) * 2 * hope_glue::NC);
}

int hope_state_norm(const double *d_lidar, const double *d_target, const double *d_mask, int n, double *d_stats, double count_before, int update,
                    void *d_scratch, float *d_out_lidar, float *d_out_target, float *d_out_mask, void *stream) {
    using namespace hope_glue;
    if (!d_lidar || !d_target || n <= 0 || !d_stats || !d_scratch || !d_out_lidar || !d_out_target || (d_out_mask && !d_mask)) return HOPE_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int blocks = (n + 63) / 64;
    if (blocks > 592) blocks = 592;
    const int rows = (n + blocks - 1) / blocks;
    blocks = (n + rows - 1) / rows;
    double *scale = static_cast<double *>(d_scratch);
    Moments *partial = reinterpret_cast<Moments *>(scale + 2 * NC);
    if (update) k_norm_partial<<<blocks, NORM_THREADS, 0, s>>>(d_lidar, d_target, n, rows, partial);
    k_norm_merge<<<1, NORM_THREADS * MERGE_GROUPS, 0, s>>>(partial, blocks, d_stats, count_before, update, scale);
    k_norm_apply<<<n < 148 * 16 ? n : 148 * 16, NORM_THREADS, 0, s>>>(d_lidar, d_target, d_mask, n, scale, d_out_lidar, d_out_target, d_out_mask);
    return cudaGetLastError() == cudaSuccess ? HOPE_OK : HOPE_ERR_CUDA;
}

int hope_masked_sample(int n, const float *d_mean, const double *d_log_std, const double *d_mask, const double *d_actions, uint64_t seed, uint64_t step,
                       double *d_action_out, int32_t *d_index_out, double *d_u_out, void *stream) {
    if (n <= 0 || !d_mean || !d_log_std || !d_mask || !d_actions || !d_action_out) return HOPE_ERR_INVALID;
    hope_glue::k_masked_sample<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(n, d_mean, d_log_std, d_mask, d_actions, seed, step, d_action_out,
                                                                                              d_index_out, d_u_out);
    return cudaGetLastError() == cudaSuccess ? HOPE_OK : HOPE_ERR_CUDA;
}

}  // extern "C"
