// qs_vecnorm.cu -- VecNormalize running statistics on the device.
//
// Replaces stable_baselines3 VecNormalize / RunningMeanStd for the reference's
//     env = VecNormalize(env, norm_obs=True, norm_reward=False)      (initial-implementation-v1/rl_train_vecN.py:11)
// SB3 2.6.0 semantics restated (common/running_mean_std.py, common/vec_env/vec_normalize.py):
//   RunningMeanStd(eps=1e-4): mean=0, var=1, count=1e-4
//   update(x):  bm = x.mean(0), bv = x.var(0), bn = len(x);  d = bm - mean; tot = count + bn
//               mean += d*bn/tot;  M2 = var*count + bv*bn + d^2*count*bn/tot;  var = M2/tot;  count = tot
//   normalize_obs: clip((obs - mean)/sqrt(var + 1e-8), -clip_obs, clip_obs).astype(float32)
//   returns = returns*gamma + reward; ret_rms.update(returns); returns[dones] = 0   (runs even with norm_reward=False)
//
// Kernels:
//   moments_partial  per-CTA column sums in float64 of (x - shift) and (x - shift)^2; block size is a multiple
//                    of lcm(32, d) so a thread keeps one column while striding through the row-major [n,d]
//                    batch with coalesced loads; shared-memory reduction inside the CTA (the lanes of a warp hold
//                    different columns -- 32 consecutive floats of a row-major [n,20] batch -- so a shuffle tree
//                    would need a transposing layout that un-coalesces the loads)
//   moments_final    fixed-order sum of the CTA partials -> (n, mean[d], M2[d]); this triplet is what ranks
//                    all-gather over NCCL (2d+1 doubles) -- Chan's merge is associative
//   merge            running stats <- merge of k triplets (one per rank), on device, no host sync
//   apply            normalise + clip, float4 vectorised when rows are 16-byte aligned
//   returns_update   the discounted-return recursion feeding ret_rms
#include "../../include/quadsim.h"
#include "qs_internal.cuh"
#include "qs_exchange.cuh"
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace qs {

constexpr int MOM_MAX_BLOCKS = 592;   // 148 SMs x 4

static int moments_block(int d) {
    int a = 32, b = d;
    while (b) { int t = a % b; a = b; b = t; }
    int bd = 32 / a * d;               // lcm(32, d)
    while (bd < 256) bd *= 2;
    return bd;
}

__global__ void moments_partial_kernel(const float* __restrict__ x, int64_t n, int d, double* __restrict__ partial /*[grid][2d]*/) {
    extern __shared__ double sred[];   // [2][blockDim]
    const int t = threadIdx.x, bd = blockDim.x;
    const int c = t % d;
    const int64_t total = n * (int64_t)d;
    const double shift = (double)__ldg(x + c);            // first row: keeps sum-of-squares well conditioned
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * bd + t; i < total; i += (int64_t)gridDim.x * bd) {
        const double v = (double)__ldcs(x + i) - shift;
        s1 += v;
        s2 = fma(v, v, s2);
    }
    sred[t] = s1;
    sred[bd + t] = s2;
    __syncthreads();
    if (t < d) {
        double a1 = 0.0, a2 = 0.0;
        for (int j = t; j < bd; j += d) { a1 += sred[j]; a2 += sred[bd + j]; }
        partial[(int64_t)blockIdx.x * 2 * d + t] = a1;
        partial[(int64_t)blockIdx.x * 2 * d + d + t] = a2;
    }
}

// Fixed-order sum of the CTA partials: 1024 threads = S slices x 2d columns (S = 1024 / 2d, 25 for d = 20); each slice adds its
// share of the partials in block order (loads unrolled so they overlap), then the slice sums are added in slice order ->
// deterministic, ~2 us instead of a 100 us single-thread chain.
// shift: per-column offset the partial sums were taken around -- row 0 of x (qs_batch_moments) or the running mean
// stats[1..d] (fused moments of the env-step kernel; x == nullptr), or zero when both are null.
// merge != nullptr: the batch triplet is also Chan-merged into the running statistics `merge` (RunningMeanStd.update), same
// arithmetic as vecnorm_merge_kernel with k = 1.
__device__ __forceinline__ void moments_final_body(const float* __restrict__ x, const double* __restrict__ stats,
                                                   const double* __restrict__ partial, int blocks, int64_t n, int d,
                                                   double* out /*[1+2d]*/, double* merge) {
    __shared__ double s[64][64];
    const int w = 2 * d, S = (1024 / w) < 64 ? (1024 / w) : 64;
    const int col = threadIdx.x % w, slice = threadIdx.x / w;
    // fused path (x == nullptr): same offset rule as env_step_kernel -- the previous triplet's mean if it has one.  Read it
    // before the barrier below; the triplet is overwritten after it.
    double shift = 0.0;
    if (threadIdx.x < d) shift = x ? (double)x[threadIdx.x] : (out[0] > 0.0 ? out[1 + threadIdx.x] : (stats ? stats[1 + threadIdx.x] : 0.0));
    if (slice < S) {
        double acc = 0.0;
        int b = slice;
        for (; b + 3 * S < blocks; b += 4 * S) {
            const double v0 = partial[(int64_t)b * w + col], v1 = partial[(int64_t)(b + S) * w + col];
            const double v2 = partial[(int64_t)(b + 2 * S) * w + col], v3 = partial[(int64_t)(b + 3 * S) * w + col];
            acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; b < blocks; b += S) acc += partial[(int64_t)b * w + col];
        s[slice][col] = acc;
    }
    __syncthreads();
    double mean_b = 0.0, m2_b = 0.0;
    const double cnt = (double)n;
    if (threadIdx.x < d) {
        const int c = threadIdx.x;
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < S; ++k) { s1 += s[k][c]; s2 += s[k][d + c]; }
        mean_b = shift + s1 / cnt;                   // batch mean
        m2_b = s2 - s1 * s1 / cnt;                   // batch M2 = sum (x - mean)^2
        out[1 + c] = mean_b;
        out[1 + d + c] = m2_b;
        if (c == 0) out[0] = cnt;
    }
    if (merge) {
        double count = 0.0, mean = 0.0, var = 0.0;
        if (threadIdx.x < d) {
            const int c = threadIdx.x;
            count = merge[0]; mean = merge[1 + c]; var = merge[1 + d + c];
            const double delta = mean_b - mean;
            const double tot = count + cnt;
            mean = mean + delta * cnt / tot;
            const double M2 = var * count + m2_b + delta * delta * count * cnt / tot;
            var = M2 / tot;
            count = tot;
        }
        __syncthreads();                              // every column has read merge[0] before it is rewritten
        if (threadIdx.x < d) {
            merge[1 + threadIdx.x] = mean;
            merge[1 + d + threadIdx.x] = var;
            if (threadIdx.x == 0) merge[0] = count;
        }
    }
}

__global__ void __launch_bounds__(1024) moments_final_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                             const double* __restrict__ partial, int blocks, int64_t n, int d,
                                                             double* __restrict__ out /*[1+2d]*/, double* merge) {
    moments_final_body(x, stats, partial, blocks, n, d, out, merge);
}
// Several ranks: the same reduction, then -- in the same launch -- the peer-memory exchange of the triplet it has just written and the
// Chan merge of every rank's triplet into the running statistics `xstats` (qs_exchange.cuh).  One kernel boundary less on the
// step -> exchange -> policy chain of a sharded rollout, where every boundary is paid at the slowest rank.
__global__ void __launch_bounds__(1024) moments_final_xchg_kernel(const double* __restrict__ stats, const double* __restrict__ partial, int blocks,
                                                                  int64_t n, int d, double* out /*[1+2d]*/, void* const* __restrict__ peers,
                                                                  int rank, int world, unsigned long long* seq, int* failed, double* xstats) {
    moments_final_body(nullptr, stats, partial, blocks, n, d, out, nullptr);
    __threadfence();
    __syncthreads();                                  // the triplet is complete (and visible to the whole CTA) before it is published
    xchg_merge_body(peers, rank, world, d, seq, failed, xstats, out);
}

__global__ void vecnorm_merge_kernel(double* __restrict__ stats, const double* __restrict__ moments, int k, int d) {
    const int c = threadIdx.x;
    if (c >= d) return;
    double count = stats[0], mean = stats[1 + c], var = stats[1 + d + c];
    for (int r = 0; r < k; ++r) {
        const double* m = moments + (int64_t)r * (1 + 2 * d);
        const double bn = m[0];
        if (bn <= 0.0) continue;
        const double delta = m[1 + c] - mean;
        const double tot = count + bn;
        mean = mean + delta * bn / tot;
        const double M2 = var * count + m[1 + d + c] + delta * delta * count * bn / tot;
        var = M2 / tot;
        count = tot;
    }
    __syncthreads();                              // every column has read stats[0] before it is rewritten
    stats[1 + c] = mean;
    stats[1 + d + c] = var;
    if (c == 0) stats[0] = count;
}

__global__ void vecnorm_apply_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t total, int d,
                                     const double* __restrict__ stats, float eps, float clip) {
    // SB3 subtracts the float64 mean from the float32 obs in float64 and only casts the result; with a nearly
    // constant column (quaternion w ~ 1, std ~ 1e-4) a float32 subtraction would lose 3 digits
    __shared__ double s_mean[32], s_istd[32];
    if (threadIdx.x < d) {
        s_mean[threadIdx.x] = stats[1 + threadIdx.x];
        s_istd[threadIdx.x] = 1.0 / sqrt(stats[1 + d + threadIdx.x] + (double)eps);
    }
    __syncthreads();
    const bool vec = (total % 4 == 0) && ((((uintptr_t)x | (uintptr_t)out) & 15) == 0);
    if (vec) {
        const int64_t nv = total / 4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
            float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
            float* pv = reinterpret_cast<float*>(&v);
            int c = (int)((i * 4) % d);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                pv[j] = fminf(fmaxf((float)(((double)pv[j] - s_mean[c]) * s_istd[c]), -clip), clip);
                c = c + 1 == d ? 0 : c + 1;
            }
            reinterpret_cast<float4*>(out)[i] = v;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int c = (int)(i % d);
            out[i] = fminf(fmaxf((float)(((double)x[i] - s_mean[c]) * s_istd[c]), -clip), clip);
        }
    }
}

// returns = returns*gamma + reward (written to `returns`); `snapshot` receives the pre-reset values (what ret_rms sees);
// then returns[done] = 0.
template <typename Real>
__global__ void returns_update_kernel(float* __restrict__ returns, const Real* __restrict__ reward, const uint8_t* __restrict__ flags,
                                      float gamma, int64_t n, float* __restrict__ snapshot) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float r = returns[i] * gamma + (float)reward[i];
        snapshot[i] = r;
        returns[i] = (flags[i] & 3) ? 0.0f : r;
    }
}

thread_local char g_vn_error[256] = "";

static int vn_check(cudaError_t err, const char* what) {
    if (err != cudaSuccess) {
        snprintf(g_vn_error, sizeof(g_vn_error), "%s: %s", what, cudaGetErrorString(err));
        return QS_ECUDA;
    }
    return QS_OK;
}

int launch_moments_final(qs_handle* h, unsigned blocks, cudaStream_t st) {
    const int d = h->cfg.env_version == 2 ? 20 : 17;
    if (h->mom_xchg) {
        const qs_xchg* x = h->mom_xchg;
        moments_final_xchg_kernel<<<1, 1024, 0, st>>>(h->mom_stats, h->mom_scratch, (int)blocks, h->cfg.n_envs, d, h->mom_out, x->peer_dev, x->rank,
                                                       x->world, x->seq, x->failed, h->mom_xchg_stats);
    } else {
        moments_final_kernel<<<1, 1024, 0, st>>>(nullptr, h->mom_stats, h->mom_scratch, (int)blocks, h->cfg.n_envs, d, h->mom_out, h->mom_merge);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error(h, "moments_final_kernel launch failed: %s", cudaGetErrorString(err));
        return QS_ECUDA;
    }
    return QS_OK;
}

}  // namespace qs

using namespace qs;

extern "C" {

const char* qs_vecnorm_last_error(void) { return g_vn_error; }

int64_t qs_moments_scratch_len(int d) { return (int64_t)MOM_MAX_BLOCKS * 2 * d; }

int qs_batch_moments(const float* x, int64_t n, int d, double* moments_out, double* scratch, void* stream) {
    if (!x || !moments_out || !scratch || n < 1 || d < 1 || d > 32) {
        snprintf(g_vn_error, sizeof(g_vn_error), "qs_batch_moments: bad argument (1 <= d <= 32, n >= 1)");
        return QS_EINVAL;
    }
    const int bd = moments_block(d);
    const int64_t total = n * d;
    int64_t blocks = (total + (int64_t)bd * 8 - 1) / ((int64_t)bd * 8);
    if (blocks > MOM_MAX_BLOCKS) blocks = MOM_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    moments_partial_kernel<<<(unsigned)blocks, bd, 2 * bd * sizeof(double), (cudaStream_t)stream>>>(x, n, d, scratch);
    moments_final_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, nullptr, scratch, (int)blocks, n, d, moments_out, nullptr);
    return vn_check(cudaGetLastError(), "qs_batch_moments");
}

int qs_vecnorm_merge(double* stats, const double* moments, int k, int d, void* stream) {
    if (!stats || !moments || k < 1 || d < 1 || d > 32) {
        snprintf(g_vn_error, sizeof(g_vn_error), "qs_vecnorm_merge: bad argument");
        return QS_EINVAL;
    }
    vecnorm_merge_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(stats, moments, k, d);
    return vn_check(cudaGetLastError(), "qs_vecnorm_merge");
}

int qs_vecnorm_apply(const float* x, float* out, int64_t n, int d, const double* stats, double eps, double clip, void* stream) {
    if (!x || !out || !stats || n < 0 || d < 1 || d > 32) {
        snprintf(g_vn_error, sizeof(g_vn_error), "qs_vecnorm_apply: bad argument");
        return QS_EINVAL;
    }
    if (n == 0) return QS_OK;
    const int64_t total = n * d;
    int64_t blocks = (total / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    vecnorm_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, total, d, stats, (float)eps, (float)clip);
    return vn_check(cudaGetLastError(), "qs_vecnorm_apply");
}

int qs_returns_update(float* returns, const void* reward, int reward_is_f64, const uint8_t* flags, float gamma, int64_t n,
                      float* snapshot, void* stream) {
    if (!returns || !reward || !flags || !snapshot || n < 0) {
        snprintf(g_vn_error, sizeof(g_vn_error), "qs_returns_update: bad argument");
        return QS_EINVAL;
    }
    if (n == 0) return QS_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (reward_is_f64)
        returns_update_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(returns, (const double*)reward, flags, gamma, n, snapshot);
    else
        returns_update_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(returns, (const float*)reward, flags, gamma, n, snapshot);
    return vn_check(cudaGetLastError(), "qs_returns_update");
}

}  // extern "C"
