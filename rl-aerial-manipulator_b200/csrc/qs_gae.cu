// qs_gae.cu -- generalized advantage estimation over a device-resident rollout buffer.
//
// Replaces stable_baselines3 RolloutBuffer.compute_returns_and_advantage (called by PPO.collect_rollouts; reference call
// sites: model.learn(...) in initial-implementation-v1/rl_train_vecN.py:36 with gamma=0.995, gae_lambda=0.9 (:22-23), and
// initial-implementation-v2/rl_train.py:56).  SB3 2.6.0 semantics restated:
//     last_gae_lam = 0
//     for step in reversed(range(T)):
//         next_non_terminal = 1 - (dones_last if step == T-1 else episode_starts[step+1])
//         next_values       = last_values    if step == T-1 else values[step+1]
//         delta = rewards[step] + gamma*next_values*next_non_terminal - values[step]
//         last_gae_lam = delta + gamma*gae_lambda*next_non_terminal*last_gae_lam
//         advantages[step] = last_gae_lam
//     returns = advantages + values
// Layout [T, N] (time-major like SB3's buffer): one thread per env walks T backwards, every access of a warp is a coalesced
// 128-byte row segment.  HBM-bound: 3 reads + 2 writes of 4 bytes per (t, env).
#include "../../include/quadsim.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace qs {

__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ episode_starts, const float* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, int T, int64_t n, float gamma, float lam,
                                                  float* __restrict__ advantages, float* __restrict__ returns) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float next_value = last_values[e];
    float next_non_terminal = last_dones[e] ? 0.f : 1.f;
    float gae = 0.f;
    for (int t = T - 1; t >= 0; --t) {
        const int64_t i = (int64_t)t * n + e;
        const float v = values[i];
        const float delta = rewards[i] + gamma * next_value * next_non_terminal - v;
        gae = delta + gamma * lam * next_non_terminal * gae;
        advantages[i] = gae;
        returns[i] = gae + v;
        next_value = v;
        next_non_terminal = episode_starts[i] ? 0.f : 1.f;
    }
}

// ---- RolloutBuffer.add for one step of all envs, slot t read from device memory (so a captured CUDA graph of the step can be replayed) ----
// What SB3's OnPolicyAlgorithm.collect_rollouts does around env.step (stable_baselines3 2.6.0; reference call sites as above):
//   before the step: buffer.add(obs, actions, ..., episode_starts = last dones, values, log_probs)
//   after the step:  rewards (VecNormalize.normalize_reward if enabled) + gamma * V(terminal_observation) for TimeLimit.truncated envs,
//                    last dones, Monitor's episode statistics
// The arithmetic is the float32 arithmetic of the torch expressions it replaces in ppo.py (no FMA contraction), so the buffers are
// bit-identical to the eager loop's.
__global__ void __launch_bounds__(256) record_pre_kernel(const long long* __restrict__ t_ptr, int64_t n, int d, const float* __restrict__ obs,
                                                         const float* __restrict__ actions, const float* __restrict__ values,
                                                         const float* __restrict__ logp, const uint8_t* __restrict__ last_dones,
                                                         float* __restrict__ obs_buf, float* __restrict__ act_buf, float* __restrict__ val_buf,
                                                         float* __restrict__ logp_buf, uint8_t* __restrict__ starts_buf) {
    const int64_t t = *t_ptr;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n * d; i += stride) obs_buf[t * n * d + i] = obs[i];
    for (int64_t i = tid; i < n * 4; i += stride) act_buf[t * n * 4 + i] = actions[i];
    for (int64_t e = tid; e < n; e += stride) {
        val_buf[t * n + e] = values[e];
        logp_buf[t * n + e] = logp[e];
        starts_buf[t * n + e] = last_dones[e];
    }
}

__global__ void __launch_bounds__(256) record_post_kernel(long long* t_ptr, int64_t n, const void* __restrict__ reward, int reward_f64,
                                                          const uint8_t* __restrict__ flags, const void* __restrict__ ep_return,
                                                          const float* __restrict__ term_values, float gamma, const double* __restrict__ ret_var,
                                                          double eps, float clip_reward, float* __restrict__ rew_buf,
                                                          uint8_t* __restrict__ last_dones, double* __restrict__ ep_stats,
                                                          double* __restrict__ partial, unsigned int* __restrict__ ticket) {
    const int64_t t = *t_ptr;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    float denom = 1.f;
    if (ret_var) denom = (float)sqrt(ret_var[0] + eps);
    double s_ret = 0.0, s_cnt = 0.0;
    for (int64_t e = tid; e < n; e += stride) {
        float r = reward_f64 ? (float)static_cast<const double*>(reward)[e] : static_cast<const float*>(reward)[e];
        if (ret_var) r = fminf(fmaxf(__fdiv_rn(r, denom), -clip_reward), clip_reward);
        const uint32_t f = flags[e] & 3u;
        if (f == 2u) r = __fadd_rn(r, __fmul_rn(gamma, term_values[e]));        // time limit only: bootstrap with gamma * V(terminal obs)
        rew_buf[t * n + e] = r;
        last_dones[e] = f != 0u;
        if (f != 0u) {
            s_ret += reward_f64 ? static_cast<const double*>(ep_return)[e] : (double)static_cast<const float*>(ep_return)[e];
            s_cnt += 1.0;
        }
    }
    // deterministic reduction: fixed tree per CTA, the last CTA to finish adds the CTA sums in index order
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_ret += __shfl_down_sync(0xffffffffu, s_ret, o);
        s_cnt += __shfl_down_sync(0xffffffffu, s_cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s_ret; sh[1][threadIdx.x >> 5] = s_cnt; }
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (unsigned k = 0; k < gridDim.x; ++k) { a += __ldcg(partial + 2 * k); b += __ldcg(partial + 2 * k + 1); }
        ep_stats[0] += a;
        ep_stats[1] += b;
        *ticket = 0u;
        *t_ptr = t + 1;
    }
}

thread_local char g_gae_error[256] = "";

}  // namespace qs

extern "C" {

const char* qs_gae_last_error(void) { return qs::g_gae_error; }

int qs_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values,
           const uint8_t* last_dones, int T, int64_t n, float gamma, float gae_lambda, float* advantages, float* returns,
           void* stream) {
    if (!rewards || !values || !episode_starts || !last_values || !last_dones || !advantages || !returns || T < 1 || n < 1) {
        snprintf(qs::g_gae_error, sizeof(qs::g_gae_error), "qs_gae: bad argument");
        return QS_EINVAL;
    }
    qs::gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, values, episode_starts, last_values, last_dones, T, n,
                                                                                  gamma, gae_lambda, advantages, returns);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        snprintf(qs::g_gae_error, sizeof(qs::g_gae_error), "qs_gae: %s", cudaGetErrorString(err));
        return QS_ECUDA;
    }
    return QS_OK;
}

int qs_rollout_record_pre(const long long* t_dev, int64_t n, int obs_dim, const float* obs, const float* actions, const float* values,
                          const float* logp, const uint8_t* last_dones, float* obs_buf, float* actions_buf, float* values_buf,
                          float* logp_buf, uint8_t* episode_starts_buf, void* stream) {
    if (!t_dev || !obs || !actions || !values || !logp || !last_dones || !obs_buf || !actions_buf || !values_buf || !logp_buf ||
        !episode_starts_buf || n < 1 || obs_dim < 1) {
        snprintf(qs::g_gae_error, sizeof(qs::g_gae_error), "qs_rollout_record_pre: bad argument");
        return QS_EINVAL;
    }
    int64_t blocks = (n * obs_dim + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    qs::record_pre_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(t_dev, n, obs_dim, obs, actions, values, logp, last_dones, obs_buf,
                                                                              actions_buf, values_buf, logp_buf, episode_starts_buf);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        snprintf(qs::g_gae_error, sizeof(qs::g_gae_error), "qs_rollout_record_pre: %s", cudaGetErrorString(err));
        return QS_ECUDA;
    }
    return QS_OK;
}

int qs_rollout_record_post(long long* t_dev, int64_t n, const void* reward, int reward_is_f64, const uint8_t* flags, const void* ep_return,
                           const float* terminal_values, float gamma, const double* ret_var, double epsilon, float clip_reward,
                           float* rewards_buf, uint8_t* last_dones, double* ep_stats, double* workspace, void* stream) {
    if (!t_dev || !reward || !flags || !ep_return || !terminal_values || !rewards_buf || !last_dones || !ep_stats || !workspace || n < 1) {
        snprintf(qs::g_gae_error, sizeof(qs::g_gae_error), "qs_rollout_record_post: bad argument");
        return QS_EINVAL;
    }
    int64_t blocks = (n + 255) / 256;
    if (blocks > QS_RECORD_MAX_BLOCKS) blocks = QS_RECORD_MAX_BLOCKS;
    // workspace: f64[2 * QS_RECORD_MAX_BLOCKS] CTA sums, then one zero-initialised 8-byte word used as the ticket
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace + 2 * QS_RECORD_MAX_BLOCKS);
    qs::record_post_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(t_dev, n, reward, reward_is_f64, flags, ep_return, terminal_values,
                                                                               gamma, ret_var, epsilon, clip_reward, rewards_buf, last_dones,
                                                                               ep_stats, workspace, ticket);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        snprintf(qs::g_gae_error, sizeof(qs::g_gae_error), "qs_rollout_record_post: %s", cudaGetErrorString(err));
        return QS_ECUDA;
    }
    return QS_OK;
}

}  // extern "C"
