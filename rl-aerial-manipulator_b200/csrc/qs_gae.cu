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

}  // extern "C"
