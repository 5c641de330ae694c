// qs_step_lsoda.cu -- the fused env-step kernel with the LSODA (Adams) port: float64 parity mode.
// Compiled with -fmad=false so products and sums round separately like the x86-64 build of ODEPACK
// the reference runs on; LSODA's step control amplifies last-bit differences to ~1e-11 in the state.
#include "qs_internal.cuh"

namespace qs {

int launch_step_lsoda(qs_handle* h, const float* actions, float* obs, double* reward, uint8_t* flags, float* term_obs,
                      double* ep_ret, int32_t* ep_len, cudaStream_t st) {
    StepParams<double> p = base_params<double>(h);
    p.actions = actions;
    p.obs_out = obs;
    p.reward_out = reward;
    p.flags_out = flags;
    p.term_obs_out = term_obs;
    p.ep_ret_out = ep_ret;
    p.ep_len_out = ep_len;
    unsigned grid = 0;
    QS_FOR_VARIANT(h,
        if (p.mom_partial) {
            auto k = env_step_kernel<double, VER, INTEG_LSODA, true>;
            grid = step_grid(h, k, STEP_BLOCK);
            k<<<grid, STEP_BLOCK, 0, st>>>(p);
        } else {
            auto k = env_step_kernel<double, VER, INTEG_LSODA, false>;
            grid = step_grid(h, k, STEP_BLOCK);
            k<<<grid, STEP_BLOCK, 0, st>>>(p);
        });
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error(h, "env_step_kernel<lsoda> launch failed: %s", cudaGetErrorString(err));
        return QS_ECUDA;
    }
    if (p.mom_partial) return launch_moments_final(h, grid, st);
    return QS_OK;
}

}  // namespace qs
