// qs_pid.cu -- batched PID baseline controller (include/quadsim.h: qs_pid_run).
//
// One thread per env: reads the vehicle state from the handle's state pool, runs the reference's cascaded PID
// (initial-implementation-v2/PID Controller/pid_controller.py:37-115) in float64 and writes the wrench and/or the env action
// that commands it.  HBM-bound like the step kernel: pool record in, 6 integrals in/out, 16-byte action out.
#include "../../include/quadsim.h"
#include "qs_internal.cuh"

#include <cuda_runtime.h>
#include <math.h>

namespace qs {

struct PidArgs {
    const void* pool;
    int64_t n;
    qs_pid_gains k;
    double dt, mass, g;
    const double *des_pos, *des_vel, *des_acc, *des_yaw, *des_yawdot;
    double* integral;
    double* wrench;
    float* actions;
    int clip;
};

// Quadcopter.attitude(): RotToRPY (utils.py:11-15) of Quaternion.as_rotation_matrix (quaternion.py:46-77), same route as the
// reference (theta = 2 arccos(qw/|q|), Rodrigues) so that near-hover attitudes carry the same rounding.
__device__ __forceinline__ void attitude_zxy(const double* q, double& phi, double& theta, double& psi) {
    const double nrm = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    const double ang = 2.0 * acos(q[0] / nrm);
    const double len = sqrt((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]);
    double v0 = q[1], v1 = q[2], v2 = q[3];
    if (len > 0.0) { v0 /= len; v1 /= len; v2 /= len; }
    const double c = cos(ang), s = sin(ang), k = 1.0 - c;
    const double r02 = v0 * v2 * k + v1 * s;
    const double r10 = v1 * v0 * k + v2 * s;
    const double r11 = v1 * v1 * k + c;
    const double r12 = v1 * v2 * k - v0 * s;
    const double r22 = v2 * v2 * k + c;
    phi = asin(r12);
    const double cphi = cos(phi);
    theta = atan2(-r02 / cphi, r22 / cphi);
    psi = atan2(-r10 / cphi, r11 / cphi);
}

__device__ __forceinline__ double clampd(double v, double lim) { return fmin(fmax(v, -lim), lim); }

template <typename Real, int VER>
__global__ void __launch_bounds__(128) pid_kernel(const PidArgs a) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    EnvState<Real, VER> st;
    pool_load<Real, VER>(a.pool, a.n, e, st);
    double y[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) y[i] = (double)st.y[i];
    double des[3], dvel[3] = {0.0, 0.0, 0.0}, dacc[3] = {0.0, 0.0, 0.0};
    if (a.des_pos) {
        des[0] = a.des_pos[e * 3]; des[1] = a.des_pos[e * 3 + 1]; des[2] = a.des_pos[e * 3 + 2];
    } else {
        Real w[3];
        st.cur_wp(w);
        des[0] = (double)w[0]; des[1] = (double)w[1]; des[2] = (double)w[2];
    }
    if (a.des_vel) { dvel[0] = a.des_vel[e * 3]; dvel[1] = a.des_vel[e * 3 + 1]; dvel[2] = a.des_vel[e * 3 + 2]; }
    if (a.des_acc) { dacc[0] = a.des_acc[e * 3]; dacc[1] = a.des_acc[e * 3 + 1]; dacc[2] = a.des_acc[e * 3 + 2]; }
    const double dyaw = a.des_yaw ? a.des_yaw[e] : (is_v2(VER) ? (double)st.final_yaw : 0.0);
    const double dyawdot = a.des_yawdot ? a.des_yawdot[e] : 0.0;
    double I[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) I[i] = a.integral[e * 6 + i];

    double phi, theta, psi;
    attitude_zxy(y + 6, phi, theta, psi);
    // position loop (pid_controller.py:52-82); products and sums kept separate (no FMA) in the reference's order
    double cmd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double err = des[i] - y[i];
        const double err_dot = dvel[i] - y[3 + i];
        I[i] = clampd(__dadd_rn(I[i], __dmul_rn(err, a.dt)), a.k.max_integral);
        cmd[i] = __dadd_rn(__dadd_rn(__dadd_rn(dacc[i], __dmul_rn(a.k.kd[i], err_dot)), __dmul_rn(a.k.kp[i], err)), __dmul_rn(a.k.ki[i], I[i]));
    }
    const double F = __dmul_rn(a.mass, __dadd_rn(a.g, cmd[2]));                                    // :85
    const double sp = sin(dyaw), cp = cos(dyaw), ig = 1.0 / a.g;
    const double des_phi = __dmul_rn(ig, __dadd_rn(__dmul_rn(cmd[0], sp), -__dmul_rn(cmd[1], cp)));   // :88
    const double des_theta = __dmul_rn(ig, __dadd_rn(__dmul_rn(cmd[0], cp), __dmul_rn(cmd[1], sp)));  // :89
    const double e_ang[3] = {des_phi - phi, des_theta - theta, dyaw - psi};                        // :90-95
    const double e_rate[3] = {0.0 - y[10], 0.0 - y[11], dyawdot - y[12]};                          // :96-98
    double M[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        I[3 + i] = clampd(__dadd_rn(I[3 + i], __dmul_rn(e_ang[i], a.dt)), a.k.max_integral);       // :101-106
        M[i] = __dadd_rn(__dadd_rn(__dmul_rn(a.k.kp[3 + i], e_ang[i]), __dmul_rn(a.k.kd[3 + i], e_rate[i])), __dmul_rn(a.k.ki[3 + i], I[3 + i]));
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) a.integral[e * 6 + i] = I[i];
    if (a.wrench) reinterpret_cast<double4*>(a.wrench)[e] = make_double4(F, M[0], M[1], M[2]);
    if (a.actions) {
        float4 act = make_float4((float)(F / (a.mass * a.g)), (float)(M[0] / 0.1), (float)(M[1] / 0.1), (float)(M[2] / 0.1));
        if (a.clip) {
            act.x = fminf(fmaxf(act.x, 0.f), 2.f);
            act.y = fminf(fmaxf(act.y, -1.f), 1.f);
            act.z = fminf(fmaxf(act.z, -1.f), 1.f);
            act.w = fminf(fmaxf(act.w, -1.f), 1.f);
        }
        reinterpret_cast<float4*>(a.actions)[e] = act;
    }
}

}  // namespace qs

extern "C" {

void qs_pid_default_gains(qs_pid_gains* out) {
    if (!out) return;
    // pid_controller.py:16-22 (x, y, z, phi, theta, psi) and :34
    const double kp[6] = {3, 3, 1000, 160, 160, 80}, kd[6] = {30, 30, 200, 3, 3, 5}, ki[6] = {1, 1, 10, 1, 1, 1};
    for (int i = 0; i < 6; ++i) { out->kp[i] = kp[i]; out->kd[i] = kd[i]; out->ki[i] = ki[i]; }
    out->max_integral = 100.0;
}

int qs_pid_run(qs_handle* h, const qs_pid_gains* gains, double dt, const double* des_pos, const double* des_vel,
               const double* des_acc, const double* des_yaw, const double* des_yawdot, double* integral,
               double* wrench_out, float* actions_out, int clip_actions, void* stream) {
    using namespace qs;
    if (!h) { set_error(nullptr, "qs_pid_run: null handle"); return QS_EINVAL; }
    if (!gains || !integral || !(dt > 0.0)) { set_error(h, "qs_pid_run: gains, integral and a positive dt are required"); return QS_EINVAL; }
    if (!h->initialized) { set_error(h, "qs_pid_run: call qs_reset (or qs_set_state) first"); return QS_EINVAL; }
    cudaError_t err = cudaSetDevice(h->cfg.device);
    if (err != cudaSuccess) { set_error(h, "qs_pid_run: %s", cudaGetErrorString(err)); return QS_ECUDA; }
    PidArgs a;
    a.pool = h->pool; a.n = h->cfg.n_envs; a.k = *gains; a.dt = dt; a.mass = h->cfg.mass; a.g = h->cfg.g;
    a.des_pos = des_pos; a.des_vel = des_vel; a.des_acc = des_acc; a.des_yaw = des_yaw; a.des_yawdot = des_yawdot;
    a.integral = integral; a.wrench = wrench_out; a.actions = actions_out; a.clip = clip_actions;
    const unsigned blocks = (unsigned)((a.n + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->cfg.precision == QS_F32) QS_FOR_VARIANT(h, (pid_kernel<float, VER><<<blocks, 128, 0, st>>>(a)););
    else QS_FOR_VARIANT(h, (pid_kernel<double, VER><<<blocks, 128, 0, st>>>(a)););
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(h, "qs_pid_run: %s", cudaGetErrorString(err)); return QS_ECUDA; }
    return QS_OK;
}

}  // extern "C"
