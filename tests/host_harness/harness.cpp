// TEST INFRASTRUCTURE ONLY -- compiles the product's __host__ __device__ headers (csrc/*.cuh) with g++ so
// their logic can be checked on the CPU-only build box.  Nothing in the product loads this library; the
// product path is CUDA-only and fails loudly without a GPU.
#include "../../rl-aerial-manipulator_b200/csrc/qs_lsoda.cuh"
#include "../../rl-aerial-manipulator_b200/csrc/qs_env.cuh"
#include "../../include/quadsim.h"

using namespace qs;

static Model<double> g_model_d;
static Model<float> g_model_f;
static ResetConsts g_rc;
static LsodaTables g_tables;
static double g_rtol, g_atol;
static uint64_t g_seed;

template <typename Real>
static void fill(Model<Real>& m, const qs_config* c) {
    m.mass = (Real)c->mass; m.inv_mass = (Real)(1.0 / c->mass); m.g = (Real)c->g; m.dt = (Real)c->dt;
    m.I00 = (Real)c->inertia[0]; m.I02 = (Real)c->inertia[2]; m.I11 = (Real)c->inertia[4]; m.I20 = (Real)c->inertia[6]; m.I22 = (Real)c->inertia[8];
    m.J00 = (Real)c->inv_inertia[0]; m.J02 = (Real)c->inv_inertia[2]; m.J11 = (Real)c->inv_inertia[4]; m.J20 = (Real)c->inv_inertia[6]; m.J22 = (Real)c->inv_inertia[8];
    for (int i = 0; i < 16; ++i) { m.mix[i] = (Real)c->mix[i]; m.inv_mix[i] = (Real)c->inv_mix[i]; }
    m.tmax = (Real)c->max_prop_thrust; m.tmin = (Real)c->min_prop_thrust;
}

template <typename Real, int VER>
static void load(EnvState<Real, VER>& s, const double* y, const double* wp, int nwp, int idx, double last_d, int step, int counter,
                 int fin, double final_yaw, double ep_ret, int episode) {
    for (int i = 0; i < 13; ++i) s.y[i] = (Real)y[i];
    for (int j = 0; j < EnvState<Real, VER>::NWP; ++j) for (int i = 0; i < 3; ++i) s.wp[j][i] = (Real)wp[j * 3 + i];
    const bool has_last = !(last_d != last_d);
    s.last_d = has_last ? (Real)last_d : Real(0);
    s.final_yaw = (Real)final_yaw; s.ep_ret = (Real)ep_ret; s.episode = (uint32_t)episode;
    s.set(step, counter, idx, nwp, fin != 0, has_last);
}

template <typename Real, int VER>
static void unload(const EnvState<Real, VER>& s, double* y, double* wp, int* ints /*nwp idx step counter fin has_last episode*/, double* reals /*last_d final_yaw ep_ret*/) {
    for (int i = 0; i < 13; ++i) y[i] = (double)s.y[i];
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) wp[j * 3 + i] = j < EnvState<Real, VER>::NWP ? (double)s.wp[j][i] : 0.0;
    ints[0] = s.n_wp(); ints[1] = s.wp_index(); ints[2] = s.step(); ints[3] = s.counter(); ints[4] = s.final_reached(); ints[5] = s.has_last();
    ints[6] = (int)s.episode;
    reals[0] = (double)s.last_d; reals[1] = (double)s.final_yaw; reals[2] = (double)s.ep_ret;
}

// one full env step on the host, mirroring env_step_kernel's per-thread body (no auto-reset)
template <typename Real, int VER>
static void host_step(const Model<Real>& m, int integ, int substeps, int scale_f32, int obs_scaled, double* y, double* wp, int* ints, double* reals,
                      const float* act, float* obs, double* reward, int* flags, int* ep_len, int* ls_int, double* ls_dbl) {
    EnvState<Real, VER> s;
    load<Real, VER>(s, y, wp, ints[0], ints[1], ints[5] ? reals[0] : (0.0 / 0.0), ints[2], ints[3], ints[4], reals[1], reals[2], ints[6]);
    Real Fcmd, Mcmd[3], F, M[3];
    scale_action<Real>(m, act, scale_f32, Fcmd, Mcmd);
    mix_and_clamp<Real>(m, Fcmd, Mcmd, F, M);
    uint32_t fl = 0;
    if (integ == QS_LSODA) {
        if constexpr (sizeof(Real) == 8) {
            LsodaResult r;
            lsoda_advance(m, g_tables, s.y, F, M, m.dt, g_rtol, g_atol, r);
            if (r.status & ~LS_WOULD_SWITCH) fl |= FLAG_LSODA_FAIL;
            ls_int[0] = r.nst; ls_int[1] = r.nfe; ls_int[2] = r.nqu; ls_int[3] = r.status;
            ls_dbl[0] = r.hu; ls_dbl[1] = r.tcur;
        }
    } else {
        rk4_step<Real>(m, s.y, F, M, substeps);
    }
    renormalise_quat<Real>(s.y);
    Real rew;
    fl |= step_logic<Real, VER>(s, rew, *ep_len);
    s.ep_ret += rew;
    make_obs<Real, VER>(s, obs_scaled, obs);
    *reward = (double)rew;
    *flags = (int)fl;
    unload<Real, VER>(s, y, wp, ints, reals);
}

extern "C" {

void hh_configure(const qs_config* c) {
    fill(g_model_d, c);
    fill(g_model_f, c);
    for (int i = 0; i < QS_TRIG_TAB; ++i) { g_rc.sin_tab[i] = c->sin_tab[i]; g_rc.cos_tab[i] = c->cos_tab[i]; }
    lsoda_tables_init(g_tables);
    g_rtol = c->lsoda_rtol; g_atol = c->lsoda_atol; g_seed = c->seed;
}

// raw LSODA advance (no renormalisation): y in/out, F and M already clamped
void hh_lsoda(double* y, double F, const double* M, double tout, int* ints, double* dbls) {
    LsodaResult r;
    lsoda_advance(g_model_d, g_tables, y, F, M, tout, g_rtol, g_atol, r);
    ints[0] = r.nst; ints[1] = r.nfe; ints[2] = r.nqu; ints[3] = r.status;
    dbls[0] = r.hu; dbls[1] = r.tcur;
}

void hh_mix(const float* act, int scale_f32, double* F, double* M) {
    double Fcmd, Mcmd[3];
    scale_action<double>(g_model_d, act, scale_f32, Fcmd, Mcmd);
    mix_and_clamp<double>(g_model_d, Fcmd, Mcmd, *F, M);
}

void hh_step(int version, int f32, int integ, int substeps, int scale_f32, int obs_scaled, double* y, double* wp, int* ints, double* reals,
             const float* act, float* obs, double* reward, int* flags, int* ep_len, int* ls_int, double* ls_dbl) {
    // version 3 = v2 with 2-3 waypoints (ENV_V2M)
    if (version == 3) {
        if (f32) host_step<float, ENV_V2M>(g_model_f, integ, substeps, scale_f32, obs_scaled, y, wp, ints, reals, act, obs, reward, flags, ep_len, ls_int, ls_dbl);
        else host_step<double, ENV_V2M>(g_model_d, integ, substeps, scale_f32, obs_scaled, y, wp, ints, reals, act, obs, reward, flags, ep_len, ls_int, ls_dbl);
    } else if (version == 2) {
        if (f32) host_step<float, ENV_V2>(g_model_f, integ, substeps, scale_f32, obs_scaled, y, wp, ints, reals, act, obs, reward, flags, ep_len, ls_int, ls_dbl);
        else host_step<double, ENV_V2>(g_model_d, integ, substeps, scale_f32, obs_scaled, y, wp, ints, reals, act, obs, reward, flags, ep_len, ls_int, ls_dbl);
    } else {
        if (f32) host_step<float, ENV_V1>(g_model_f, integ, substeps, scale_f32, obs_scaled, y, wp, ints, reals, act, obs, reward, flags, ep_len, ls_int, ls_dbl);
        else host_step<double, ENV_V1>(g_model_d, integ, substeps, scale_f32, obs_scaled, y, wp, ints, reals, act, obs, reward, flags, ep_len, ls_int, ls_dbl);
    }
}

void hh_reset(int version, int obs_scaled, uint64_t env_gid, int episode, double* y, double* wp, int* ints, double* reals, float* obs) {
    if (version == 3) {
        EnvState<double, ENV_V2M> s; s.episode = (uint32_t)episode;
        reset_env<double, ENV_V2M>(s, g_rc, g_seed, env_gid);
        make_obs<double, ENV_V2M>(s, obs_scaled, obs);
        unload<double, ENV_V2M>(s, y, wp, ints, reals);
    } else if (version == 2) {
        EnvState<double, ENV_V2> s; s.episode = (uint32_t)episode;
        reset_env<double, ENV_V2>(s, g_rc, g_seed, env_gid);
        make_obs<double, ENV_V2>(s, obs_scaled, obs);
        unload<double, ENV_V2>(s, y, wp, ints, reals);
    } else {
        EnvState<double, ENV_V1> s; s.episode = (uint32_t)episode;
        reset_env<double, ENV_V1>(s, g_rc, g_seed, env_gid);
        make_obs<double, ENV_V1>(s, obs_scaled, obs);
        unload<double, ENV_V1>(s, y, wp, ints, reals);
    }
}

void hh_uniforms(uint64_t seed, uint64_t env_gid, uint32_t episode, double* u /*[QS_N_UNIFORMS]*/) { reset_uniforms(seed, env_gid, episode, u); }

void hh_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }

void hh_rpy(const double* q, double* rpy) { quat_to_rpy<double>(q, rpy[0], rpy[1], rpy[2]); }

}  // extern "C"
