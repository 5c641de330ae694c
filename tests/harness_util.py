"""Test helper: builds tests/host_harness/harness.cpp with g++ (the product's __host__ __device__ headers
compiled for the CPU) and wraps it with ctypes.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_harness", "harness.cpp")
OUT = os.path.join(ROOT, "build", "libqs_host.so")
CSRC = os.path.join(ROOT, "rl-aerial-manipulator_b200", "csrc")


def build_host_harness() -> str:
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        # -ffp-contract=off: products and sums round separately, as in the reference's x86-64 NumPy/ODEPACK
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", OUT])
    return OUT


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostHarness:
    def __init__(self, cfg):
        self.lib = C.CDLL(build_host_harness())
        self.lib.hh_configure(C.byref(cfg))
        self.cfg = cfg
        L = self.lib
        L.hh_lsoda.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.hh_mix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hh_step.argtypes = [C.c_int] * 6 + [C.c_void_p] * 11
        L.hh_reset.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int] + [C.c_void_p] * 5
        L.hh_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.hh_philox.argtypes = [C.c_void_p] * 3
        L.hh_rpy.argtypes = [C.c_void_p] * 2

    def mix(self, act, scale_f32=True):
        act = np.ascontiguousarray(act, dtype=np.float32)
        F = C.c_double()
        M = np.zeros(3)
        self.lib.hh_mix(ptr(act), int(scale_f32), C.byref(F), ptr(M))
        return F.value, M

    def lsoda(self, y, F, M, tout=0.005):
        y = np.array(y, dtype=np.float64)
        M = np.ascontiguousarray(M, dtype=np.float64)
        ints = np.zeros(4, dtype=np.int32)
        dbl = np.zeros(2)
        self.lib.hh_lsoda(ptr(y), float(F), ptr(M), float(tout), ptr(ints), ptr(dbl))
        return y, dict(nst=int(ints[0]), nfe=int(ints[1]), nqu=int(ints[2]), status=int(ints[3]), hu=dbl[0], tcur=dbl[1])

    def step(self, version, st, act, f32=False, integ="lsoda", substeps=1, scale_f32=True, obs_scaled=True):
        """st: dict(y, wp_list[3,3], n_wp, wp_index, last_distance, current_step, counter, final_reached, final_yaw)."""
        y = np.array(st["y"], dtype=np.float64)
        wp = np.zeros(9)
        wp[:] = np.asarray(st["wp_list"], dtype=np.float64).reshape(-1)[:9]
        ld = st["last_distance"]
        has_last = not (ld is None or np.isnan(ld))
        ints = np.array([st["n_wp"], st["wp_index"], st["current_step"], st.get("counter", 0),
                         int(st.get("final_reached", False)), int(has_last), st.get("episode", 0)], dtype=np.int32)
        reals = np.array([ld if has_last else 0.0, st.get("final_yaw", 0.0), st.get("ep_return", 0.0)], dtype=np.float64)
        act = np.ascontiguousarray(act, dtype=np.float32)
        obs = np.zeros(20 if version in (2, 3) else 17, dtype=np.float32)
        rew = C.c_double()
        flags = C.c_int()
        ep_len = C.c_int()
        ls_i = np.zeros(4, dtype=np.int32)
        ls_d = np.zeros(2)
        self.lib.hh_step(version, int(f32), 1 if integ == "lsoda" else 0, substeps, int(scale_f32), int(obs_scaled), ptr(y), ptr(wp),
                         ptr(ints), ptr(reals), ptr(act), ptr(obs), C.byref(rew), C.byref(flags), C.byref(ep_len), ptr(ls_i), ptr(ls_d))
        out = dict(y=y, wp_list=wp.reshape(3, 3), n_wp=int(ints[0]), wp_index=int(ints[1]), current_step=int(ints[2]),
                   counter=int(ints[3]), final_reached=bool(ints[4]), last_distance=reals[0] if ints[5] else np.nan,
                   final_yaw=reals[1], ep_return=reals[2])
        return out, obs, rew.value, flags.value, ep_len.value, ls_i

    def reset(self, version, env_gid, episode, obs_scaled=True):
        y = np.zeros(13)
        wp = np.zeros(9)
        ints = np.zeros(7, dtype=np.int32)
        reals = np.zeros(3)
        obs = np.zeros(20 if version in (2, 3) else 17, dtype=np.float32)
        self.lib.hh_reset(version, int(obs_scaled), env_gid, episode, ptr(y), ptr(wp), ptr(ints), ptr(reals), ptr(obs))
        return dict(y=y, wp_list=wp.reshape(3, 3), n_wp=int(ints[0]), wp_index=int(ints[1]), current_step=int(ints[2]),
                    counter=int(ints[3]), final_reached=bool(ints[4]), has_last=bool(ints[5]), final_yaw=reals[1]), obs

    def uniforms(self, seed, env_gid, episode):
        u = np.zeros(18)
        self.lib.hh_uniforms(seed, env_gid, episode, ptr(u))
        return u

    def philox(self, ctr, key):
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        self.lib.hh_philox(ptr(c), ptr(k), ptr(o))
        return o


def golden_state(g, prefix, i):
    return dict(y=g[prefix + "y"][i], wp_list=g[prefix + "wp_list"][i], n_wp=int(g[prefix + "n_wp"][i]),
                wp_index=int(g[prefix + "wp_index"][i]), last_distance=float(g[prefix + "last_distance"][i]),
                current_step=int(g[prefix + "current_step"][i]), counter=int(g[prefix + "counter"][i]),
                final_reached=bool(g[prefix + "final_reached"][i]), final_yaw=float(g[prefix + "final_yaw"][i]))


def flags_from_golden(g, i):
    """QS_FLAG_* byte the kernel must produce for golden case i."""
    f = 0
    if g["terminated"][i]:
        f |= 0x01
    if g["truncated"][i]:
        f |= 0x02
    info = int(g["info"][i])
    f |= (info & 1) << 2 | ((info >> 1) & 1) << 3 | ((info >> 2) & 1) << 4 | ((info >> 3) & 1) << 5
    return f
