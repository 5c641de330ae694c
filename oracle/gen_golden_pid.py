"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/pid.npz from the UNMODIFIED reference PID controller.

Runs only in the build container (needs /root/reference):
    python oracle/gen_golden_pid.py
Imports `initial-implementation-v2/PID Controller/pid_controller.py` and its own `model.quadcopter.Quadcopter` as they are,
puts each recorded state into a Quadcopter, seeds the module-level `integral_error` dict, calls `run(quad, des_state, dt)` and
records (state, desired state, integral before, dt) -> (F, M, integral after).  A closed-loop block follows one vehicle for 400
controller+`Quadcopter.update` steps (the loop of `PID Controller/runsim.py:27-32`) so integral carry-over is pinned as well.
"""
from __future__ import annotations

import os
import sys
from collections import namedtuple

import numpy as np

sys.dont_write_bytecode = True
REF = os.path.join(os.environ.get("QS_REFERENCE_ROOT", "/root/reference"), "initial-implementation-v2", "PID Controller")
sys.path.insert(0, REF)
import pid_controller as pid  # noqa: E402
from model.quadcopter import Quadcopter  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pid.npz")
DesiredState = namedtuple("DesiredState", "pos vel acc yaw yawdot")   # trajGen3D.py:13
KEYS = ["x", "y", "z", "phi", "theta", "psi"]


def rand_quat(rng, spread):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = rng.uniform(-spread, spread)
    return np.array([np.cos(ang / 2), *(np.sin(ang / 2) * ax)])


def call(state, des, integ, dt):
    quad = Quadcopter((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    quad.state = np.array(state, dtype=np.float64)
    for k, v in zip(KEYS, integ):
        pid.integral_error[k] = float(v)
    F, M = pid.run(quad, des, dt)
    return float(F), np.asarray(M, dtype=np.float64).reshape(3), np.array([float(pid.integral_error[k]) for k in KEYS])


def main():
    rng = np.random.default_rng(20251018)
    rows = {k: [] for k in ("y", "des_pos", "des_vel", "des_acc", "des_yaw", "des_yawdot", "integral_in", "dt", "F", "M", "integral_out")}
    n = 400
    for i in range(n):
        kind = i % 8
        spread = [0.0, 1e-6, 1e-3, 0.1, 0.5, 1.2, 3.0, 0.3][kind]
        q = rand_quat(rng, spread) * (1.0 if kind != 5 else rng.uniform(0.9, 1.1))   # un-normalised quaternions too
        y = np.concatenate([rng.uniform(-3, 3, 3), rng.normal(size=3) * (0 if kind == 0 else 1.5), q, rng.normal(size=3) * (0 if kind == 0 else 2.0)])
        des = DesiredState(rng.uniform(-3, 3, 3), rng.normal(size=3) * 0.5, rng.normal(size=3) * 0.3, float(rng.uniform(-np.pi, np.pi)),
                           float(rng.normal() * 0.2))
        if kind == 1:   # hover target on the spot: zero errors
            des = DesiredState(y[0:3].copy(), np.zeros(3), np.zeros(3), 0.0, 0.0)
        integ = rng.normal(size=6) * 2.0
        if kind == 6:   # at / beyond the anti-windup limit, both signs
            integ = np.array([99.999, -99.999, 100.0, -100.0, 150.0, -150.0]) * rng.choice([-1.0, 1.0])
        dt = [0.005, 0.01][i % 2]
        F, M, integ_out = call(y, des, integ, dt)
        for k, v in zip(rows, (y, des.pos, des.vel, des.acc, des.yaw, des.yawdot, integ, dt, F, M, integ_out)):
            rows[k].append(np.asarray(v, dtype=np.float64))
    out = {k: np.stack(v) for k, v in rows.items()}

    # closed loop, PID Controller/runsim.py:27-32 with a fixed hover target instead of the minimum-snap trajectory
    quad = Quadcopter((0.5, 0.0, 0.0), (0.0, 0.0, 0.0))
    for k in KEYS:
        pid.integral_error[k] = 0.0
    des = DesiredState(np.array([0.8, -0.3, 0.6]), np.zeros(3), np.zeros(3), 0.4, 0.0)
    ys, Fs, Ms, Is = [], [], [], []
    for _ in range(400):
        ys.append(quad.state.copy())
        F, M = pid.run(quad, des, 0.01)
        Fs.append(float(F)); Ms.append(np.asarray(M, dtype=np.float64).reshape(3)); Is.append([pid.integral_error[k] for k in KEYS])
        quad.update(0.01, F, M)
    out.update(loop_y=np.stack(ys), loop_F=np.array(Fs), loop_M=np.stack(Ms), loop_integral=np.array(Is, dtype=np.float64),
               loop_des_pos=des.pos, loop_des_yaw=np.float64(des.yaw), loop_y_end=quad.state.copy(),
               gains_kp=np.array([pid.k_p_x, pid.k_p_y, pid.k_p_z, pid.k_p_phi, pid.k_p_theta, pid.k_p_psi], dtype=np.float64),
               gains_kd=np.array([pid.k_d_x, pid.k_d_y, pid.k_d_z, pid.k_d_phi, pid.k_d_theta, pid.k_d_psi], dtype=np.float64),
               gains_ki=np.array([pid.k_i_x, pid.k_i_y, pid.k_i_z, pid.k_i_phi, pid.k_i_theta, pid.k_i_psi], dtype=np.float64),
               max_integral=np.float64(pid.MAX_INTEGRAL))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
