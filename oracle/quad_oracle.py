"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched WaypointQuadEnv step.

A NumPy/SciPy restatement (float64, vectorised over N envs) of the reference's
hot path, used as the checker for the CUDA kernels.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product package never does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this module
against `tests/golden/*.npz`, which `oracle/gen_golden.py` produced by running
the unmodified reference (`oracle/ref_harness.py`) in the build container
(NumPy 2.3.5, SciPy 1.18.1); when /root/reference is present the same tests
also run the live reference side by side.

What is restated (reference file:line, relative to /root/reference):
  constants          initial-implementation-v2/simul_files/model/params.py:10-43
  state derivative   .../simul_files/model/quadcopter.py:66-103
                     .../simul_files/utils/quaternion.py:46-77 (axis-angle -> matrix)
  mixer/clamp/update .../simul_files/model/quadcopter.py:105-114
  integrator         scipy.integrate.odeint (LSODA) exactly as called at quadcopter.py:113
                     -- third-party, version of this image (SciPy 1.18.1); the
                     fixed-step RK4 path is this repo's throughput mode ("Oracle B")
  v2 step/reward/obs initial-implementation-v2/rl_env_scaledObs.py:98-231
  v2 reset           initial-implementation-v2/rl_env_scaledObs.py:40-96,
                     initial-implementation-v2/utils2/utils.py:12-94
  euler angles       initial-implementation-v2/utils2/utils.py:4-9 (scipy Rotation.as_euler('xyz'))
  v1 step/reward/obs initial-implementation-v1/rl_env_scaledObs.py:65-168 (rl_env.py: raw obs)
  v1 reset           initial-implementation-v1/rl_env_scaledObs.py:32-63
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.integrate import odeint

# ---------------------------------------------------------------------------
# constants (params.py:10-43) -- computed with the same NumPy calls so that
# invA / invI are bit-identical to the reference's.
# ---------------------------------------------------------------------------
MASS = 0.18
GRAV = 9.81
INERTIA = np.array([(0.00025, 0, 2.55e-6), (0, 0.000232, 0), (2.55e-6, 0, 0.0003738)])
INV_INERTIA = np.linalg.inv(INERTIA)
ARM = 0.086
MIN_F = 0.0
MAX_F = 2.0 * MASS * GRAV
KM = 1.5e-9
KF = 6.11e-8
R_YAW = KM / KF
MIX = np.array([[1, 1, 1, 1], [0, ARM, 0, -ARM], [-ARM, 0, ARM, 0], [R_YAW, -R_YAW, R_YAW, -R_YAW]])
INV_MIX = np.linalg.inv(MIX)
DT = 1.0 / 200.0

INFO_SUCCESS = 1
INFO_STOPPED = 2
INFO_CRASHED = 4
INFO_OOB = 8

N_UNIFORMS = 18  # unit uniforms one reset may consume (v2 <= 16, v2 with a drawn waypoint count 17, v1 <= 11)

OBS_DIM = {"v1": 17, "v1_raw": 17, "v2": 20}
MAX_STEPS = {"v1": 1200, "v1_raw": 1200, "v2": 2000}
COUNTER_LIMIT = 500


# ---------------------------------------------------------------------------
# dynamics
# ---------------------------------------------------------------------------
def state_dot(y, F, M):
    """d/dt of the 13-state [pos, vel, quat(wxyz), omega]; y is [...,13], F [...], M [...,3].

    quadcopter.py:66-103.  The reference builds the rotation matrix from the
    axis-angle form of the normalised quaternion (quaternion.py:46-77); the
    closed form below is the same matrix, and only its third row is needed
    because the thrust vector is [0,0,F] in the body frame.
    """
    y = np.asarray(y, dtype=np.float64)
    qw, qx, qy, qz = y[..., 6], y[..., 7], y[..., 8], y[..., 9]
    p, q, r = y[..., 10], y[..., 11], y[..., 12]
    n2 = qw * qw + qx * qx + qy * qy + qz * qz
    inv_n = 1.0 / np.sqrt(n2)
    w, x, yy, z = qw * inv_n, qx * inv_n, qy * inv_n, qz * inv_n
    out = np.empty_like(y)
    out[..., 0:3] = y[..., 3:6]
    out[..., 3] = (2.0 * (x * z - w * yy)) * F / MASS
    out[..., 4] = (2.0 * (yy * z + w * x)) * F / MASS
    out[..., 5] = (1.0 - 2.0 * (x * x + yy * yy)) * F / MASS - GRAV
    qerr = 2.0 * (1.0 - n2)
    out[..., 6] = -0.5 * (-p * qx - q * qy - r * qz) + qerr * qw
    out[..., 7] = -0.5 * (p * qw - r * qy + q * qz) + qerr * qx
    out[..., 8] = -0.5 * (q * qw + r * qx - p * qz) + qerr * qy
    out[..., 9] = -0.5 * (r * qw - q * qx + p * qy) + qerr * qz
    I = INERTIA
    Iw0 = I[0, 0] * p + I[0, 1] * q + I[0, 2] * r
    Iw1 = I[1, 0] * p + I[1, 1] * q + I[1, 2] * r
    Iw2 = I[2, 0] * p + I[2, 1] * q + I[2, 2] * r
    t0 = M[..., 0] - (q * Iw2 - r * Iw1)
    t1 = M[..., 1] - (r * Iw0 - p * Iw2)
    t2 = M[..., 2] - (p * Iw1 - q * Iw0)
    J = INV_INERTIA
    out[..., 10] = J[0, 0] * t0 + J[0, 1] * t1 + J[0, 2] * t2
    out[..., 11] = J[1, 0] * t0 + J[1, 1] * t1 + J[1, 2] * t2
    out[..., 12] = J[2, 0] * t0 + J[2, 1] * t1 + J[2, 2] * t2
    return out


def _state_dot_scalar(y, t, F, M0, M1, M2):
    """Scalar (python float) state derivative for the odeint callback.

    Follows the reference's own arithmetic path (quadcopter.py:66-103): rotation matrix from the
    axis-angle form of the normalised quaternion, theta = 2*arccos(qw/|q|), v = q_xyz/|q_xyz|
    (quaternion.py:46-77), thrust = third row of that matrix times F.  `state_dot` above is the
    closed form of the same thing; the two differ by ~1e-13 relative when qw is close to 1
    (arccos is ill-conditioned there), which LSODA's step control amplifies to ~1e-11 in the state.
    """
    px, py, pz, vx, vy, vz, qw, qx, qy, qz, p, q, r = y
    n2 = qw * qw + qx * qx + qy * qy + qz * qz
    theta = 2.0 * math.acos(qw / math.sqrt(n2))
    ln = math.sqrt(qx * qx + qy * qy + qz * qz)
    if ln > 0.0:
        v0, v1, v2 = qx / ln, qy / ln, qz / ln
    else:
        v0, v1, v2 = qx, qy, qz
    c = math.cos(theta)
    s = math.sin(theta)
    r20 = v2 * v0 * (1. - c) - v1 * s
    r21 = v2 * v1 * (1. - c) + v0 * s
    r22 = v2 * v2 * (1. - c) + c
    inv_m = 1.0 / MASS
    qerr = 2.0 * (1.0 - n2)
    I = _I_LIST
    Iw0 = I[0][0] * p + I[0][1] * q + I[0][2] * r
    Iw1 = I[1][0] * p + I[1][1] * q + I[1][2] * r
    Iw2 = I[2][0] * p + I[2][1] * q + I[2][2] * r
    t0 = M0 - (q * Iw2 - r * Iw1)
    t1 = M1 - (r * Iw0 - p * Iw2)
    t2 = M2 - (p * Iw1 - q * Iw0)
    J = _J_LIST
    return (
        vx, vy, vz,
        inv_m * (r20 * F),
        inv_m * (r21 * F),
        inv_m * (r22 * F - MASS * GRAV),
        -0.5 * (-p * qx - q * qy - r * qz) + qerr * qw,
        -0.5 * (p * qw - r * qy + q * qz) + qerr * qx,
        -0.5 * (q * qw + r * qx - p * qz) + qerr * qy,
        -0.5 * (r * qw - q * qx + p * qy) + qerr * qz,
        J[0][0] * t0 + J[0][1] * t1 + J[0][2] * t2,
        J[1][0] * t0 + J[1][1] * t1 + J[1][2] * t2,
        J[2][0] * t0 + J[2][1] * t1 + J[2][2] * t2,
    )


_I_LIST = INERTIA.tolist()
_J_LIST = INV_INERTIA.tolist()


def scale_action(actions, action_f32: bool = True):
    """Commanded thrust/moments from the 4-D action (v2 :125-126, v1 :87-88).

    With float32 actions NumPy >= 2 keeps the products in float32
    (`a0*0.18` -> f32, `*9.81` -> f32, `a*0.1` -> f32) before the mixer
    promotes them to float64.
    """
    if action_f32:
        a = np.asarray(actions, dtype=np.float32)
        F = (a[..., 0] * np.float32(MASS)) * np.float32(GRAV)
        M = a[..., 1:4] * np.float32(0.1)
        return F.astype(np.float64), M.astype(np.float64)
    a = np.asarray(actions, dtype=np.float64)
    return a[..., 0] * MASS * GRAV, a[..., 1:4] * 0.1


def mix_and_clamp(F, M):
    """Per-prop thrusts through invA, clamp to [minF/4, maxF/4], re-mix (quadcopter.py:109-112)."""
    cmd = np.stack([F, M[..., 0], M[..., 1], M[..., 2]], axis=-1)
    t = cmd @ INV_MIX.T
    t = np.maximum(np.minimum(t, MAX_F / 4), MIN_F / 4)
    Fc = ((t[..., 0] + t[..., 1]) + t[..., 2]) + t[..., 3]
    Mc = t @ MIX[1:].T
    return Fc, Mc, t


def integrate_lsoda(y, F, M, dt=DT, full_output=False):
    """One env step with SciPy's LSODA, called exactly as quadcopter.py:113 does (defaults)."""
    y = np.asarray(y, dtype=np.float64)
    flat = y.reshape(-1, 13)
    Ff = np.asarray(F, dtype=np.float64).reshape(-1)
    Mf = np.asarray(M, dtype=np.float64).reshape(-1, 3)
    out = np.empty_like(flat)
    stats = []
    for i in range(flat.shape[0]):
        args = (float(Ff[i]), float(Mf[i, 0]), float(Mf[i, 1]), float(Mf[i, 2]))
        if full_output:
            sol, info = odeint(_state_dot_scalar, flat[i], [0, dt], args=args, full_output=True)
            stats.append({k: info[k][0] for k in ("nst", "nfe", "nqu", "hu", "tcur", "mused")})
        else:
            sol = odeint(_state_dot_scalar, flat[i], [0, dt], args=args)
        out[i] = sol[1]
    out = out.reshape(y.shape)
    return (out, stats) if full_output else out


def integrate_rk4(y, F, M, dt=DT, substeps=1):
    """Classical RK4 with `substeps` equal sub-intervals (this repo's throughput mode)."""
    y = np.array(y, dtype=np.float64)
    h = dt / substeps
    for _ in range(substeps):
        k1 = state_dot(y, F, M)
        k2 = state_dot(y + (0.5 * h) * k1, F, M)
        k3 = state_dot(y + (0.5 * h) * k2, F, M)
        k4 = state_dot(y + h * k3, F, M)
        y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return y


def physics_update(y, actions, integrator="lsoda", substeps=1, action_f32=True):
    """Quadcopter.update (quadcopter.py:105-114): scale, mix, clamp, integrate, renormalise."""
    F, M = scale_action(actions, action_f32)
    Fc, Mc, _ = mix_and_clamp(F, M)
    if integrator == "lsoda":
        y2 = integrate_lsoda(y, Fc, Mc)
    elif integrator == "rk4":
        y2 = integrate_rk4(y, Fc, Mc, substeps=substeps)
    else:
        raise ValueError(integrator)
    y2 = np.array(y2)
    qn = np.sqrt(np.sum(y2[..., 6:10] ** 2, axis=-1, keepdims=True))
    y2[..., 6:10] = y2[..., 6:10] / qn
    return y2


def quat_to_rpy(qw, qx, qy, qz):
    """scipy Rotation.from_quat([x,y,z,w]).as_euler('xyz') (utils2/utils.py:4-9), restated.

    SciPy's algorithm (Bernardes & Viollet 2022; scipy/spatial/transform/_rotation_xp.py
    `_get_angles`, SciPy 1.18.1): half-angle sum/difference with a 1e-7 gimbal-lock
    window in which the third angle is set to zero.
    """
    qw, qx, qy, qz = (np.asarray(v, dtype=np.float64) for v in (qw, qx, qy, qz))
    n = np.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    x, y, z, w = qx / n, qy / n, qz / n, qw / n
    a = w - y
    b = x + z
    c = y + w
    d = z - x
    half_sum = np.arctan2(b, a)
    half_diff = np.arctan2(d, c)
    ang1 = 2.0 * np.arctan2(np.hypot(c, d), np.hypot(a, b))
    case1 = np.abs(ang1) <= 1e-7
    case2 = np.abs(ang1 - np.pi) <= 1e-7
    case0 = ~(case1 | case2)
    roll = np.where(case0, half_sum - half_diff, np.where(case1, 2.0 * half_sum, -2.0 * half_diff))
    yaw = np.where(case0, half_sum + half_diff, 0.0)
    pitch = ang1 - np.pi / 2
    # the compiled NumPy backend (_rotation_cy) wraps with comparisons, so exactly +pi stays +pi
    wrap = lambda v: np.where(v < -np.pi, v + 2 * np.pi, np.where(v > np.pi, v - 2 * np.pi, v))
    return wrap(roll), wrap(pitch), wrap(yaw)


# ---------------------------------------------------------------------------
# env state container
# ---------------------------------------------------------------------------
@dataclass
class EnvBatch:
    """Struct-of-arrays internal state of N reference envs (one row per env)."""

    version: str
    y: np.ndarray                 # [N,13] f64
    wp_list: np.ndarray           # [N,K,3] f64
    n_wp: np.ndarray              # [N] int
    wp_index: np.ndarray          # [N] int
    cur_wp: np.ndarray            # [N,3] f64
    last_distance: np.ndarray     # [N] f64, NaN == None
    current_step: np.ndarray      # [N] int
    counter: np.ndarray           # [N] int (v2)
    final_reached: np.ndarray     # [N] bool (v2; == counter_activated)
    final_yaw: np.ndarray         # [N] f64 (v2)
    episode: np.ndarray = field(default=None)  # [N] int, number of resets so far (RNG key)

    @property
    def n(self) -> int:
        return self.y.shape[0]

    @classmethod
    def empty(cls, version: str, n: int, max_wp: int | None = None) -> "EnvBatch":
        k = max_wp or (1 if version == "v2" else 2)
        z = lambda *s, dt=np.float64: np.zeros(s, dtype=dt)
        return cls(version, z(n, 13), z(n, k, 3), np.ones(n, dtype=np.int64), z(n, dt=np.int64), z(n, 3),
                   np.full(n, np.nan), z(n, dt=np.int64), z(n, dt=np.int64), z(n, dt=bool), z(n),
                   z(n, dt=np.int64))

    def copy(self) -> "EnvBatch":
        return EnvBatch(self.version, *(np.array(getattr(self, f)) for f in
                                        ("y", "wp_list", "n_wp", "wp_index", "cur_wp", "last_distance",
                                         "current_step", "counter", "final_reached", "final_yaw", "episode")))


# ---------------------------------------------------------------------------
# reset from a block of unit uniforms
# ---------------------------------------------------------------------------
class UniformBlock:
    """np.random look-alike fed by a fixed block of unit uniforms u in [0,1), consumed in order.

    uniform(lo,hi) = lo + (hi-lo)*u ; rand() = u ; randint(lo,hi) = lo + floor(u*(hi-lo)).
    The first two are NumPy's own (legacy) definitions; the third is this repo's definition of an
    unbiased integer draw from one uniform (NumPy's MT19937 randint uses masked rejection instead).
    """

    def __init__(self, u):
        self.u = np.asarray(u, dtype=np.float64)
        self.k = 0

    def rand(self):
        v = self.u[self.k]
        self.k += 1
        return float(v)

    def uniform(self, lo, hi):
        return lo + (hi - lo) * self.rand()

    def randint(self, lo, hi):
        return lo + int(math.floor(self.rand() * (hi - lo)))


def reset_env(b: EnvBatch, i: int, rng) -> None:
    """Reset env `i` drawing from `rng` (np.random itself, or a UniformBlock) in the reference's order.

    v2 (rl_env_scaledObs.py:40-79): x,y,(z) ~U(-1,1)x3, z~U(1,2), three attitude draws and one
    rand() that are consumed but unused (:49-52), trajectory mixture .3/.42/.28 (:63-68;
    utils2/utils.py:12-94), final_yaw ~U(-pi,pi) (:94-96).
    v1 (rl_env_scaledObs.py:32-63): x,y,(z), z~U(1,2), n_wp=randint(1,3), n_wp x [U(-1,1),U(-1,1),U(1,3)].
    """
    uni = rng.uniform
    start = np.array([uni(-1, 1), uni(-1, 1), uni(-1, 1)])
    start[2] = uni(1, 2)
    b.y[i] = 0.0
    b.y[i, 0:3] = start
    b.y[i, 6] = 1.0  # attitude (0,0,0) -> quat (1,0,-0,-0)  (quadcopter.py:25-38, utils.py:25-60)
    b.current_step[i] = 0
    b.wp_index[i] = 0
    b.last_distance[i] = np.nan
    b.wp_list[i] = 0.0
    if b.version == "v2":
        nw = 1  # num_waypoints = 1 (:47)
        if getattr(b, "v2_random_waypoints", False):
            nw = int(rng.randint(2, 4))  # the alternative kept commented out at :46, drawn at that position
        uni(-np.pi / 2, np.pi / 2), uni(-np.pi / 2, np.pi / 2), uni(-np.pi, np.pi)  # unused attitude draws
        rng.rand()  # `rand() < 0`: never true
        b.counter[i] = 0
        b.final_reached[i] = False
        if rng.rand() < 0.3:
            kind = 0
        elif rng.rand() < 0.6:
            kind = 1
        else:
            kind = 2
        wps = []
        if kind in (0, 1):
            end = np.array([uni(-1, 1), uni(-1, 1), uni(-1, 1)])
            end[2] = uni(0.5, 3)
            axis = -1
            if kind == 1:
                axis = 2 - int(rng.randint(0, 3))  # 0 -> z bump, 1 -> y, 2 -> x
            for j in range(1, nw + 1):
                t = j / nw
                wp = start + t * (end - start)
                if kind == 1:
                    bump = np.zeros(3)
                    bump[axis] = np.sin(2 * t * np.pi)
                    wp = wp + bump
                    wp[2] = max(wp[2], 0.2)
                wps.append(wp)
        else:
            total = 2 * np.pi * 1
            for j in range(1, nw + 1):
                ang = (j / nw) * total
                wp = np.array([start[0] + 0.8 * np.cos(ang), start[1] + 0.8 * np.sin(ang), start[2] + j * 0.4])
                wp[2] = max(wp[2], 0.2)
                wps.append(wp)
        b.final_yaw[i] = uni(-np.pi, np.pi)
    else:
        nw = int(rng.randint(1, 3))
        wps = [np.array([uni(-1, 1), uni(-1, 1), uni(1, 3)]) for _ in range(nw)]
    b.n_wp[i] = nw
    for j, wp in enumerate(wps):
        b.wp_list[i, j] = wp
    b.cur_wp[i] = wps[0]


def reset_from_uniforms(b: EnvBatch, idx, u):
    """Reset envs `idx`, env idx[r] consuming the unit uniforms u[r, :N_UNIFORMS] in order."""
    u = np.asarray(u, dtype=np.float64)
    for row, i in enumerate(np.asarray(idx)):
        reset_env(b, int(i), UniformBlock(u[row]))


# ---------------------------------------------------------------------------
# observation
# ---------------------------------------------------------------------------
def observe(b: EnvBatch) -> np.ndarray:
    """_get_observation: v2 :98-121 (20-D), v1 :65-83 (17-D scaled), rl_env.py (17-D raw)."""
    n = b.n
    pos, vel, quat, om = b.y[:, 0:3], b.y[:, 3:6], b.y[:, 6:10], b.y[:, 10:13]
    rel = b.cur_wp - pos
    ar = np.arange(n)
    if b.version == "v2":
        has_next = b.wp_index < b.n_wp - 1
        nxt_i = np.minimum(b.wp_index + 1, b.wp_list.shape[1] - 1)
        rel_next = np.where(has_next[:, None], b.wp_list[ar, nxt_i] - b.cur_wp, 0.0)
        obs = np.concatenate([pos / 10.0, vel / 5.0, quat, om / 5.0, rel / 2.0, rel_next / 2.0,
                              (b.final_yaw / np.pi)[:, None]], axis=1)
    else:
        last = b.wp_list[ar, b.n_wp - 1]
        is_final = np.all(np.abs(b.cur_wp - last) <= 1e-8 + 1e-5 * np.abs(last), axis=1)  # np.allclose
        flag = np.where(is_final, 1.0, 0.0)[:, None]
        if b.version == "v1":
            obs = np.concatenate([pos / 10.0, vel / 5.0, quat, om / 5.0, rel / 2.0, flag], axis=1)
        else:
            obs = np.concatenate([pos, vel, quat, om, rel, flag], axis=1)
    return obs.astype(np.float32)


def _norm3(v):
    return np.sqrt(np.sum(v * v, axis=-1))


# ---------------------------------------------------------------------------
# step (no auto-reset): mutates `b`, returns obs/reward/terminated/truncated/info bits
# ---------------------------------------------------------------------------
def step(b: EnvBatch, actions, integrator="lsoda", substeps=1, action_f32=True):
    b.y = physics_update(b.y, actions, integrator, substeps, action_f32)
    return step_logic(b)


def step_logic(b: EnvBatch):
    """Everything in step() after quadcopter.update(): reward, state machine, observation."""
    if b.version == "v2":
        return _step_v2(b)
    return _step_v1(b)


def _base_reward(b: EnvBatch, dist_gain: float):
    """_calculate_reward common part (v2 :198-222, v1 :142-168); updates last_distance."""
    pos, vel, om = b.y[:, 0:3], b.y[:, 3:6], b.y[:, 10:13]
    d = _norm3(pos - b.cur_wp)
    dist_r = -d * dist_gain
    vn, wn = _norm3(vel), _norm3(om)
    speed = -0.1 * vn ** 2
    speed = np.where(wn > 0.1, speed - 0.01 * wn ** 2, speed)
    has_last = ~np.isnan(b.last_distance)
    prog = 20 * (np.where(has_last, b.last_distance, d) - d)
    prog = np.where(prog > 0, prog + 2, prog)
    prog = np.where(has_last, prog, 0.0)
    b.last_distance = d.copy()
    return d, dist_r, speed, prog, vn, wn


def _step_v2(b: EnvBatch):
    n = b.n
    pos, vel = b.y[:, 0:3], b.y[:, 3:6]
    d, dist_r, speed, prog, vn, wn = _base_reward(b, 10.0)
    time_pen = np.full(n, -0.1)
    fr0 = b.final_reached.copy()
    prog = np.where(fr0, 0.0, prog)
    time_pen = np.where(fr0, 0.0, time_pen)
    dist_r = np.where(fr0 & (d < 0.1), 1.0, dist_r)
    reward = dist_r + speed + time_pen + prog

    roll, pitch, yaw = quat_to_rpy(b.y[:, 6], b.y[:, 7], b.y[:, 8], b.y[:, 9])
    truncated = b.current_step >= MAX_STEPS["v2"]
    b.current_step = b.current_step + 1
    terminated = np.zeros(n, dtype=bool)
    info = np.zeros(n, dtype=np.uint8)

    reached = d < 0.1
    inc = reached & ~fr0
    b.wp_index = np.where(inc, b.wp_index + 1, b.wp_index)
    reward = np.where(inc, reward + 100.0, reward)
    more = reached & (b.wp_index < b.n_wp)
    ar = np.arange(n)
    nxt = b.wp_list[ar, np.minimum(b.wp_index, b.wp_list.shape[1] - 1)]
    b.cur_wp = np.where(more[:, None], nxt, b.cur_wp)

    dyaw = np.abs(yaw - b.final_yaw)
    stopped = (vn < 0.1) & (wn < 0.1)
    # first arrival at the final waypoint (:156-164)
    first = reached & ~more & ~fr0
    stop_b = np.where(vn < 1, 150.0 * (1 - vn ** 2), 0.0)
    yaw_b = np.where(dyaw < 2 * np.pi, 100.0 * (1 - dyaw / (2 * np.pi)), 0.0)
    reward = np.where(first, reward + 200.0 + stop_b + yaw_b, reward)
    b.final_reached = b.final_reached | first
    info = np.where(first, INFO_SUCCESS | np.where(stopped, INFO_STOPPED, 0), info)
    # hold phase (:165-179)
    hold = reached & ~more & fr0
    yaw_h = np.where(dyaw < 2 * np.pi, 30.0 * (1 - dyaw / (2 * np.pi)), 0.0)
    roll_h = np.where(np.abs(roll) < 0.2, 10.0 * (1 - np.abs(roll) / 0.2), -.1 * np.abs(roll))
    pit_h = np.where(np.abs(pitch) < 0.2, 10.0 * (1 - np.abs(pitch) / 0.2), -.1 * np.abs(pitch))
    reward = np.where(hold, reward + yaw_h + roll_h + pit_h, reward)
    hold_run = hold & (b.counter <= COUNTER_LIMIT)
    hold_end = hold & ~hold_run
    b.counter = np.where(hold_run, b.counter + 1, b.counter)
    terminated |= hold_end
    info = np.where(hold, INFO_SUCCESS | np.where(stopped, INFO_STOPPED, 0), info)

    early = first | hold
    # fall-through path (:181-196); `more` falls through too
    rest = ~early
    b.counter = np.where(rest & fr0, b.counter + 1, b.counter)  # counter_activated as of entry
    crash = rest & (pos[:, 2] < 0.1)
    reward = np.where(crash, reward - 100, reward)
    reward = np.where(crash & (vel[:, 2] < 0), reward + vel[:, 2] * 100.0, reward)
    oob = rest & ~crash & (_norm3(pos) > 10)
    reward = np.where(oob, reward - 100.0, reward)
    terminated |= crash | oob
    info = np.where(crash, INFO_CRASHED, info)
    info = np.where(oob, INFO_OOB, info)
    return observe(b), reward, terminated, truncated, info.astype(np.uint8)


def _step_v1(b: EnvBatch):
    n = b.n
    pos, vel = b.y[:, 0:3], b.y[:, 3:6]
    d, dist_r, speed, prog, vn, wn = _base_reward(b, 2.0)
    reward = dist_r + speed + (-0.1) + prog
    wdir = b.cur_wp - pos
    with np.errstate(invalid="ignore", divide="ignore"):
        unit = wdir / _norm3(wdir)[:, None]
    vt = np.sum(vel * unit, axis=1)
    near = d < 0.5
    with np.errstate(invalid="ignore"):
        plus = near & (vt > 0.1)
        minus = near & ~plus & (vt < 0.1)
    reward = np.where(plus, reward + 10.0, reward)
    reward = np.where(minus, reward - 10.0, reward)
    reached = d < 0.1
    reward = np.where(reached, reward + 100.0, reward)
    b.wp_index = np.where(reached, b.wp_index + 1, b.wp_index)
    more = reached & (b.wp_index < b.n_wp)
    ar = np.arange(n)
    nxt = b.wp_list[ar, np.minimum(b.wp_index, b.wp_list.shape[1] - 1)]
    b.cur_wp = np.where(more[:, None], nxt, b.cur_wp)
    success = reached & ~more
    rot_b = np.where(wn < 0.1, 100.0, -20.0 * wn)
    stop_b = np.where(vn < 0.1, 100.0, -10.0 * vn)
    reward = np.where(success, reward + 400.0 + stop_b + rot_b, reward)
    info = np.where(success, INFO_SUCCESS | np.where(vn < 0.1, INFO_STOPPED, 0), 0)
    rest = ~success
    truncated = rest & (b.current_step >= MAX_STEPS["v1"])  # success path: literal False, step not counted
    b.current_step = np.where(rest, b.current_step + 1, b.current_step)
    crash = rest & (pos[:, 2] < 0.1)
    reward = np.where(crash, reward - 100, reward)
    reward = np.where(crash & (vel[:, 2] < 0), reward + vel[:, 2] * 100.0, reward)
    oob = rest & ~crash & (_norm3(pos) > 10)
    reward = np.where(oob, reward - 100.0, reward)
    terminated = success | crash | oob
    info = np.where(crash, INFO_CRASHED, info)
    info = np.where(oob, INFO_OOB, info)
    return observe(b), reward, terminated, truncated, info.astype(np.uint8)


# ---------------------------------------------------------------------------
# vec-env stepping with auto-reset (SB3 DummyVecEnv.step_wait semantics, restated)
# ---------------------------------------------------------------------------
class VecOracle:
    """N reference envs with DummyVecEnv-style auto-reset and Monitor-style episode stats.

    `uniforms(env_ids, episodes) -> [len, N_UNIFORMS]` supplies the reset draws, so the CUDA
    path and this oracle can be driven by the very same Philox stream.
    """

    def __init__(self, version, n, uniforms, integrator="lsoda", substeps=1, action_f32=True, max_wp=None, v2_random_waypoints=False):
        self.b = EnvBatch.empty(version, n, 3 if v2_random_waypoints else max_wp)
        self.b.v2_random_waypoints = bool(v2_random_waypoints)     # rl_env_scaledObs.py:46 alternative (see reset_env)
        self.uniforms = uniforms
        self.integrator, self.substeps, self.action_f32 = integrator, substeps, action_f32
        self.ep_return = np.zeros(n)
        self.ep_len = np.zeros(n, dtype=np.int64)

    def reset(self):
        ids = np.arange(self.b.n)
        self.b.episode[:] = 0
        reset_from_uniforms(self.b, ids, self.uniforms(ids, self.b.episode[ids]))
        self.ep_return[:] = 0
        self.ep_len[:] = 0
        return observe(self.b)

    def step(self, actions):
        obs, rew, term, trunc, info = step(self.b, actions, self.integrator, self.substeps, self.action_f32)
        self.ep_return += rew
        self.ep_len += 1
        done = term | trunc
        out = {"terminal_obs": obs.copy(), "ep_return": self.ep_return.copy(), "ep_len": self.ep_len.copy(),
               "terminated": term, "truncated": trunc, "info": info}
        ids = np.nonzero(done)[0]
        if len(ids):
            self.b.episode[ids] += 1
            reset_from_uniforms(self.b, ids, self.uniforms(ids, self.b.episode[ids]))
            self.ep_return[ids] = 0
            self.ep_len[ids] = 0
            obs = observe(self.b)
        return obs, rew, done, out
