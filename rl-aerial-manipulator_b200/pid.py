"""Batched PID baseline controller over libquadsim (`qs_pid_run`, csrc/qs_pid.cu).

Host-side mirror of `initial-implementation-v2/PID Controller/pid_controller.py`: the reference exposes module-level gains
(:16-22), one module-level `integral_error` dict (:24-31) and `run(quad, des_state, dt) -> (F, M)` (:37-115).  Here one
`BatchedPID` serves all envs of a `BatchedQuadEnv`: `run(des_state, dt)` returns F [n] and M [n,3] as float64 CUDA tensors,
`actions(des_state, dt)` the float32 env actions that command them (ready for `env.step`), and the integrals live in
`integral_error` f64[n,6].  `DesiredState` is the namedtuple of `PID Controller/trajGen3D.py:13`.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

import torch

from ._cabi import QsPidGains, check

DesiredState = namedtuple("DesiredState", "pos vel acc yaw yawdot", defaults=(None, None, None, None, None))
KEYS = ("x", "y", "z", "phi", "theta", "psi")


def min_jerk_state(p0: torch.Tensor, p1: torch.Tensor, t: torch.Tensor, duration: float, yaw=None) -> DesiredState:
    """Rest-to-rest quintic from p0 [n,3] to p1 [n,3], sampled at t [n] seconds: the DesiredState (pos, vel, acc) a trajectory
    generator hands to the controller.  Stand-in for the reference's minimum-snap generator (`PID Controller/trajGen3D.py`,
    out of scope): with the reference's gains the controller needs a smooth reference -- a far step target saturates the
    thrust clamp and the attitude loop loses authority."""
    tau = (t / duration).clamp(0.0, 1.0).unsqueeze(1).to(torch.float64)
    s = 10 * tau ** 3 - 15 * tau ** 4 + 6 * tau ** 5
    sd = (30 * tau ** 2 - 60 * tau ** 3 + 30 * tau ** 4) / duration
    sdd = (60 * tau - 180 * tau ** 2 + 120 * tau ** 3) / duration ** 2
    d = (p1 - p0).to(torch.float64)
    return DesiredState(p0.to(torch.float64) + d * s, d * sd, d * sdd, yaw, None)


class BatchedPID:
    def __init__(self, env, gains: dict | None = None):
        """gains: optional {"kp"|"kd"|"ki": 6 values in the order x, y, z, phi, theta, psi, "max_integral": float};
        defaults are the reference's."""
        self.env = env
        self.lib = env.lib
        self.gains = QsPidGains()
        self.lib.qs_pid_default_gains(C.byref(self.gains))
        for name, val in (gains or {}).items():
            if name == "max_integral":
                self.gains.max_integral = float(val)
            elif name in ("kp", "kd", "ki"):
                getattr(self.gains, name)[:] = [float(v) for v in val]
            else:
                raise ValueError(f"unknown gain group {name!r}")
        self.integral_error = torch.zeros((env.n_envs, 6), dtype=torch.float64, device=env.device)
        self._wrench = torch.empty((env.n_envs, 4), dtype=torch.float64, device=env.device)
        self._actions = torch.empty((env.n_envs, 4), dtype=torch.float32, device=env.device)

    def reset_integral(self, mask: torch.Tensor | None = None) -> None:
        """Zero the integrals (of the masked envs) -- e.g. with StepOut.done after an auto-reset."""
        if mask is None:
            self.integral_error.zero_()
        else:
            self.integral_error.masked_fill_(mask.to(self.integral_error.device).bool().unsqueeze(1), 0.0)

    def _ptr(self, t, shape):
        if t is None:
            return None, None
        t = torch.as_tensor(t, dtype=torch.float64, device=self.env.device)
        t = t.expand(shape).contiguous() if tuple(t.shape) != tuple(shape) else t.contiguous()
        return C.c_void_p(t.data_ptr()), t

    def _call(self, des_state, dt, wrench, actions, clip):
        n = self.env.n_envs
        des_state = des_state or DesiredState()
        keep = []
        ptrs = []
        for val, shape in ((des_state.pos, (n, 3)), (des_state.vel, (n, 3)), (des_state.acc, (n, 3)), (des_state.yaw, (n,)),
                           (des_state.yawdot, (n,))):
            p, t = self._ptr(val, shape)
            ptrs.append(p)
            keep.append(t)
        st = C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        rc = self.lib.qs_pid_run(self.env._h, C.byref(self.gains), float(dt), *ptrs, C.c_void_p(self.integral_error.data_ptr()),
                                 C.c_void_p(wrench.data_ptr()) if wrench is not None else None,
                                 C.c_void_p(actions.data_ptr()) if actions is not None else None, int(bool(clip)), st)
        check(self.lib, self.env._h, rc, "qs_pid_run")
        if any(t is not None for t in keep):
            torch.cuda.current_stream(self.env.device).synchronize()      # temporaries must outlive the kernel

    def run(self, des_state: DesiredState | None = None, dt: float | None = None):
        """(F f64[n], M f64[n,3]) like pid_controller.run; des_state None = hover at each env's current waypoint."""
        self._call(des_state, dt if dt is not None else self.env.cfg.dt, self._wrench, None, False)
        return self._wrench[:, 0], self._wrench[:, 1:4]

    def actions(self, des_state: DesiredState | None = None, dt: float | None = None, clip: bool = True) -> torch.Tensor:
        """float32[n,4] env actions commanding the PID wrench (clipped to the action box like SB3 does before env.step)."""
        self._call(des_state, dt if dt is not None else self.env.cfg.dt, None, self._actions, clip)
        return self._actions
