"""WaypointQuadEnv -- the reference's gymnasium single-env surface on top of one GPU env slot.

Same constructor / `reset(seed=None) -> (obs, {})` / `step(action) -> (obs, reward, terminated, truncated, info)` and the
public attributes the reference's callers read:
    env.quadcopter.{state, position(), velocity(), omega(), attitude(), world_frame()}
    env.waypoint_list, env.current_waypoint, env.waypoint_index, env.final_yaw, env.F, env.M, env.dt
(initial-implementation-v2/runsim_scaledObs.py:29,60-66,90; initial-implementation-v2/simul_files/quadPlot.py:299-326;
class: initial-implementation-v2/rl_env_scaledObs.py:9, initial-implementation-v1/rl_env_scaledObs.py:8, rl_env.py).

It is the compatibility path for `runsim_*.py` / `evaluate_policy`-style single-env loops: every step is one kernel launch
for one env plus a 200-byte state read-back, so it is latency-, not throughput-oriented.  Use QuadVecEnv / BatchedQuadEnv for
batches.  `precision="f64", integrator="lsoda"` (default here) reproduces the reference step to 1e-9.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .quad_constants import QUAD
from .batched_env import BatchedQuadEnv
from .vec_env import _box


class QuadcopterView:
    """Read-only stand-in for simul_files.model.quadcopter.Quadcopter (quadcopter.py:25-64) fed from the GPU state."""

    def __init__(self):
        self.state = np.zeros(13)
        self.state[6] = 1.0

    def position(self):
        return self.state[0:3]

    def velocity(self):
        return self.state[3:6]

    def omega(self):
        return self.state[10:13]

    def rotation_matrix(self) -> np.ndarray:
        """Rotation matrix of the normalised quaternion (utils/quaternion.py:60-77, closed form)."""
        w, x, y, z = self.state[6:10] / np.linalg.norm(self.state[6:10])
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    def attitude(self):
        """RotToRPY of the rotation matrix (utils/utils.py:11-15)."""
        R = self.rotation_matrix()
        phi = math.asin(R[1, 2])
        theta = math.atan2(-R[0, 2] / math.cos(phi), R[2, 2] / math.cos(phi))
        psi = math.atan2(-R[1, 0] / math.cos(phi), R[1, 1] / math.cos(phi))
        return phi, theta, psi

    def world_frame(self) -> np.ndarray:
        """3x6 world coordinates of the four motors, the origin and the hub top (quadcopter.py:40-51)."""
        wHb = np.r_[np.c_[self.rotation_matrix(), self.state[0:3]], np.array([[0, 0, 0, 1]])]
        return wHb.dot(QUAD.body_frame.T)[0:3]


class WaypointQuadEnv:
    metadata = {"render_modes": []}

    def __init__(self, env_version: int = 2, obs_scaled: bool = True, precision: str = "f64", integrator: str = "lsoda",
                 substeps: int = 1, device: int | None = None, seed: int = 0, v2_random_waypoints: bool = False):
        self._sim = BatchedQuadEnv(1, env_version=env_version, precision=precision, integrator=integrator, substeps=substeps,
                                   obs_scaled=obs_scaled, auto_reset=False, device=device, seed=seed,
                                   v2_random_waypoints=v2_random_waypoints)
        d = self._sim.obs_dim
        self.observation_space = _box(-np.inf, np.inf, (d,), np.float32)
        self.action_space = _box(np.array([0, -1, -1, -1], dtype=np.float32), np.array([2.0, 1, 1, 1], dtype=np.float32))
        self.env_version = env_version
        self.quadcopter = None
        self.current_waypoint = None
        self.waypoint_list = []
        self.waypoint_index = 0
        self.dt = 1.0 / 200.0
        self.last_distance = None
        self.final_yaw = None
        self.final_waypoint_reached = None
        self.counter = None
        self.counter_activated = None
        self.counter_limit = 500 if env_version == 2 else None
        self.max_episode_steps = 2000 if env_version == 2 else 1200
        self.current_step = 0
        self.F = None
        self.M = None
        self._act = torch.zeros((1, 4), dtype=torch.float32, device=self._sim.device)
        self._first = True

    def _pull(self):
        st = {k: v.cpu().numpy() for k, v in self._sim.get_state().items()}
        if self.quadcopter is None:
            self.quadcopter = QuadcopterView()
        self.quadcopter.state = st["y"][0].copy()
        nwp = int(st["n_wp"][0])
        self.num_waypoints = nwp
        self.waypoint_list = [st["wp_list"][0, j].copy() for j in range(nwp)]
        self.waypoint_index = int(st["wp_index"][0])
        self.current_waypoint = self.waypoint_list[min(self.waypoint_index, nwp - 1)]
        ld = float(st["last_distance"][0])
        self.last_distance = None if math.isnan(ld) else ld
        self.current_step = int(st["current_step"][0])
        if self.env_version == 2:
            self.final_yaw = float(st["final_yaw"][0])
            self.counter = int(st["counter"][0])
            self.final_waypoint_reached = bool(st["final_reached"][0])
            self.counter_activated = self.final_waypoint_reached

    def reset(self, seed=None, options=None):
        if self._first:
            obs = self._sim.reset()
            self._first = False
        else:
            obs = self._sim.reset(torch.ones(1, dtype=torch.uint8))
        self._pull()
        return obs[0].cpu().numpy().copy(), {}

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(4)
        self.F = a[0] * np.float32(QUAD.mass) * np.float32(QUAD.gravity)     # float32 arithmetic, like the reference under NumPy >= 2
        self.M = a[1:4] * np.float32(0.1)
        self._act.copy_(torch.from_numpy(a).reshape(1, 4))
        out = self._sim.step(self._act)
        f = int(out.flags[0])
        self._pull()
        info = {}
        if f & 0x04:
            info = {"success": True, "stopped": bool(f & 0x08)}
        elif f & 0x10:
            info = {"success": False, "crashed": True}
        elif f & 0x20:
            info = {"success": False, "out_of_bounds": True}
        return out.obs[0].cpu().numpy().copy(), float(out.reward[0]), bool(f & 0x01), bool(f & 0x02), info

    def close(self):
        self._sim.close()
