"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference env.

Imports `WaypointQuadEnv` straight from /root/reference (read-only) so golden
vectors can be generated from the reference's own Python/NumPy/SciPy step.
/root/reference exists only in the build container, never on the GPU box, so
nothing in `-m gpu` tests, `smoke()` or `bench.py` may import this module; it
is used by `oracle/gen_golden.py` and by the container-only tests that are
skipped when the reference tree is absent.

The reference needs `gymnasium`, which is not installed here and cannot be
(no network).  The stub below provides the two names the reference touches:
`gymnasium.Env` (whose `reset(seed=...)` the reference calls via `super()`)
and `gymnasium.spaces.Box`.

Reference entry points loaded (file:line):
  initial-implementation-v2/rl_env_scaledObs.py:9   WaypointQuadEnv (20-D obs)
  initial-implementation-v1/rl_env_scaledObs.py:8   WaypointQuadEnv (17-D scaled obs)
  initial-implementation-v1/rl_env.py               WaypointQuadEnv (17-D raw obs)
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("QS_REFERENCE_ROOT", "/root/reference")

_VARIANTS = {
    "v2": ("initial-implementation-v2", "rl_env_scaledObs.py"),
    "v1": ("initial-implementation-v1", "rl_env_scaledObs.py"),
    "v1_raw": ("initial-implementation-v1", "rl_env.py"),
}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "initial-implementation-v2"))


def _install_gymnasium_stub() -> None:
    if "gymnasium" in sys.modules:
        return
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Env:  # minimal gymnasium.Env: reset() only seeds nothing
        def reset(self, seed=None, options=None):
            return None

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    gym.Env = Env
    spaces.Box = Box
    gym.spaces = spaces
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces


_loaded: dict[str, types.ModuleType] = {}


def load_reference(variant: str) -> types.ModuleType:
    """Return the reference module for `variant` ("v1", "v1_raw", "v2").

    v1 and v2 both use the top-level names `simul_files` / `utils2`, so each
    variant is imported with its own directory first on sys.path and its
    support packages are purged from sys.modules afterwards (the env module
    keeps its own references to them).
    """
    if variant in _loaded:
        return _loaded[variant]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    _install_gymnasium_stub()
    subdir, fname = _VARIANTS[variant]
    root = os.path.join(REFERENCE_ROOT, subdir)
    purge = [m for m in sys.modules if m.split(".")[0] in ("simul_files", "utils2")]
    for m in purge:
        del sys.modules[m]
    sys.path.insert(0, root)
    try:
        spec = importlib.util.spec_from_file_location(f"_qs_reference_{variant}", os.path.join(root, fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(root)
        for m in [m for m in sys.modules if m.split(".")[0] in ("simul_files", "utils2")]:
            del sys.modules[m]
    _loaded[variant] = mod
    return mod


@contextlib.contextmanager
def quiet():
    """The v2 trajectory generators and the hold-phase exit print(); mute them."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class ScriptedRandom:
    """Stand-in for the three np.random calls the reference reset makes.

    Each call consumes unit uniforms u in [0,1) from `stream` in order:
      uniform(lo, hi[, size]) -> lo + (hi - lo) * u      (NumPy legacy definition)
      rand()                  -> u
      randint(lo, hi)         -> lo + floor(u * (hi - lo))
    so a reset of the reference can be replayed on exactly the uniforms the
    CUDA kernel drew from Philox (see csrc/quadsim_reset.cuh).
    """

    def __init__(self, stream):
        self.stream = [float(u) for u in stream]
        self.pos = 0

    def _next(self) -> float:
        u = self.stream[self.pos]
        self.pos += 1
        return u

    def uniform(self, low=0.0, high=1.0, size=None):
        if size is None:
            return low + (high - low) * self._next()
        n = int(np.prod(size))
        return np.array([low + (high - low) * self._next() for _ in range(n)]).reshape(size)

    def rand(self):
        return self._next()

    def randint(self, low, high):
        return low + int(np.floor(self._next() * (high - low)))


@contextlib.contextmanager
def scripted_random(stream):
    rng = ScriptedRandom(stream)
    saved = (np.random.uniform, np.random.rand, np.random.randint)
    np.random.uniform, np.random.rand, np.random.randint = rng.uniform, rng.rand, rng.randint
    try:
        yield rng
    finally:
        np.random.uniform, np.random.rand, np.random.randint = saved


def make_env(variant: str):
    return load_reference(variant).WaypointQuadEnv()


def inject(env, variant: str, s: dict) -> None:
    """Put a reference env into an arbitrary internal state (dict of plain values)."""
    mod = load_reference(variant)
    quad = mod.Quadcopter(np.zeros(3), (0.0, 0.0, 0.0))
    quad.state = np.array(s["y"], dtype=np.float64).copy()
    env.quadcopter = quad
    env.waypoint_list = [np.array(w, dtype=np.float64) for w in s["wp_list"]]
    env.num_waypoints = len(env.waypoint_list)
    env.waypoint_index = int(s["wp_index"])
    env.current_waypoint = np.array(s["cur_wp"], dtype=np.float64)
    ld = s["last_distance"]
    env.last_distance = None if (ld is None or np.isnan(ld)) else float(ld)
    env.current_step = int(s["current_step"])
    if variant == "v2":
        env.max_episode_steps = 2000
        env.counter_limit = 500
        env.counter = int(s["counter"])
        env.final_waypoint_reached = bool(s["final_reached"])
        env.counter_activated = bool(s["final_reached"])
        env.final_yaw = float(s["final_yaw"])
    else:
        env.max_episode_steps = 1200


def extract(env, variant: str) -> dict:
    out = {
        "y": np.array(env.quadcopter.state, dtype=np.float64).copy(),
        "wp_list": np.array(env.waypoint_list, dtype=np.float64).copy(),
        "wp_index": int(env.waypoint_index),
        "cur_wp": np.array(env.current_waypoint, dtype=np.float64).copy(),
        "last_distance": np.nan if env.last_distance is None else float(env.last_distance),
        "current_step": int(env.current_step),
    }
    if variant == "v2":
        out["counter"] = int(env.counter)
        out["final_reached"] = bool(env.final_waypoint_reached)
        out["final_yaw"] = float(env.final_yaw)
    return out
