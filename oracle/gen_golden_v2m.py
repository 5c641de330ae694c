"""TEST INFRASTRUCTURE ONLY -- golden resets for v2 with 2-3 waypoints (tests/golden/reset_v2m.npz).

The reference ships `self.num_waypoints = 1` (initial-implementation-v2/rl_env_scaledObs.py:47) and keeps the alternative
`#self.num_waypoints = np.random.randint(2, 4)` commented out one line above (:46).  This script loads the reference module from
its read-only source with exactly that one switch flipped IN MEMORY (line 46 uncommented, line 47 dropped -- nothing is written
anywhere, nothing else is changed), replays `reset()` on scripted blocks of unit uniforms (oracle/ref_harness.ScriptedRandom)
and records uniforms -> internal state + observation.  The multi-waypoint STEP cases are already part of step_v2.npz (built by
state injection into the unmodified reference).  Container only:
    python oracle/gen_golden_v2m.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True

from oracle import ref_harness as rh  # noqa: E402
from oracle.gen_golden import pack_states  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reset_v2m.npz")


def load_v2_with_line46():
    root = os.path.join(rh.REFERENCE_ROOT, "initial-implementation-v2")
    src = open(os.path.join(root, "rl_env_scaledObs.py")).read()
    off, on = "#self.num_waypoints = np.random.randint(2, 4)", "self.num_waypoints = np.random.randint(2, 4)"
    fixed = "self.num_waypoints =  1"
    assert src.count(off) == 1 and src.count(fixed) == 1, "reference source changed"
    src = src.replace(off, on).replace(fixed, "pass")
    rh._install_gymnasium_stub()
    for m in [m for m in sys.modules if m.split(".")[0] in ("simul_files", "utils2")]:
        del sys.modules[m]
    sys.path.insert(0, root)
    try:
        mod = types.ModuleType("_qs_reference_v2m")
        mod.__file__ = os.path.join(root, "rl_env_scaledObs.py")
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    finally:
        sys.path.remove(root)
        for m in [m for m in sys.modules if m.split(".")[0] in ("simul_files", "utils2")]:
            del sys.modules[m]
    return mod


def main(n=160):
    mod = load_v2_with_line46()
    rng = np.random.default_rng(46)
    U = rng.random((n, 18))
    # slot 4 = randint(2,4); slots 9/10 = trajectory mixture; slot 15 = curved axis.  Cover K x kind x axis, and the 0.3 boundary
    U[0:80, 4], U[80:160, 4] = 0.25, 0.75                            # K = 2 / K = 3
    for base in (0, 80):
        U[base:base + 16, 9] = 0.1                                   # linear
        U[base + 16:base + 52, 9], U[base + 16:base + 52, 10] = 0.5, 0.3    # curved
        U[base + 16:base + 28, 15], U[base + 28:base + 40, 15], U[base + 40:base + 52, 15] = 0.1, 0.5, 0.9
        U[base + 52:base + 68, 9], U[base + 52:base + 68, 10] = 0.9, 0.8    # helical
        U[base + 68, 9] = 0.3
    states, obs = [], []
    for i in range(n):
        env = mod.WaypointQuadEnv()
        with rh.quiet(), rh.scripted_random(U[i]) as sr:
            o, _ = env.reset()
        assert sr.pos <= 18
        states.append(rh.extract(env, "v2"))
        obs.append(o)
    out = {"uniforms": U, "obs": np.array(obs, dtype=np.float32)}
    for k, v in pack_states(states, "v2").items():
        out[k] = v
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "n_wp histogram", np.bincount(out["n_wp"]))


if __name__ == "__main__":
    main()
