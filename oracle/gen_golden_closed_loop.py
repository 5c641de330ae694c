"""TEST INFRASTRUCTURE ONLY -- closed-loop golden episodes from the UNMODIFIED reference v2 env driven by the
reference's own best checkpoint.

Runs only in the build container (needs /root/reference):
    python oracle/gen_golden_closed_loop.py
What the bench's rollout workload runs on the GPU is exactly this loop (v2 env + the MlpPolicy of
checkpoints_from_8_6M/ppo_model_2300000_steps.zip), so this pins it episode by episode:

  * the reference `WaypointQuadEnv` (initial-implementation-v2/rl_env_scaledObs.py:9, LSODA, float64) is reset on the unit
    uniforms the CUDA kernel draws from Philox for (seed, global env id e, episode 0), e = 0..N-1 -- generated here with the
    product's own Philox code compiled for the host (tests/harness_util.py) and fed through a scripted np.random;
  * actions are `model.predict(obs, deterministic=True)` as in runsim_scaledObs.py:54-60: torch float32 forward of the
    shipped actor, clipped to the action box (SB3 clips in predict());
  * each episode runs until terminated or truncated.

Recorded per episode: length, return, final flags / info bits, the step of the first arrival, the final state; for the first
TRAJ episodes also the full state / reward / action trajectory (closed-loop drift check).
Output: tests/golden/closed_loop_v2.npz
"""
from __future__ import annotations

import io
import os
import sys
import time
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ZIP = "initial-implementation-v2/checkpoints_from_8_6M/ppo_model_2300000_steps.zip"
SEED = 20260
N_EPISODES = 64
TRAJ = 8
MAX_LEN = 2100


def info_bits(info: dict) -> int:
    return (1 if info.get("success", False) else 0) | (2 if info.get("stopped", False) else 0) | \
           (4 if info.get("crashed", False) else 0) | (8 if info.get("out_of_bounds", False) else 0)


def main():
    import torch

    import rl_aerial_manipulator_b200 as qsim
    from harness_util import HostHarness

    hh = HostHarness(qsim.make_config(env_version=2, precision="f64", integrator="lsoda", seed=SEED))
    z = zipfile.ZipFile(os.path.join(rh.REFERENCE_ROOT, ZIP))
    sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True, map_location="cpu")

    def predict(obs: np.ndarray) -> np.ndarray:
        x = torch.from_numpy(obs[None].astype(np.float32))
        for i in (0, 2, 4):
            x = torch.tanh(x @ sd[f"mlp_extractor.policy_net.{i}.weight"].T + sd[f"mlp_extractor.policy_net.{i}.bias"])
        a = (x @ sd["action_net.weight"].T + sd["action_net.bias"])[0].numpy()
        return np.clip(a, [0, -1, -1, -1], [2, 1, 1, 1]).astype(np.float32)

    env = rh.make_env("v2")
    U = np.stack([hh.uniforms(SEED, e, 0) for e in range(N_EPISODES)])
    length = np.zeros(N_EPISODES, dtype=np.int32)
    ret = np.zeros(N_EPISODES)
    term = np.zeros(N_EPISODES, dtype=bool)
    trunc = np.zeros(N_EPISODES, dtype=bool)
    info = np.zeros(N_EPISODES, dtype=np.uint8)
    arrival = np.full(N_EPISODES, -1, dtype=np.int32)
    y0 = np.zeros((N_EPISODES, 13))
    y_final = np.zeros((N_EPISODES, 13))
    wp = np.zeros((N_EPISODES, 3))
    final_yaw = np.zeros(N_EPISODES)
    traj_y = np.zeros((TRAJ, MAX_LEN, 13))
    traj_r = np.zeros((TRAJ, MAX_LEN))
    traj_a = np.zeros((TRAJ, MAX_LEN, 4), dtype=np.float32)
    t0 = time.time()
    for e in range(N_EPISODES):
        with rh.quiet(), rh.scripted_random(U[e]):
            obs, _ = env.reset()
        y0[e] = env.quadcopter.state
        wp[e] = env.current_waypoint
        final_yaw[e] = env.final_yaw
        for t in range(MAX_LEN):
            a = predict(obs)
            with rh.quiet(), np.errstate(all="ignore"):
                obs, r, te, tr, inf = env.step(a.copy())
            ret[e] += r
            if e < TRAJ:
                traj_y[e, t], traj_r[e, t], traj_a[e, t] = env.quadcopter.state, r, a
            if arrival[e] < 0 and env.final_waypoint_reached:
                arrival[e] = t
            if te or tr:
                length[e], term[e], trunc[e], info[e] = t + 1, te, tr, info_bits(inf)
                break
        else:
            raise RuntimeError("episode did not end")
        y_final[e] = env.quadcopter.state
        print(f"episode {e}: len {length[e]} return {ret[e]:.3f} term {te} trunc {tr} info {info[e]} arrival {arrival[e]}  ({time.time() - t0:.0f} s)")
    L = int(length[:TRAJ].max())
    np.savez_compressed(os.path.join(OUT, "closed_loop_v2.npz"), seed=np.int64(SEED), uniforms=U, length=length, ep_return=ret,
                        terminated=term, truncated=trunc, info=info, arrival=arrival, y0=y0, y_final=y_final, waypoint=wp,
                        final_yaw=final_yaw, traj_y=traj_y[:, :L], traj_reward=traj_r[:, :L], traj_action=traj_a[:, :L],
                        source=np.array(ZIP))
    print(f"closed_loop_v2: {N_EPISODES} episodes, success {int((info & 1).astype(bool).sum())}, mean length {length.mean():.1f}, "
          f"mean return {ret.mean():.1f}")


if __name__ == "__main__":
    main()
