"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference):
    python oracle/gen_golden.py
It drives the reference's own `WaypointQuadEnv` (v1, v1_raw, v2) through
`oracle/ref_harness.py` (state injection + scripted np.random) and records

  step_<variant>.npz   single-step known answers: internal state before, float32 action,
                       internal state after, obs, reward, terminated, truncated, info bits, and the
                       LSODA counters (nst, nfe, nqu, hu, tcur) of the very same odeint call --
                       random states plus one block per branch of the step state machine
  reset_<variant>.npz  reset known answers: block of unit uniforms -> internal state + obs
  traj_<variant>.npz   multi-env trajectories with DummyVecEnv-style auto-reset
  seed0_v2.npz         np.random.seed(0); reset(); 3 steps -- the SURVEY appendix-A vector
  policy_<v>.npz       SB3 MlpPolicy weights (from the shipped zips) + torch forward outputs
  vecnorm_v1.npz       obs_rms / ret_rms snapshot from initial-implementation-v1/vec_normalize.pkl

Everything is float64 unless the reference itself produces float32 (obs, actions).
"""
from __future__ import annotations

import io
import os
import pickle
import sys
import zipfile

import numpy as np
from scipy.integrate import odeint

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
INFO_BITS = {"success": 1, "stopped": 2, "crashed": 4, "out_of_bounds": 8}
MAXWP = 3


def info_bits(info: dict) -> int:
    b = 0
    if info.get("success", False):
        b |= 1
    if info.get("stopped", False):
        b |= 2
    if info.get("crashed", False):
        b |= 4
    if info.get("out_of_bounds", False):
        b |= 8
    return b


def rand_quat(rng, spread):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = rng.uniform(-spread, spread)
    return np.array([np.cos(ang / 2), *(np.sin(ang / 2) * ax)])


def random_state(rng, variant, spread=1.0):
    nwp = int(rng.integers(1, 3)) if variant != "v2" else 1
    y = np.zeros(13)
    y[0:3] = rng.uniform(-2, 2, 3)
    y[2] = rng.uniform(0.5, 3.0)
    y[3:6] = rng.normal(size=3) * 0.8
    y[6:10] = rand_quat(rng, spread)
    y[10:13] = rng.normal(size=3) * 1.5
    wps = [np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1, 3)]) for _ in range(nwp)]
    wp_index = int(rng.integers(0, nwp))
    d = np.linalg.norm(y[0:3] - wps[wp_index])
    s = dict(y=y, wp_list=wps, wp_index=wp_index, cur_wp=wps[wp_index].copy(),
             last_distance=d + rng.normal() * 0.01, current_step=int(rng.integers(0, 1000)),
             counter=0, final_reached=False, final_yaw=rng.uniform(-np.pi, np.pi))
    return s


def random_action(rng):
    return np.array([rng.uniform(0, 2), rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)], dtype=np.float32)


def hover_near(rng, s, offset):
    """Place the vehicle `offset` metres from its current waypoint, slow and level."""
    dirn = rng.normal(size=3)
    dirn /= np.linalg.norm(dirn)
    s["y"][0:3] = s["cur_wp"] + dirn * offset
    s["y"][3:6] = rng.normal(size=3) * 0.05
    s["y"][6:10] = rand_quat(rng, 0.1)
    s["y"][10:13] = rng.normal(size=3) * 0.05
    s["last_distance"] = offset + 0.001


def build_cases(variant, rng):
    """List of (name, state-dict, action) covering every branch of the step state machine."""
    cases = []
    add = lambda name, s, a: cases.append((name, s, np.asarray(a, dtype=np.float32)))
    hover = np.array([1, 0, 0, 0], dtype=np.float32)
    for i in range(160):
        add("random", random_state(rng, variant, spread=[0.3, 1.0, 3.1][i % 3]), random_action(rng))
    for i in range(12):
        s = random_state(rng, variant)
        s["last_distance"] = None
        s["current_step"] = 0
        add("first_step", s, random_action(rng))
    for a in ([2, 1, 1, 1], [0, -1, 0.5, -0.25], [2, -1, -1, -1], [0, 0, 0, 0], [2, 0, 0, 0], [0, 1, -1, 1],
              [1.2, 0.3, -0.2, 0.1]):
        add("saturated", random_state(rng, variant, 0.3), a)
    for i in range(10):  # crash, falling and rising
        s = random_state(rng, variant, 0.3)
        s["y"][2] = rng.uniform(-0.05, 0.09)
        s["y"][5] = -abs(s["y"][5]) if i % 2 == 0 else abs(s["y"][5]) + 0.5
        add("crash", s, random_action(rng))
    for i in range(8):
        s = random_state(rng, variant, 0.3)
        s["y"][0:3] = s["y"][0:3] / np.linalg.norm(s["y"][0:3]) * rng.uniform(10.2, 12)
        s["y"][2] = abs(s["y"][2]) + 0.5
        add("out_of_bounds", s, random_action(rng))
    limit = 2000 if variant == "v2" else 1200
    for st in (limit - 2, limit - 1, limit, limit + 1):
        for _ in range(2):
            s = random_state(rng, variant, 0.3)
            s["current_step"] = st
            add("truncation", s, random_action(rng))
    if variant == "v2":
        for i in range(16):  # first arrival at the final waypoint
            s = random_state(rng, variant)
            hover_near(rng, s, rng.uniform(0.01, 0.09))
            if i % 4 == 1:
                s["y"][3:6] = rng.normal(size=3) * 1.5  # |v| >= 1 -> no stopping bonus
            if i % 4 == 2:
                s["y"][6:10] = rand_quat(rng, 3.1)
            add("final_first", s, hover if i % 2 else random_action(rng))
        for c in (0, 1, 250, 499, 500, 501, 502, 900):  # hold phase around counter_limit
            for j in range(2):
                s = random_state(rng, variant)
                hover_near(rng, s, rng.uniform(0.01, 0.09))
                if j == 1:
                    s["y"][6:10] = rand_quat(rng, 1.2)  # |roll|,|pitch| > 0.2 branches
                s["final_reached"], s["counter"], s["wp_index"] = True, c, 1
                add("hold", s, hover if j == 0 else random_action(rng))
        for c in (3, 500, 700):  # left the 0.1 m ball after arrival: counter keeps running
            s = random_state(rng, variant)
            hover_near(rng, s, rng.uniform(0.15, 0.6))
            s["final_reached"], s["counter"], s["wp_index"] = True, c, 1
            add("hold_left", s, random_action(rng))
        for i in range(6):  # reached + crash / reached + truncation combinations
            s = random_state(rng, variant)
            s["cur_wp"] = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), 0.12])
            s["wp_list"] = [s["cur_wp"].copy()]
            hover_near(rng, s, 0.03)
            s["y"][2] = 0.095
            s["current_step"] = 2000 if i % 2 else 10
            s["final_reached"] = i >= 3
            s["wp_index"] = 1 if i >= 3 else 0
            add("reach_low", s, random_action(rng))
        for k in (2, 3):  # multi-waypoint lists (the commented `randint(2,4)` alternative, :46)
            for idx in range(k):
                s = random_state(rng, variant)
                s["wp_list"] = [np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1, 3)]) for _ in range(k)]
                s["wp_index"], s["cur_wp"] = idx, s["wp_list"][idx].copy()
                hover_near(rng, s, rng.uniform(0.01, 0.09))
                add("multi_wp_reach", s, random_action(rng))
                s2 = random_state(rng, variant)
                s2["wp_list"] = [w.copy() for w in s["wp_list"]]
                s2["wp_index"], s2["cur_wp"] = idx, s["wp_list"][idx].copy()
                s2["last_distance"] = np.linalg.norm(s2["y"][0:3] - s2["cur_wp"])
                add("multi_wp_far", s2, random_action(rng))
    else:
        for i in range(16):  # approach bonus / penalty inside 0.5 m
            s = random_state(rng, variant, 0.3)
            hover_near(rng, s, rng.uniform(0.15, 0.45))
            to_wp = s["cur_wp"] - s["y"][0:3]
            to_wp /= np.linalg.norm(to_wp)
            s["y"][3:6] = to_wp * [0.5, -0.5, 0.05, 0.1][i % 4]
            add("approach", s, hover)
        for nwp in (1, 2):
            for idx in range(nwp):
                for _ in range(4):
                    s = random_state(rng, variant, 0.3)
                    s["wp_list"] = [np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1, 3)]) for _ in range(nwp)]
                    s["wp_index"], s["cur_wp"] = idx, s["wp_list"][idx].copy()
                    hover_near(rng, s, rng.uniform(0.01, 0.09))
                    if _ % 2:
                        s["y"][3:6] *= 10
                        s["y"][10:13] *= 10
                    s["current_step"] = [5, 1200][_ // 2]
                    add("reach", s, hover if _ % 2 else random_action(rng))
        s = random_state(rng, variant, 0.3)  # duplicate waypoints: is_final via np.allclose on values
        s["wp_list"] = [s["wp_list"][0].copy(), s["wp_list"][0].copy() + 5e-9]
        s["wp_index"], s["cur_wp"] = 0, s["wp_list"][0].copy()
        add("allclose_final", s, random_action(rng))
    return cases


def ref_post_update(variant, s, action):
    """Post-update dynamics state of the reference for (s, action) without touching the env logic."""
    env = rh.make_env(variant)
    rh.inject(env, variant, s)
    mod = rh.load_reference(variant)
    F = action[0] * mod.params.mass * mod.params.g
    M = action[1:4] * 0.1
    env.quadcopter.update(env.dt, F, M.reshape(-1, 1))
    return env.quadcopter.state.copy()


def lsoda_stats(variant, s, action):
    """Counters of the odeint call the reference makes for (s, action): same f, same args."""
    mod = rh.load_reference(variant)
    P = mod.params
    quad = mod.Quadcopter(np.zeros(3), (0, 0, 0))
    F = action[0] * P.mass * P.g
    M = (action[1:4] * 0.1).reshape(-1, 1)
    t = P.invA.dot(np.r_[np.array([[F]]), M])
    tc = np.maximum(np.minimum(t, P.maxF / 4), P.minF / 4)
    Fc = np.sum(tc)
    Mc = P.A[1:].dot(tc)
    sol, info = odeint(quad.state_dot, np.array(s["y"], dtype=np.float64), [0, 1.0 / 200.0], args=(Fc, Mc), full_output=True)
    return dict(nst=info["nst"][0], nfe=info["nfe"][0], nqu=info["nqu"][0], hu=info["hu"][0], tcur=info["tcur"][0],
                mused=info["mused"][0], F_clamped=float(Fc), M_clamped=Mc.reshape(3), y_raw=sol[1])


def pack_states(states, variant):
    n = len(states)
    out = dict(y=np.zeros((n, 13)), wp_list=np.zeros((n, MAXWP, 3)), n_wp=np.zeros(n, dtype=np.int64),
               wp_index=np.zeros(n, dtype=np.int64), cur_wp=np.zeros((n, 3)), last_distance=np.full(n, np.nan),
               current_step=np.zeros(n, dtype=np.int64), counter=np.zeros(n, dtype=np.int64),
               final_reached=np.zeros(n, dtype=bool), final_yaw=np.zeros(n))
    for i, s in enumerate(states):
        out["y"][i] = s["y"]
        wl = np.asarray(s["wp_list"], dtype=np.float64).reshape(-1, 3)
        out["n_wp"][i] = wl.shape[0]
        out["wp_list"][i, : wl.shape[0]] = wl
        out["wp_index"][i] = s["wp_index"]
        out["cur_wp"][i] = s["cur_wp"]
        ld = s["last_distance"]
        out["last_distance"][i] = np.nan if ld is None else ld
        out["current_step"][i] = s["current_step"]
        if variant == "v2":
            out["counter"][i] = s["counter"]
            out["final_reached"][i] = s["final_reached"]
            out["final_yaw"][i] = s["final_yaw"]
    return out


def gen_step(variant, seed):
    rng = np.random.default_rng(seed)
    cases = build_cases(variant, rng)
    if variant != "v2":
        # exactly-on-waypoint case: distance 0 -> unit direction NaN (v1 :100-110)
        s = random_state(rng, variant, 0.3)
        a = random_action(rng)
        s["wp_list"] = [s["wp_list"][0]]
        s["wp_index"] = 0
        post = ref_post_update(variant, s, a)
        s["cur_wp"] = post[0:3].copy()
        s["wp_list"] = [post[0:3].copy()]
        cases.append(("on_waypoint_nan", s, a))
    names, pre, post, acts, obs, rew, term, trunc, info, stats = [], [], [], [], [], [], [], [], [], []
    for name, s, a in cases:
        env = rh.make_env(variant)
        rh.inject(env, variant, s)
        st = lsoda_stats(variant, s, a)
        with rh.quiet(), np.errstate(all="ignore"):
            o, r, te, tr, inf = env.step(a.copy())
        names.append(name)
        pre.append(s)
        post.append(rh.extract(env, variant))
        acts.append(a)
        obs.append(o)
        rew.append(float(r))
        term.append(bool(te))
        trunc.append(bool(tr))
        info.append(info_bits(inf))
        stats.append(st)
    out = {"case": np.array(names), "action": np.array(acts, dtype=np.float32), "obs": np.array(obs, dtype=np.float32),
           "reward": np.array(rew), "terminated": np.array(term), "truncated": np.array(trunc),
           "info": np.array(info, dtype=np.uint8)}
    for k, v in pack_states(pre, variant).items():
        out["pre_" + k] = v
    for k, v in pack_states(post, variant).items():
        out["post_" + k] = v
    for k in ("nst", "nfe", "nqu", "hu", "tcur", "mused", "F_clamped", "M_clamped", "y_raw"):
        out["lsoda_" + k] = np.array([s[k] for s in stats])
    np.savez_compressed(os.path.join(OUT, f"step_{variant}.npz"), **out)
    kinds, counts = np.unique(out["case"], return_counts=True)
    print(f"step_{variant}: {len(names)} cases", dict(zip(kinds.tolist(), counts.tolist())),
          "terminated", int(out["terminated"].sum()), "truncated", int(out["truncated"].sum()))


def gen_reset(variant, seed, n=96):
    rng = np.random.default_rng(seed)
    U = rng.random((n, 16))
    # make sure every trajectory kind / axis / waypoint count shows up
    if variant == "v2":
        U[0:8, 8] = 0.1                       # linear
        U[8:32, 8], U[8:32, 9] = 0.5, 0.3     # curved
        U[8:16, 14], U[16:24, 14], U[24:32, 14] = 0.1, 0.5, 0.9  # randint slot for curved: z / y / x bump
        U[32:40, 8], U[32:40, 9] = 0.9, 0.8   # helical
        U[40, 8] = 0.3                         # boundary: 0.3 < 0.3 is False
    else:
        U[0:8, 4], U[8:16, 4] = 0.2, 0.7       # one / two waypoints
    states, obs = [], []
    for i in range(n):
        env = rh.make_env(variant)
        with rh.quiet(), rh.scripted_random(U[i]):
            o, _ = env.reset()
        states.append(rh.extract(env, variant))
        obs.append(o)
    out = {"uniforms": U, "obs": np.array(obs, dtype=np.float32)}
    for k, v in pack_states(states, variant).items():
        out[k] = v
    np.savez_compressed(os.path.join(OUT, f"reset_{variant}.npz"), **out)
    print(f"reset_{variant}: {n} cases")


def gen_traj(variant, seed, n_env=4, n_steps=600):
    """DummyVecEnv-style rollout of n_env reference envs; resets consume scripted uniform blocks."""
    rng = np.random.default_rng(seed)
    max_ep = 64
    U = rng.random((n_env, max_ep, 16))
    envs = [rh.make_env(variant) for _ in range(n_env)]
    episode = np.zeros(n_env, dtype=np.int64)
    D = 20 if variant == "v2" else 17
    obs0 = np.zeros((n_env, D), dtype=np.float32)
    for i, env in enumerate(envs):
        with rh.quiet(), rh.scripted_random(U[i, 0]):
            obs0[i], _ = env.reset()
    acts = np.zeros((n_steps, n_env, 4), dtype=np.float32)
    obs = np.zeros((n_steps, n_env, D), dtype=np.float32)
    tobs = np.zeros((n_steps, n_env, D), dtype=np.float32)
    rew = np.zeros((n_steps, n_env))
    term = np.zeros((n_steps, n_env), dtype=bool)
    trunc = np.zeros((n_steps, n_env), dtype=bool)
    info = np.zeros((n_steps, n_env), dtype=np.uint8)
    ys = np.zeros((n_steps, n_env, 13))
    for t in range(n_steps):
        for i, env in enumerate(envs):
            if i == 0:
                a = np.array([1, 0, 0, 0], dtype=np.float32)       # hover
            elif i == 1:
                a = random_action(rng)                               # uniform over the action box
            else:                                                    # crude P-controller: keeps episodes alive longer
                pos, vel = env.quadcopter.position(), env.quadcopter.velocity()
                e = env.current_waypoint - pos
                a = np.array([np.clip(1 + 0.6 * e[2] - 0.4 * vel[2], 0, 2), np.clip(-0.02 * e[1] + 0.02 * vel[1] - 0.05 * env.quadcopter.omega()[0], -1, 1),
                              np.clip(0.02 * e[0] - 0.02 * vel[0] - 0.05 * env.quadcopter.omega()[1], -1, 1), 0.0], dtype=np.float32)
                a += (rng.normal(size=4) * [0.05, 0.002, 0.002, 0.002]).astype(np.float32)
                a = np.clip(a, [0, -1, -1, -1], [2, 1, 1, 1]).astype(np.float32)
            acts[t, i] = a
            with rh.quiet(), np.errstate(all="ignore"):
                o, r, te, tr, inf = env.step(a.copy())
            tobs[t, i], rew[t, i], term[t, i], trunc[t, i], info[t, i] = o, r, te, tr, info_bits(inf)
            ys[t, i] = env.quadcopter.state
            if te or tr:
                episode[i] += 1
                with rh.quiet(), rh.scripted_random(U[i, episode[i]]):
                    o, _ = env.reset()
            obs[t, i] = o
    np.savez_compressed(os.path.join(OUT, f"traj_{variant}.npz"), uniforms=U, obs0=obs0, action=acts, obs=obs,
                        terminal_obs=tobs, reward=rew, terminated=term, truncated=trunc, info=info, y=ys)
    print(f"traj_{variant}: {n_env} envs x {n_steps} steps, episodes finished per env {episode.tolist()}")


def gen_seed0():
    """np.random.seed(0); reset(); three float32 actions -- with the real MT19937 draws."""
    env = rh.make_env("v2")
    np.random.seed(0)
    with rh.quiet():
        o0, _ = env.reset()
    s0 = rh.extract(env, "v2")
    acts = np.array([[1.2, 0.3, -0.2, 0.1], [2, 1, 1, 1], [0, -1, 0.5, -0.25]], dtype=np.float32)
    ys, obs, rew = [], [], []
    for a in acts:
        with rh.quiet():
            o, r, te, tr, inf = env.step(a.copy())
        ys.append(env.quadcopter.state.copy())
        obs.append(o)
        rew.append(r)
    np.savez_compressed(os.path.join(OUT, "seed0_v2.npz"), y0=s0["y"], cur_wp=s0["cur_wp"], final_yaw=s0["final_yaw"],
                        obs0=o0, action=acts, y=np.array(ys), obs=np.array(obs), reward=np.array(rew))
    print("seed0_v2: reset state", s0["y"][:3], "wp", s0["cur_wp"], "yaw", s0["final_yaw"])


def gen_policy(tag, zip_rel, obs_dim, seed):
    """MlpPolicy weights from a shipped SB3 zip + torch CPU forward on random observations."""
    import torch

    z = zipfile.ZipFile(os.path.join(rh.REFERENCE_ROOT, zip_rel))
    sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True, map_location="cpu")
    rng = np.random.default_rng(seed)
    obs = rng.normal(size=(64, obs_dim)).astype(np.float32)
    obs[:, 6] = 1.0

    def mlp(x, prefix, dt):
        for i in (0, 2, 4):
            w, b = sd[f"mlp_extractor.{prefix}.{i}.weight"].to(dt), sd[f"mlp_extractor.{prefix}.{i}.bias"].to(dt)
            x = torch.tanh(x @ w.T + b)
        return x

    out = {"obs": obs}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        x = torch.from_numpy(obs).to(dt)
        pi, vf = mlp(x, "policy_net", dt), mlp(x, "value_net", dt)
        out["mean_" + name] = (pi @ sd["action_net.weight"].to(dt).T + sd["action_net.bias"].to(dt)).numpy()
        out["value_" + name] = (vf @ sd["value_net.weight"].to(dt).T + sd["value_net.bias"].to(dt)).numpy()[:, 0]
    for k, v in sd.items():
        out["w." + k] = v.numpy()
    out["source"] = np.array(zip_rel)
    np.savez_compressed(os.path.join(OUT, f"policy_{tag}.npz"), **out)
    print(f"policy_{tag}: {sum(v.numel() for v in sd.values())} parameters from {zip_rel}")


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.__dict__.update(st)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith(("stable_baselines3", "gymnasium")):
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


def gen_vecnorm():
    with open(os.path.join(rh.REFERENCE_ROOT, "initial-implementation-v1", "vec_normalize.pkl"), "rb") as f:
        vn = _Unpickler(f).load()
    d = vn.__dict__
    o, r = d["obs_rms"].__dict__, d["ret_rms"].__dict__
    np.savez_compressed(os.path.join(OUT, "vecnorm_v1.npz"), obs_mean=o["mean"], obs_var=o["var"], obs_count=o["count"],
                        ret_mean=r["mean"], ret_var=r["var"], ret_count=r["count"], clip_obs=d["clip_obs"],
                        clip_reward=d["clip_reward"], gamma=d["gamma"], epsilon=d["epsilon"],
                        norm_obs=d["norm_obs"], norm_reward=d["norm_reward"])
    print("vecnorm_v1: count", o["count"], "clip_obs", d["clip_obs"], "gamma", d["gamma"], "eps", d["epsilon"])


def main():
    os.makedirs(OUT, exist_ok=True)
    for variant, seed in (("v2", 1001), ("v1", 1002), ("v1_raw", 1003)):
        gen_step(variant, seed)
        gen_reset(variant, seed + 10)
    gen_traj("v2", 2001)
    gen_traj("v1", 2002)
    gen_seed0()
    gen_policy("v2", "initial-implementation-v2/checkpoints_from_8_6M/ppo_model_2300000_steps.zip", 20, 3001)
    gen_policy("v1", "initial-implementation-v1/waypoint_controller_scaledObs_4M.zip", 17, 3002)
    gen_vecnorm()


if __name__ == "__main__":
    main()
