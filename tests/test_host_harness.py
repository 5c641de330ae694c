"""The product's device code (csrc/*.cuh), compiled for the host with g++, against the golden vectors of the
unmodified reference and against the oracle.  Runs on the CPU-only build box; the same checks run on the
B200 through the C-ABI in tests/test_gpu_parity.py.

Tolerances (north_star): float64 single step 1e-9 relative; float32 1e-4; flags bit-exact.
"""
import os

import numpy as np
import pytest

import rl_aerial_manipulator_b200 as qsim
from oracle import quad_oracle as qo
from harness_util import HostHarness, golden_state, flags_from_golden

VERS = {"v2": 2, "v1": 1, "v1_raw": 1}


@pytest.fixture(scope="module")
def hh():
    return HostHarness(qsim.make_config(env_version=2, precision="f64", integrator="lsoda", seed=1234))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_philox_known_answers(hh):
    """Random123 known-answer vectors for philox4x32-10."""
    np.testing.assert_array_equal(hh.philox([0, 0, 0, 0], [0, 0]), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8])
    np.testing.assert_array_equal(hh.philox([0xffffffff] * 4, [0xffffffff] * 2), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd])
    np.testing.assert_array_equal(hh.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]),
                                  [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])


def test_uniforms_are_uniform(hh):
    u = np.array([hh.uniforms(7, e, ep) for e in range(400) for ep in range(3)])
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    assert len(np.unique(u)) == u.size  # distinct (env, episode) -> distinct draws
    # different seed / env / episode all change the block
    assert not np.array_equal(hh.uniforms(7, 1, 0), hh.uniforms(8, 1, 0))
    assert not np.array_equal(hh.uniforms(7, 1, 0), hh.uniforms(7, 1, 1))
    assert not np.array_equal(hh.uniforms(7, 1, 0), hh.uniforms(7, 2**32 + 1, 0))


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_lsoda_port_matches_scipy_counters(hh, golden_dir, variant):
    """nst / nfe / nqu / hu / tcur and the raw result of the reference's own odeint calls."""
    g = load(golden_dir, f"step_{variant}.npz")
    n = len(g["reward"])
    exact = 0
    hu_close = 0
    worst = 0.0
    for i in range(n):
        F, M = hh.mix(g["action"][i])
        assert abs(F - g["lsoda_F_clamped"][i]) <= 1e-15 * max(1.0, abs(F))  # BLAS gemv (FMA) vs plain products: <= 2 ulp
        np.testing.assert_allclose(M, g["lsoda_M_clamped"][i], rtol=1e-13, atol=1e-17)  # yaw moment cancels to ~1e-4 of its terms
        y, st = hh.lsoda(g["pre_y"][i], F, M)
        assert st["status"] == 0
        same = (st["nst"], st["nfe"], st["nqu"]) == (g["lsoda_nst"][i], g["lsoda_nfe"][i], g["lsoda_nqu"][i])
        exact += same
        if same:
            hu_close += abs(st["hu"] - g["lsoda_hu"][i]) < 1e-4 * g["lsoda_hu"][i]
        err = np.abs(y - g["lsoda_y_raw"][i]) / np.maximum(np.abs(g["lsoda_y_raw"][i]), 1.0)
        worst = max(worst, err.max())
    assert exact >= 0.98 * n, f"only {exact}/{n} calls reproduce scipy's step sequence"
    # measured here: 491/491 calls reproduce scipy's (nst, nfe, nqu); last step size within 7e-5 relative;
    # worst state error 4.8e-12.  (With the closed-form rotation instead of the reference's arccos route the
    # near-hover calls pick visibly different steps and the error grows to 1.6e-9 -- see qs_model.cuh.)
    assert hu_close >= 0.99 * n
    assert worst < 1e-10, worst


@pytest.mark.parametrize("variant", ["v2", "v1", "v1_raw"])
def test_f64_lsoda_step_vs_reference_golden(hh, golden_dir, variant):
    g = load(golden_dir, f"step_{variant}.npz")
    ver = VERS[variant]
    checked = 0
    for i in range(len(g["reward"])):
        v = 3 if ver == 2 and g["pre_n_wp"][i] > 1 else ver   # multi-waypoint v2 lists: the ENV_V2M variant (wider record)
        st, obs, rew, flags, ep_len, ls = hh.step(v, golden_state(g, "pre_", i), g["action"][i], f32=False, integ="lsoda",
                                                 obs_scaled=(variant != "v1_raw"))
        case = g["case"][i]
        np.testing.assert_allclose(st["y"], g["post_y"][i], rtol=1e-9, atol=1e-10, err_msg=f"{i} {case}")
        assert st["wp_index"] == g["post_wp_index"][i] and st["current_step"] == g["post_current_step"][i], (i, case)
        if ver == 2:
            assert st["counter"] == g["post_counter"][i] and st["final_reached"] == g["post_final_reached"][i], (i, case)
        if case != "on_waypoint_nan":
            assert flags == flags_from_golden(g, i), (i, case, flags)
            np.testing.assert_allclose(rew, g["reward"][i], rtol=1e-9, atol=1e-9, err_msg=f"{i} {case}")
        np.testing.assert_allclose(obs, g["obs"][i], rtol=3e-7, atol=1e-9, err_msg=f"{i} {case}")
        np.testing.assert_allclose(st["last_distance"], g["post_last_distance"][i], rtol=1e-9, atol=1e-10)
        checked += 1
    assert checked > 200


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_step_logic_exact_on_reference_state(hh, golden_dir, variant):
    """Feed the golden post-update dynamics state through a zero-length physics step: logic must be exact."""
    # RK4 with dt -> the harness always integrates, so this test drives the logic through the oracle instead:
    g = load(golden_dir, f"step_{variant}.npz")
    ver = VERS[variant]
    b = qo.EnvBatch.empty(variant, len(g["reward"]), max_wp=3)
    for f in ("y", "wp_list", "n_wp", "wp_index", "cur_wp", "last_distance", "current_step", "counter", "final_reached", "final_yaw"):
        setattr(b, f, np.array(g["pre_" + f]))
    with np.errstate(all="ignore"):
        obs_o, rew_o, term_o, trunc_o, info_o = qo.step(b, g["action"], integrator="rk4", substeps=4)
    for i in range(len(g["reward"])):
        v = 3 if ver == 2 and g["pre_n_wp"][i] > 1 else ver
        st, obs, rew, flags, ep_len, _ = hh.step(v, golden_state(g, "pre_", i), g["action"][i], f32=False, integ="rk4", substeps=4)
        np.testing.assert_allclose(st["y"], b.y[i], rtol=1e-12, atol=1e-13)
        if g["case"][i] == "on_waypoint_nan":
            continue
        want = int(term_o[i]) | int(trunc_o[i]) << 1 | (int(info_o[i]) & 0xF) << 2
        assert flags == want, (i, g["case"][i])
        np.testing.assert_allclose(rew, rew_o[i], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(obs, obs_o[i], rtol=2e-7, atol=1e-12)


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_f32_rk4_step_within_1e4(hh, golden_dir, variant):
    """float32 throughput mode vs the float64 reference: 1e-4 relative (north_star), flags equal away from thresholds."""
    g = load(golden_dir, f"step_{variant}.npz")
    ver = VERS[variant]
    n_flag = 0
    for i in range(len(g["reward"])):
        if ver == 2 and g["pre_n_wp"][i] > 1:
            continue
        case = g["case"][i]
        if case == "on_waypoint_nan":
            continue
        st, obs, rew, flags, ep_len, _ = hh.step(ver, golden_state(g, "pre_", i), g["action"][i], f32=True, integ="rk4", substeps=1)
        np.testing.assert_allclose(st["y"], g["post_y"][i], rtol=1e-4, atol=1e-5, err_msg=f"{i} {case}")
        np.testing.assert_allclose(obs, g["obs"][i], rtol=1e-4, atol=1e-5, err_msg=f"{i} {case}")
        assert flags == flags_from_golden(g, i), (i, case)
        n_flag += 1
        # the +2 progress bonus flips when |delta distance| is below float32 resolution; skip those few
        ld, d = g["pre_last_distance"][i], g["post_last_distance"][i]
        if not np.isnan(ld) and abs(ld - d) < 2e-6:
            continue
        np.testing.assert_allclose(rew, g["reward"][i], rtol=1e-4, atol=2e-4, err_msg=f"{i} {case}")
    assert n_flag > 200


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_reset_matches_oracle_on_same_uniforms(hh, variant):
    ver = VERS[variant]
    n = 300
    b = qo.EnvBatch.empty(variant, n, max_wp=3)
    U = np.array([hh.uniforms(1234, 10_000_000_000 + e, e % 5) for e in range(n)])
    qo.reset_from_uniforms(b, np.arange(n), U)
    obs_o = qo.observe(b)
    kinds = set()
    for e in range(n):
        st, obs = hh.reset(ver, 10_000_000_000 + e, e % 5)
        np.testing.assert_array_equal(st["y"], b.y[e])
        assert st["n_wp"] == b.n_wp[e] and st["wp_index"] == 0 and st["current_step"] == 0 and not st["has_last"]
        k = 1 if ver == 2 else 2
        np.testing.assert_array_equal(st["wp_list"][:k], b.wp_list[e][:k])
        if ver == 2:
            assert st["final_yaw"] == b.final_yaw[e]
        np.testing.assert_array_equal(obs, obs_o[e])
        kinds.add(int(b.n_wp[e]) if ver == 1 else (0 if U[e, 8] < 0.3 else 1 if U[e, 9] < 0.6 else 2))
    assert len(kinds) == (3 if ver == 2 else 2)


def test_reset_v2_multi_waypoint_matches_oracle_and_reference(hh, golden_dir):
    """ENV_V2M reset (num_waypoints = randint(2, 4), rl_env_scaledObs.py:46): the oracle reproduces the reference's reset with that
    line enabled bit for bit on the golden uniform blocks, and the device code reproduces the oracle on its own Philox uniforms."""
    g = load(golden_dir, "reset_v2m.npz")
    n = g["uniforms"].shape[0]
    b = qo.EnvBatch.empty("v2", n, max_wp=3)
    b.v2_random_waypoints = True
    qo.reset_from_uniforms(b, np.arange(n), g["uniforms"])
    for f in ("y", "wp_list", "n_wp", "final_yaw"):
        np.testing.assert_array_equal(getattr(b, f), g[f], err_msg=f)
    np.testing.assert_array_equal(qo.observe(b), g["obs"])
    assert set(np.unique(g["n_wp"])) == {2, 3}
    m = 400
    b = qo.EnvBatch.empty("v2", m, max_wp=3)
    b.v2_random_waypoints = True
    U = np.array([hh.uniforms(1234, 5_000_000_000 + e, e % 3) for e in range(m)])   # 1234: the seed the harness is configured with
    qo.reset_from_uniforms(b, np.arange(m), U)
    obs_o = qo.observe(b)
    seen = set()
    for e in range(m):
        st, obs = hh.reset(3, 5_000_000_000 + e, e % 3)
        assert st["n_wp"] == b.n_wp[e] and st["n_wp"] in (2, 3)
        np.testing.assert_array_equal(st["y"], b.y[e])
        np.testing.assert_array_equal(st["wp_list"][:st["n_wp"]], b.wp_list[e][:st["n_wp"]])
        assert st["final_yaw"] == b.final_yaw[e]
        np.testing.assert_array_equal(obs, obs_o[e])
        seen.add((st["n_wp"], 0 if U[e, 9] < 0.3 else 1 if U[e, 10] < 0.6 else 2))
    assert len(seen) == 6


def test_reset_known_answers_from_reference(hh, golden_dir):
    """Golden reset vectors were produced by the reference on scripted uniforms; replay them through the
    oracle (already pinned) is covered elsewhere -- here: the kernel's draw order equals the oracle's for the
    reference's uniform blocks by checking trajectory-kind statistics of the device generator."""
    n = 6000
    kinds = np.zeros(3)
    for e in range(n):
        u = hh.uniforms(99, e, 0)
        kinds[0 if u[8] < 0.3 else 1 if u[9] < 0.6 else 2] += 1
    np.testing.assert_allclose(kinds / n, [0.30, 0.42, 0.28], atol=0.02)  # mixture weights of rl_env_scaledObs.py:63-68


def test_euler_matches_oracle(hh):
    rng = np.random.default_rng(3)
    q = rng.normal(size=(500, 4))
    q[0] = [np.sqrt(0.5), 0, np.sqrt(0.5), 0]
    q[1] = [np.sqrt(0.5), 0, -np.sqrt(0.5), 0]
    out = np.zeros((500, 3))
    import ctypes as C
    for i in range(500):
        qi = np.ascontiguousarray(q[i])
        hh.lib.hh_rpy(qi.ctypes.data_as(C.c_void_p), out[i].ctypes.data_as(C.c_void_p))
    r, p, y = qo.quat_to_rpy(q[:, 0], q[:, 1], q[:, 2], q[:, 3])
    np.testing.assert_allclose(out, np.stack([r, p, y], 1), rtol=0, atol=1e-14)


@pytest.mark.parametrize("variant", ["v2", "v1", "v2m"])
def test_randomised_states_step_matches_oracle(hh, variant):
    """Differential test beyond the golden cases: 1,500 random internal states per variant -- waypoint lists of every length,
    counters and step counts at and around their limits, positions on both sides of the reach / crash / bounds thresholds,
    first-step states -- through the device code (float64, RK4 x 2) and through the oracle: flags, counters and waypoint
    bookkeeping exact, reward and state to 1e-11."""
    rng = np.random.default_rng({"v2": 1, "v1": 2, "v2m": 3}[variant])
    oname = "v1" if variant == "v1" else "v2"
    ver = {"v2": 2, "v1": 1, "v2m": 3}[variant]
    n = 1500
    kmax = {"v2": 1, "v1": 2, "v2m": 3}[variant]
    b = qo.EnvBatch.empty(oname, n, max_wp=3)
    b.n_wp[:] = rng.integers(1, kmax + 1, n)
    b.wp_list[:] = np.stack([rng.uniform(-1, 1, (n, 3)), rng.uniform(-1, 1, (n, 3)), rng.uniform(0.3, 3, (n, 3))], -1)
    b.wp_index[:] = np.minimum(rng.integers(0, 3, n), b.n_wp - (rng.random(n) < 0.5))      # some already past the last entry
    b.wp_index[:] = np.clip(b.wp_index, 0, b.n_wp)
    ar = np.arange(n)
    b.cur_wp = b.wp_list[ar, np.minimum(b.wp_index, b.n_wp - 1)].copy()
    b.y[:] = 0
    # position: a third right around the 0.1 m reach sphere, a few near the ground / far away, the rest anywhere
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.where(rng.random(n) < 0.35, rng.uniform(0.09, 0.11, n), rng.uniform(0.0, 2.5, n))
    b.y[:, 0:3] = b.cur_wp + d * r[:, None]
    low = rng.random(n) < 0.08
    b.y[low, 2] = rng.uniform(0.09, 0.11, low.sum())
    far = rng.random(n) < 0.04
    b.y[far, 0] = rng.choice([-1, 1], far.sum()) * rng.uniform(9.9, 10.1, far.sum())
    b.y[:, 3:6] = rng.normal(size=(n, 3)) * np.where(rng.random(n) < 0.3, 0.05, 1.0)[:, None]
    q = rng.normal(size=(n, 4)) * 0.3 + [1, 0, 0, 0]
    b.y[:, 6:10] = q / np.linalg.norm(q, axis=1, keepdims=True)
    b.y[:, 10:13] = rng.normal(size=(n, 3)) * np.where(rng.random(n) < 0.3, 0.05, 1.5)[:, None]
    b.last_distance[:] = np.where(rng.random(n) < 0.1, np.nan, np.linalg.norm(b.y[:, 0:3] - b.cur_wp, axis=1) + rng.normal(size=n) * 0.01)
    limit = qo.MAX_STEPS[oname]
    b.current_step[:] = np.where(rng.random(n) < 0.3, limit + rng.integers(-2, 2, n), rng.integers(0, limit, n))
    if oname == "v2":
        b.final_reached[:] = (b.wp_index >= b.n_wp) & (rng.random(n) < 0.9)
        b.wp_index[:] = np.where(b.final_reached, b.n_wp, np.minimum(b.wp_index, b.n_wp - 1))
        b.cur_wp = b.wp_list[ar, np.minimum(b.wp_index, b.n_wp - 1)].copy()
        b.counter[:] = np.where(b.final_reached, np.where(rng.random(n) < 0.5, 500 + rng.integers(-2, 3, n), rng.integers(0, 500, n)), 0)
        b.final_yaw[:] = rng.uniform(-np.pi, np.pi, n)
    else:
        b.wp_index[:] = np.minimum(b.wp_index, b.n_wp - 1)
        b.cur_wp = b.wp_list[ar, b.wp_index].copy()
    act = np.stack([rng.uniform(0, 2, n), *rng.uniform(-1, 1, (3, n))], 1).astype(np.float32)
    pre = b.copy()
    with np.errstate(all="ignore"):
        obs_o, rew_o, term_o, trunc_o, info_o = qo.step(b, act, integrator="rk4", substeps=2)
    seen = set()
    for i in range(n):
        st = dict(y=pre.y[i], wp_list=pre.wp_list[i], n_wp=int(pre.n_wp[i]), wp_index=int(pre.wp_index[i]),
                  last_distance=float(pre.last_distance[i]), current_step=int(pre.current_step[i]), counter=int(pre.counter[i]),
                  final_reached=bool(pre.final_reached[i]), final_yaw=float(pre.final_yaw[i]))
        out, obs, rew, flags, ep_len, _ = hh.step(ver, st, act[i], f32=False, integ="rk4", substeps=2)
        want = int(term_o[i]) | int(trunc_o[i]) << 1 | (int(info_o[i]) & 0xF) << 2
        # a state exactly on a threshold may round to either side in the two implementations' last bit: none is expected, but
        # say so if it happens instead of failing on a measure-zero tie
        dist = np.linalg.norm(b.y[i, 0:3] - pre.cur_wp[i])
        tie = min(abs(dist - 0.1), abs(dist - 0.5), abs(b.y[i, 2] - 0.1), abs(np.linalg.norm(b.y[i, 0:3]) - 10)) < 1e-12
        assert flags == want or tie, (i, flags, want)
        np.testing.assert_allclose(out["y"], b.y[i], rtol=1e-12, atol=1e-13)
        assert out["wp_index"] == min(int(b.wp_index[i]), 3) and out["current_step"] == b.current_step[i]
        if oname == "v2":
            assert out["counter"] == b.counter[i] and out["final_reached"] == bool(b.final_reached[i])
        np.testing.assert_allclose(rew, rew_o[i], rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(obs, obs_o[i], rtol=2e-7, atol=1e-12)
        seen.add(want)
    assert len(seen) >= (6 if oname == "v2" else 5), seen          # plain, truncated, success(+stopped), crashed, out of bounds ...


@pytest.mark.parametrize("variant", ["v2", "v1", "v2m"])
def test_reset_distributions_and_bounds(hh, variant):
    """4,000 resets of the device reset code per variant: value ranges and frequencies the reference's reset implies
    (rl_env_scaledObs.py:40-79 / v1:32-63, utils2/utils.py:12-94)."""
    ver = {"v2": 2, "v1": 1, "v2m": 3}[variant]
    n = 4000
    nw = np.zeros(n, dtype=int)
    start = np.zeros((n, 3))
    wps, yaws = [], []
    for e in range(n):
        st, obs = hh.reset(ver, 777_000 + e, e % 4)
        nw[e] = st["n_wp"]
        start[e] = st["y"][0:3]
        wps.append(st["wp_list"][: st["n_wp"]])
        yaws.append(st["final_yaw"])
        assert st["y"][6] == 1.0 and not np.any(st["y"][3:6]) and not np.any(st["y"][7:13])      # at rest, attitude (0,0,0)
        assert st["wp_index"] == 0 and st["current_step"] == 0 and st["counter"] == 0 and not st["final_reached"]
    assert np.all((start[:, :2] >= -1) & (start[:, :2] < 1)) and np.all((start[:, 2] >= 1) & (start[:, 2] < 2))
    assert abs(start[:, 0].mean()) < 0.05 and abs(start[:, 2].mean() - 1.5) < 0.03
    allw = np.concatenate(wps)
    if variant == "v1":
        assert set(np.unique(nw)) == {1, 2} and abs((nw == 1).mean() - 0.5) < 0.04                 # randint(1, 3)
        assert np.all((allw[:, :2] >= -1) & (allw[:, :2] < 1)) and np.all((allw[:, 2] >= 1) & (allw[:, 2] < 3))
    else:
        if variant == "v2":
            assert np.all(nw == 1)
        else:
            assert set(np.unique(nw)) == {2, 3} and abs((nw == 2).mean() - 0.5) < 0.04             # randint(2, 4)
        yaws = np.array(yaws)
        assert np.all((yaws >= -np.pi) & (yaws < np.pi)) and abs(yaws.mean()) < 0.15
        assert allw[:, 2].min() >= 0.2 - 1e-15                       # curved / helical clamp; linear ends at z >= 0.5
        assert np.all(np.abs(allw[:, :2]) < 2.9) and allw[:, 2].max() < 4.0 + 1e-9
        # the LAST waypoint of a linear / curved trajectory is the drawn end point (+ sin(2 pi) ~ -2.4e-16 on one axis), the helix ends
        # above its start: both keep the end inside the arena
        last = np.array([w[-1] for w in wps])
        assert np.all(last[:, 2] >= 0.5 - 1e-9)


# ------------------------------------------------------------------------------------------------------
# multi-step drift of the fixed-step throughput modes against the reference's trajectories (same device code, host build);
# the B200 runs the same checks through the C-ABI in tests/test_gpu_trajectory.py
# ------------------------------------------------------------------------------------------------------
def _state_of(b, i):
    return dict(y=b.y[i].copy(), wp_list=b.wp_list[i].copy(), n_wp=int(b.n_wp[i]), wp_index=int(b.wp_index[i]),
                last_distance=float(b.last_distance[i]), current_step=int(b.current_step[i]), counter=int(b.counter[i]),
                final_reached=bool(b.final_reached[i]), final_yaw=float(b.final_yaw[i]))


@pytest.mark.parametrize("variant,precision,substeps", [("v2", "f32", 1), ("v2", "f64", 1), ("v1", "f32", 1)])   # the full matrix runs on the GPU
def test_rk4_trajectory_drift_vs_reference_host(hh, golden_dir, variant, precision, substeps):
    from drift_bounds import check_step
    g = load(golden_dir, f"traj_{variant}.npz")
    n_steps, n_env = g["action"].shape[:2]
    for i in range(n_env):
        b = qo.EnvBatch.empty(variant, n_env, max_wp=3)
        qo.reset_from_uniforms(b, np.arange(n_env), g["uniforms"][:, 0])
        st, episode, h, last_d = _state_of(b, i), 0, 0, np.nan
        for t in range(n_steps):
            cw = st["wp_list"][min(st["wp_index"], st["n_wp"] - 1)]
            st, obs, rew, flags, ep_len, _ = hh.step(VERS[variant], st, g["action"][t, i], f32=(precision == "f32"), integ="rk4", substeps=substeps)
            h += 1
            assert flags == (int(g["terminated"][t, i]) | int(g["truncated"][t, i]) << 1 | (int(g["info"][t, i]) & 0xF) << 2), (i, t)
            d_ref = np.linalg.norm(g["y"][t, i, :3] - cw)
            dd = 1.0 if np.isnan(last_d) else last_d - d_ref
            last_d = d_ref
            check_step("traj", precision, np.array([float(h)]), st["y"][None], g["y"][t, i][None], np.array([rew]), g["reward"][t, i][None],
                       np.array([dd]), obs[None], g["terminal_obs"][t, i][None], tag=f"env {i} t={t}")
            if flags & 3:
                episode += 1
                bb = qo.EnvBatch.empty(variant, n_env, max_wp=3)
                qo.reset_from_uniforms(bb, np.arange(n_env), g["uniforms"][:, episode])
                st, h, last_d = _state_of(bb, i), 0, np.nan


@pytest.mark.parametrize("precision,substeps", [("f32", 1)])
def test_policy_action_replay_drift_vs_reference_host(hh, golden_dir, precision, substeps):
    from drift_bounds import check_step
    g = load(golden_dir, "closed_loop_v2.npz")
    n = 2
    b = qo.EnvBatch.empty("v2", n, max_wp=3)
    qo.reset_from_uniforms(b, np.arange(n), g["uniforms"][:n])
    for i in range(n):
        st, last_d = _state_of(b, i), np.nan
        L = int(g["length"][i])
        for t in range(L):
            st, obs, rew, flags, ep_len, _ = hh.step(2, st, g["traj_action"][i, t], f32=(precision == "f32"), integ="rk4", substeps=substeps)
            d_ref = np.linalg.norm(g["traj_y"][i, t, :3] - g["waypoint"][i])
            dd = 1.0 if np.isnan(last_d) else last_d - d_ref
            last_d = d_ref
            check_step("policy", precision, np.array([t + 1.0]), st["y"][None], g["traj_y"][i, t][None], np.array([rew]),
                       g["traj_reward"][i, t][None], np.array([dd]), tag=f"ep {i} t={t}")
            assert bool(flags & 3) == (t == L - 1), (i, t, flags)
        assert flags == (int(g["terminated"][i]) | int(g["truncated"][i]) << 1 | (int(g["info"][i]) & 0xF) << 2)
