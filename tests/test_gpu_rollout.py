"""GPU tests of the warp-specialised pipeline (csrc/qs_rollout.cu) through the C-ABI:
  * its policy part alone (`impl="tensor_pipeline"` of qs_policy_forward) against the float64 SB3 restatement, the torch fp32
    forward and the golden forwards of the shipped zips -- same stated tolerances as the chain kernels (POLICY_TOL["tensor"]);
  * the fused rollout step (qs_rollout_step) against the SAME work done by separate calls (policy forward + env step + moment
    kernels): identical flags, state/obs/reward to float32 rounding, VecNormalize statistics to 1e-12;
  * the in-kernel Philox noise against its NumPy restatement, its distribution, shard independence and graph-replay freshness.
"""
import os

import numpy as np
import pytest
import torch

from oracle import sb3_oracle as so
from test_gpu_policy_vecnorm import POLICY_TOL, t2n

pytestmark = pytest.mark.gpu


def make(n, golden_dir, seed=3, env_version=2, **kw):
    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
    from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
    env = BatchedQuadEnv(n, env_version=env_version, precision="f32", seed=seed, **kw)
    env.reset()
    pol = MlpPolicyKernel.from_npz(os.path.join(golden_dir, f"policy_v{env_version}.npz"), device="cuda", impl="tensor_pipeline")
    return env, pol


def status_ok():
    from rl_aerial_manipulator_b200 import load_library
    assert load_library().qs_rollout_status() == 0, "an in-kernel hand-over timed out"


@pytest.mark.parametrize("n", [1, 127, 129, 4099, 148 * 128 * 2 + 77, 262144])
def test_pipeline_policy_forward_vs_oracle(golden_dir, n):
    from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
    pol = MlpPolicyKernel.from_npz(os.path.join(golden_dir, "policy_v2.npz"), device="cuda", impl="tensor_pipeline")
    tol_a, tol_v = POLICY_TOL["tensor"]
    g = torch.Generator(device="cuda").manual_seed(n)
    obs = torch.randn((n, 20), device="cuda", generator=g) * 0.7
    noise = torch.randn((n, 4), device="cuda", generator=g)
    a, v, lp = pol.forward(obs, noise)
    status_ok()
    m = min(n, 4096)
    ao, vo, lpo, _ = so.mlp_policy_forward(pol.state_dict, t2n(obs[:m]), t2n(noise[:m]))
    np.testing.assert_allclose(t2n(a[:m]), ao, rtol=0, atol=tol_a)
    np.testing.assert_allclose(t2n(v[:m]), vo, rtol=0, atol=tol_v)
    np.testing.assert_allclose(t2n(lp[:m]), lpo, rtol=1e-5, atol=1e-5)
    mean_t, value_t = so.torch_policy_forward(pol.state_dict, obs)
    a_t = mean_t + torch.exp(torch.from_numpy(pol.state_dict["log_std"]).cuda()) * noise
    assert (a - a_t).abs().max() < tol_a and (v - value_t).abs().max() < tol_v
    lo, hi = torch.tensor([0.0, -1, -1, -1], device="cuda"), torch.tensor([2.0, 1, 1, 1], device="cuda")
    assert torch.equal(pol.actions_clipped, torch.minimum(torch.maximum(a, lo), hi))
    # deterministic: a second launch reproduces the first bit for bit
    a1, v1 = a.clone(), v.clone()
    a2, v2, _ = pol.forward(obs, noise)
    assert torch.equal(a1, a2) and torch.equal(v1, v2)


@pytest.mark.parametrize("tag,obs_dim", [("v2", 20), ("v1", 17)])
def test_pipeline_policy_forward_golden_and_vecnormalize(golden_dir, tag, obs_dim):
    from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
    from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd
    z = np.load(os.path.join(golden_dir, f"policy_{tag}.npz"))
    pol = MlpPolicyKernel.from_npz(os.path.join(golden_dir, f"policy_{tag}.npz"), device="cuda", impl="tensor_pipeline")
    tol_a, tol_v = POLICY_TOL["tensor"]
    a, v, _ = pol.forward(torch.from_numpy(z["obs"]).cuda())
    np.testing.assert_allclose(t2n(a), z["mean_f64"], rtol=0, atol=tol_a)
    np.testing.assert_allclose(t2n(v), z["value_f64"], rtol=0, atol=tol_v)
    n = 10000
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn((n, obs_dim), device="cuda", generator=g) * 3 + 1
    obs[:, 6] = 1.0 - 1e-4 * torch.rand(n, device="cuda", generator=g)
    rms = DeviceRunningMeanStd(obs_dim, "cuda")
    rms.update(obs)
    normed = rms.normalize(obs)
    a1, v1, _ = pol.forward(normed)
    a1, v1 = a1.clone(), v1.clone()
    out = torch.empty_like(obs)
    a2, v2, _ = pol.forward(obs, norm_stats=rms.stats, obs_norm_out=out)
    status_ok()
    assert torch.allclose(out, normed, atol=2e-6)
    assert torch.allclose(a1, a2, atol=1e-4) and torch.allclose(v1, v2, atol=1e-2)


@pytest.mark.parametrize("n,env_version", [(4099, 2), (300, 1), (148 * 128 + 5, 2)])
def test_fused_rollout_step_equals_separate_calls(golden_dir, n, env_version):
    """qs_rollout_step == qs_policy_forward (same pipeline kernel, policy part) + qs_step + the VecNormalize moment kernels, on
    the same noise, for 40 steps with auto-reset: flags equal, state / obs / reward equal to float32 rounding (the two kernels
    inline the same device functions; the compiler may contract a*b+c differently, and the action means differ by summation
    order), running statistics to 1e-9."""
    from rl_aerial_manipulator_b200.rollout import FusedRollout
    from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd
    env_a, pol = make(n, golden_dir, seed=11, env_version=env_version)
    env_b, _ = make(n, golden_dir, seed=11, env_version=env_version)
    d = env_a.obs_dim
    y = env_a.get_state(["y"])["y"]
    y[: n // 3, 2] = 0.13                          # a third of the envs start just above the crash height: terminations + auto-resets
    y[: n // 3, 5] = -0.5
    env_a.set_state(y=y)
    env_b.set_state(y=y)
    rms_a, rms_b = DeviceRunningMeanStd(d, "cuda"), DeviceRunningMeanStd(d, "cuda")
    for rms, env in ((rms_a, env_a), (rms_b, env_b)):
        rms.update(env.obs)
        rms.attach(env, merge=True)
    fused = FusedRollout(env_a, pol, vecnorm=rms_a, sample="noise", store_obs_norm=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    obs_norm_b = torch.empty((n, d), device="cuda")
    n_done = 0
    for t in range(40):
        noise = torch.randn((n, 4), device="cuda", generator=g)
        # separate calls on env_b
        a_b, v_b, lp_b = pol.forward(env_b.obs, noise, norm_stats=rms_b.stats, obs_norm_out=obs_norm_b)
        a_b, v_b, lp_b, ac_b = a_b.clone(), v_b.clone(), lp_b.clone(), pol.actions_clipped.clone()
        out_b = env_b.step(ac_b)
        rms_b.update_from_moments()
        # fused on env_a
        out_a = fused.step(noise)
        assert fused.status() == 0
        # same arithmetic, but the two builds of the pipeline add the head's partial sums in a different order (one epilogue warp per
        # row in the fused configuration, two in the policy-only one): float32 summation order over 64 products of weights up to
        # O(10) -- each build is within POLICY_TOL["tensor"] (1e-4 / 1e-2) of the float64 forward, so they are that close to each
        # other (measured: 7 of 16,396 means differ by more than 2e-5, the largest by 4.9e-5)
        torch.testing.assert_close(fused.actions, a_b, rtol=0, atol=POLICY_TOL["tensor"][0])
        torch.testing.assert_close(fused.values, v_b, rtol=0, atol=POLICY_TOL["tensor"][1])
        torch.testing.assert_close(fused.actions_clipped, ac_b, rtol=0, atol=POLICY_TOL["tensor"][0])
        assert float((fused.actions - a_b).abs().mean()) < 2e-6 and float((fused.values - v_b).abs().mean()) < 5e-4
        # the log-prob depends on the noise alone.  The normalised observations: the two envs' observations agree to 1e-5 (below), and
        # the normalisation multiplies that by 1/std of the column -- up to ~10 for the narrow angular-rate columns (measured: 7e-6)
        assert torch.equal(fused.logp, lp_b), f"t={t}"
        torch.testing.assert_close(fused.obs_norm, obs_norm_b, rtol=1e-4, atol=1e-4)
        same = out_a.flags == out_b.flags
        assert (~same).sum() <= max(1, n // 2000), f"t={t}: {(~same).sum()} flag mismatches"

        def realign():
            # the action means of the two builds differ by float32 summation order (up to 5e-5): fed back through the dynamics for 40
            # steps that would grow, so env_a restarts every step from env_b's state and the comparison stays a per-step one
            env_a.set_state(**{k: v for k, v in env_b.get_state().items()})
            env_a.obs.copy_(env_b.obs)
            rms_a.stats.copy_(rms_b.stats)
            rms_a._moments.copy_(rms_b._moments)
        if not bool(same.all()):               # an env straddling a threshold by a float32 ulp: skip the value checks of this step
            realign()
            continue
        torch.testing.assert_close(out_a.obs, out_b.obs, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out_a.reward, out_b.reward, rtol=1e-4, atol=2e-3)
        done = out_b.done
        n_done += int(done.sum())
        if bool(done.any()):
            torch.testing.assert_close(out_a.terminal_obs[done], out_b.terminal_obs[done], rtol=1e-5, atol=1e-5)
            assert torch.equal(out_a.ep_len[done], out_b.ep_len[done])
            torch.testing.assert_close(out_a.ep_return[done], out_b.ep_return[done], rtol=1e-4, atol=5e-2)
        sa, sb = env_a.get_state(["y", "episode", "current_step"]), env_b.get_state(["y", "episode", "current_step"])
        torch.testing.assert_close(sa["y"], sb["y"], rtol=1e-5, atol=1e-5)
        assert torch.equal(sa["episode"], sb["episode"]) and torch.equal(sa["current_step"], sb["current_step"])
        np.testing.assert_allclose(t2n(rms_a.stats), t2n(rms_b.stats), rtol=1e-6, atol=1e-9)
        realign()
    assert n_done > n // 10
    env_a.close()
    env_b.close()


def philox_normal_oracle(seed, gids, step):
    """NumPy restatement of philox_normal4 (csrc/qs_rollout.cu): Philox4x32-10, counter (gid lo, gid hi, step lo, step hi), key
    seed ^ (0x85A308D3, 0x243F6A88); two Box-Muller pairs from (w0, w1), (w2, w3)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    gids = np.asarray(gids, dtype=np.uint64)
    c = [gids & 0xFFFFFFFF, gids >> np.uint64(32), np.full_like(gids, step & 0xFFFFFFFF), np.full_like(gids, step >> 32)]
    k0, k1 = (seed & 0xFFFFFFFF) ^ 0x85A308D3, (seed >> 32) ^ 0x243F6A88
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & 0xFFFFFFFF, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    z = np.zeros((len(gids), 4))
    for j in range(2):
        u1 = ((c[2 * j] >> np.uint64(9)).astype(np.float64) + 0.5) * 2.0 ** -23
        th = (c[2 * j + 1] >> np.uint64(8)).astype(np.float64) * (2 * np.pi * 2.0 ** -24)
        r = np.sqrt(-2.0 * np.log(u1))
        z[:, 2 * j], z[:, 2 * j + 1] = r * np.cos(th), r * np.sin(th)
    return z


def test_in_kernel_philox_noise(golden_dir):
    """sample="philox": eps = (action - mean) / std reproduces the NumPy restatement of the generator; it is standard normal,
    fresh every step (also when the step is replayed from a CUDA graph), and keyed on the GLOBAL env id (shard independent)."""
    from rl_aerial_manipulator_b200.rollout import FusedRollout
    n, seed = 8192, 0x1234567890ABCDEF
    env, pol = make(n, golden_dir, seed=2, env_id_offset=5_000_000_000)
    fused = FusedRollout(env, pol, sample="philox", noise_seed=seed)
    det = FusedRollout(*make(n, golden_dir, seed=2, env_id_offset=5_000_000_000), sample="mean")
    std = np.exp(pol.state_dict["log_std"].astype(np.float64))
    gids = np.arange(n, dtype=np.uint64) + np.uint64(5_000_000_000)
    eps_all = []
    for t in range(3):
        det.step()
        fused.step()
        assert fused.status() == 0
        if t == 0:      # same state on both: mean actions are comparable only on the first step
            eps = (t2n(fused.actions).astype(np.float64) - t2n(det.actions)) / std
            np.testing.assert_allclose(eps, philox_normal_oracle(seed, gids, t), rtol=0, atol=2e-5 * np.maximum(1, std.max()))
        lp = t2n(fused.logp).astype(np.float64)
        eps_sq = -2.0 * (lp + pol.state_dict["log_std"].sum() + 2 * np.log(2 * np.pi))
        np.testing.assert_allclose(eps_sq, (philox_normal_oracle(seed, gids, t) ** 2).sum(1), rtol=1e-4, atol=1e-4)
        eps_all.append(philox_normal_oracle(seed, gids, t))
    assert int(fused.noise_step.item()) == 3
    e = np.concatenate(eps_all).ravel()
    assert abs(e.mean()) < 0.02 and abs(e.var() - 1) < 0.02 and abs((e ** 4).mean() - 3) < 0.15
    assert not np.allclose(eps_all[0], eps_all[1])
    # graph replay: the counter lives on the device, so every replay draws new noise
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fused.step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fused.step()
    seen = []
    for _ in range(3):
        graph.replay()
        seen.append(fused.logp.clone())
    torch.cuda.synchronize()
    assert fused.status() == 0
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    assert int(fused.noise_step.item()) == 3 + 1 + 3          # warm-up + three replays (capture itself executes nothing)
    # shard independence: the second half of the batch as its own handle draws the same noise
    env.close()
    half = n // 2
    env_h, pol_h = make(half, golden_dir, seed=2, env_id_offset=5_000_000_000 + half)
    fh = FusedRollout(env_h, pol_h, sample="philox", noise_seed=seed)
    env_w, pol_w = make(n, golden_dir, seed=2, env_id_offset=5_000_000_000)
    fw = FusedRollout(env_w, pol_w, sample="philox", noise_seed=seed)
    for t in range(5):
        fh.step()
        fw.step()
    assert torch.equal(fh.actions, fw.actions[half:]) and torch.equal(env_h.obs, env_w.obs[half:])
    assert torch.equal(env_h.flags, env_w.flags[half:])


def test_fused_rollout_1M_envs_properties(golden_dir):
    """BASELINE config 4 size: 1,048,576 envs, 30 fused steps with in-kernel noise: finite outputs, auto-reset keeps every env
    inside its bounds, episode counters advance, VecNormalize statistics equal a float64 recomputation of the last batch."""
    from rl_aerial_manipulator_b200.rollout import FusedRollout
    from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd
    n = 1 << 20
    env, pol = make(n, golden_dir, seed=9)
    rms = DeviceRunningMeanStd(20, "cuda")
    rms.update(env.obs)
    rms.attach(env, merge=True)
    fused = FusedRollout(env, pol, vecnorm=rms, sample="philox", noise_seed=1)
    dones = torch.zeros(n, dtype=torch.int32, device="cuda")
    for t in range(30):
        out = fused.step()
        dones += out.done.int()
    assert fused.status() == 0
    assert torch.isfinite(env.obs).all() and torch.isfinite(env.reward).all() and torch.isfinite(fused.values).all()
    st = env.get_state(["y", "episode"])
    assert torch.equal(st["episode"], dones)
    assert bool((st["y"][:, 2] >= 0.0).all()) and bool((st["y"][:, :3].norm(dim=1) < 10.5).all())
    m = t2n(rms._moments)
    x = env.obs.double()
    assert m[0] == n
    np.testing.assert_allclose(m[1:21], t2n(x.mean(0)), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(m[21:], t2n(x.var(0, unbiased=False)) * n, rtol=1e-7, atol=1e-9)
    assert abs(float(rms.count) - (1e-4 + 31 * n)) < 1.0
    env.close()


def test_policy_forward_with_in_kernel_noise(golden_dir):
    """qs_policy_forward_philox (the policy kernel of the two-launch rollout): eps = (action - mean) / std reproduces the NumPy
    restatement of the generator for (seed, global env id, step); the step counter advances per launch; mean / value / log-prob
    equal those of qs_policy_forward fed with that noise."""
    from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
    n, seed, off = 20000, 0x0123456789ABCDEF, 1 << 33
    pol = MlpPolicyKernel.from_npz(os.path.join(golden_dir, "policy_v2.npz"), device="cuda", impl="tensor_pipeline")
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn((n, 20), device="cuda", generator=g)
    mean = pol.forward(obs, None)[0].clone()
    std = torch.from_numpy(np.exp(pol.state_dict["log_std"])).cuda()
    for step in range(3):
        a, v, lp = pol.forward_sampled(obs, noise_seed=seed, env_id_offset=off)
        a, v, lp, ac = a.clone(), v.clone(), lp.clone(), pol.actions_clipped.clone()
        eps = ((a - mean) / std).cpu().numpy()
        want = philox_normal_oracle(seed, off + np.arange(n), step)
        np.testing.assert_allclose(eps, want, rtol=0, atol=2e-4)           # device __sincosf / __log2f vs NumPy, and the division by std
        a2, v2, lp2 = pol.forward(obs, torch.from_numpy(want.astype(np.float32)).cuda())
        torch.testing.assert_close(a, a2, rtol=0, atol=3e-4)
        torch.testing.assert_close(v, v2, rtol=0, atol=1e-6)
        torch.testing.assert_close(lp, lp2, rtol=0, atol=5e-3)
        assert torch.equal(ac, torch.minimum(torch.maximum(a, torch.tensor([0.0, -1, -1, -1], device="cuda")), torch.tensor([2.0, 1, 1, 1], device="cuda")))
    assert int(pol._philox_counter[0]) == 3 and int(pol._philox_counter[1]) == 0
    assert pol.lib.qs_rollout_status() == 0
