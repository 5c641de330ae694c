"""qs_step_many (T env steps per launch, hidden state in registers; csrc/qs_step_many.cu) against T calls of qs_step -- the
kernel that is itself pinned to the oracle and the reference goldens (tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(n, **kw):
    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
    env = BatchedQuadEnv(n, **kw)
    env.reset()
    y = env.get_state(["y"])["y"]
    y[: n // 3, 2] = 0.13                          # a third of the envs start just above the crash height: terminations + auto-resets
    y[: n // 3, 5] = -0.5
    env.set_state(y=y)
    return env


def random_actions(T, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lo = torch.tensor([0.0, -1, -1, -1], device="cuda")
    return (lo + 2.0 * torch.rand((T, n, 4), device="cuda", generator=g)).contiguous()


@pytest.mark.parametrize("n,ver,precision", [(4099, 2, "f32"), (1000, 1, "f32"), (2052, 2, "f64"), (96, 2, "f32")])
def test_step_many_equals_repeated_single_steps(n, ver, precision):
    """T = 12 steps in one launch == 12 launches of the same kernel with T = 1 (bit for bit: same code), and == 12 calls of qs_step
    (same device functions in another kernel: the compiler may contract a*b+c differently, so float rounding; flags equal except
    for an env straddling a threshold by an ulp)."""
    T = 12
    kw = dict(env_version=ver, precision=precision, seed=3)
    env_m, env_1, env_s = make(n, **kw), make(n, **kw), make(n, **kw)
    acts = random_actions(T, n, seed=n)
    out = env_m.step_many(T, actions=acts, store_terminal=True)
    obs_m, rew_m, fl_m = out["obs"].clone(), out["reward"].clone(), out["flags"].clone()
    tobs_m, er_m, el_m = out["terminal_obs"].clone(), out["ep_return"].clone(), out["ep_len"].clone()
    n_done = 0
    tol = dict(rtol=1e-5, atol=1e-5) if precision == "f32" else dict(rtol=1e-12, atol=1e-12)
    for t in range(T):
        o1 = env_1.step_many(1, actions=acts[t:t + 1].contiguous(), store_terminal=True)
        assert torch.equal(o1["obs"][0], obs_m[t]) and torch.equal(o1["reward"][0], rew_m[t]) and torch.equal(o1["flags"][0], fl_m[t]), f"t={t}"
        s = env_s.step(acts[t])
        same = s.flags == fl_m[t]
        assert int((~same).sum()) <= max(1, n // 2000), f"t={t}"
        if not bool(same.all()):                 # re-align the single-step twin (an ulp across a threshold diverges from here on)
            env_s.set_state(**env_1.get_state())
            continue
        torch.testing.assert_close(s.obs, obs_m[t], **tol)
        torch.testing.assert_close(s.reward, rew_m[t], rtol=1e-4 if precision == "f32" else 1e-11, atol=2e-3 if precision == "f32" else 1e-9)
        done = s.done
        n_done += int(done.sum())
        if bool(done.any()):
            torch.testing.assert_close(s.terminal_obs[done], tobs_m[t][done], **tol)
            assert torch.equal(s.ep_len[done], el_m[t][done])
            torch.testing.assert_close(s.ep_return[done], er_m[t][done], rtol=1e-4, atol=5e-2)
    assert n_done >= n // 10
    st_m, st_1 = env_m.get_state(), env_1.get_state()
    same_bits = lambda a, b: torch.equal(torch.nan_to_num(a.double(), nan=-7.0), torch.nan_to_num(b.double(), nan=-7.0))   # last_distance is NaN after a reset
    assert all(same_bits(st_m[k], st_1[k]) for k in st_m), [k for k in st_m if not same_bits(st_m[k], st_1[k])]    # the pool after T steps, bit for bit
    assert torch.equal(env_m.obs, obs_m[T - 1])
    for e in (env_m, env_1, env_s):
        e.close()


def test_step_many_in_kernel_actions():
    """actions=None: uniform over the action box from Philox keyed on (seed, GLOBAL env id, step); fresh every launch (device step
    counter, CUDA-graph replays included); a batch split into shards draws what the whole batch draws; the recorded actions
    replayed through qs_step give the same rollout."""
    n, T = 8192, 6
    kw = dict(env_version=2, precision="f32", seed=7)
    whole = make(n, **kw)
    a = whole.step_many(T, action_seed=11, store_actions=True)
    acts0, obs0, fl0 = a["actions"].clone(), a["obs"].clone(), a["flags"].clone()
    lo = torch.tensor([0.0, -1, -1, -1], device="cuda")
    hi = torch.tensor([2.0, 1, 1, 1], device="cuda")
    assert bool((acts0 >= lo).all()) and bool((acts0 < hi).all())
    u = ((acts0 - lo) / (hi - lo)).flatten().double()
    assert abs(float(u.mean()) - 0.5) < 0.005 and abs(float(u.var()) - 1 / 12) < 0.002
    assert float((acts0[0] - acts0[1]).abs().max()) > 0.1                         # a new draw every step
    acts1 = whole.step_many(T, action_seed=11, store_actions=True)["actions"].clone()
    assert not torch.equal(acts0, acts1)                                          # and every launch (step counter advanced by T)
    # shards: two handles with env_id_offset draw the halves of the whole batch's actions and produce its rollout
    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
    half = n // 2
    for off in (0, half):
        sh = BatchedQuadEnv(half, env_id_offset=off, **kw)
        sh.reset()
        y = sh.get_state(["y"])["y"]
        ref = make(n, **kw).get_state(["y"])["y"][off:off + half]
        sh.set_state(y=ref.clone())
        o = sh.step_many(T, action_seed=11, store_actions=True)
        assert torch.equal(o["actions"], acts0[:, off:off + half]) and torch.equal(o["obs"], obs0[:, off:off + half])
        assert torch.equal(o["flags"], fl0[:, off:off + half])
        sh.close()
    # graph replay: the counter lives on the device
    env = make(n, **kw)
    env.step_many(T, action_seed=5, store_actions=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.step_many(T, action_seed=5, store_actions=True)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        out = env.step_many(T, action_seed=5, store_actions=True)
    g.replay()
    r1 = out["actions"].clone()
    g.replay()
    r2 = out["actions"].clone()
    assert not torch.equal(r1, r2)
    # the recorded actions through the single-step kernel
    twin = make(n, **kw)
    for t in range(T):
        s = twin.step(acts0[t])
        same = s.flags == fl0[t]
        assert int((~same).sum()) <= 4
        if not bool(same.all()):
            break
        torch.testing.assert_close(s.obs, obs0[t], rtol=1e-5, atol=1e-5)
    for e in (whole, env, twin):
        e.close()


def test_step_many_argument_checks():
    from rl_aerial_manipulator_b200._cabi import QuadsimError
    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
    env = BatchedQuadEnv(64, env_version=2, precision="f64", integrator="lsoda")
    env.reset()
    with pytest.raises(QuadsimError, match="RK4"):
        env.step_many(2)
    env.close()
    env = BatchedQuadEnv(33, env_version=1, precision="f32")        # 33 * 17 is not a multiple of 4: time-major obs rows would be misaligned
    env.reset()
    with pytest.raises(QuadsimError, match="multiple of 4"):
        env.step_many(2)
    out = env.step_many(2, obs_last_only=True)
    assert out["obs"].shape == (33, 17)
    env.close()
