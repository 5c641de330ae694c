"""PPO pieces (rl-aerial-manipulator_b200/ppo.py, csrc/qs_gae.cu) against NumPy restatements of SB3 2.6.0."""
import math
import os

import numpy as np
import pytest
import torch


def sb3_gae(rewards, values, episode_starts, last_values, dones, gamma, lam):
    """RolloutBuffer.compute_returns_and_advantage, restated."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = 0.0
    for step in reversed(range(T)):
        if step == T - 1:
            nnt, nv = 1.0 - dones.astype(np.float32), last_values
        else:
            nnt, nv = 1.0 - episode_starts[step + 1], values[step + 1]
        delta = rewards[step] + gamma * nv * nnt - values[step]
        last = delta + gamma * lam * nnt * last
        adv[step] = last
    return adv, adv + values


def test_ppo_loss_matches_numpy_restatement():
    from rl_aerial_manipulator_b200.ppo import ppo_loss
    rng = np.random.default_rng(0)
    n = 512
    values, logp, old, adv, ret = (rng.normal(size=n).astype(np.float32) for _ in range(5))
    ent = np.full(n, 3.3, np.float32)
    t = lambda a: torch.from_numpy(a)
    loss, pg, vf, e = ppo_loss(t(values), t(logp), t(ent), t(old), t(adv), t(ret), 0.2, 0.01, 0.5, True)
    a = (adv - adv.mean()) / (adv.std(ddof=1) + 1e-8)                   # torch .std() is the unbiased one SB3 uses
    ratio = np.exp(logp - old)
    pg_np = -np.minimum(a * ratio, a * np.clip(ratio, 0.8, 1.2)).mean()
    vf_np = ((ret - values) ** 2).mean()
    want = pg_np + 0.01 * (-ent.mean()) + 0.5 * vf_np
    assert abs(float(loss) - want) < 1e-5 * max(1, abs(want)) and abs(float(pg) - pg_np) < 1e-5 and abs(float(vf) - vf_np) < 1e-4


def test_init_and_packing_roundtrip():
    from rl_aerial_manipulator_b200.policy import pack_params
    from rl_aerial_manipulator_b200.ppo import TorchActorCritic, init_state_dict
    sd = init_state_dict(20, seed=3)
    assert sd["action_net.weight"].shape == (4, 64) and np.allclose(sd["log_std"], 0)
    w = sd["mlp_extractor.policy_net.2.weight"]
    np.testing.assert_allclose(w @ w.T, 2 * np.eye(64), atol=1e-4)       # orthogonal rows, gain sqrt(2)
    net = TorchActorCritic(sd, 20)
    np.testing.assert_array_equal(net.packed().numpy(), pack_params(sd, 20))
    obs, act = torch.randn(16, 20), torch.randn(16, 4)
    v, lp, ent = net.evaluate_actions(obs, act)
    assert v.shape == (16,) and lp.shape == (16,) and abs(float(ent[0]) - 4 * (0.5 + 0.5 * math.log(2 * math.pi))) < 1e-6


@pytest.mark.gpu
def test_gae_kernel_vs_sb3_restatement():
    from rl_aerial_manipulator_b200.ppo import gae
    rng = np.random.default_rng(1)
    for T, n in ((1, 7), (16, 1000), (64, 4099)):
        rewards = rng.normal(size=(T, n)).astype(np.float32) * 10
        values = rng.normal(size=(T, n)).astype(np.float32) * 50
        starts = (rng.random((T, n)) < 0.1).astype(np.uint8)
        last_values = rng.normal(size=n).astype(np.float32) * 50
        dones = (rng.random(n) < 0.1).astype(np.uint8)
        c = lambda a: torch.from_numpy(a).cuda()
        adv, ret = gae(c(rewards), c(values), c(starts), c(last_values), c(dones), 0.995, 0.9)
        a_np, r_np = sb3_gae(rewards.astype(np.float64), values.astype(np.float64), starts.astype(np.float64), last_values.astype(np.float64),
                             dones, 0.995, 0.9)
        np.testing.assert_allclose(adv.cpu().numpy(), a_np, rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(ret.cpu().numpy(), r_np, rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
def test_quad_ppo_iterations_run_and_kernel_tracks_torch_weights():
    """Two collect/train iterations on 4096 envs with VecNormalize: finite losses, parameters move, the rollout kernel's forward
    equals the torch module's forward on the re-packed weights, rollout bookkeeping is consistent."""
    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
    from rl_aerial_manipulator_b200.ppo import QuadPPO
    from rl_aerial_manipulator_b200.vec_normalize import DeviceVecNormalize
    n = 4096
    env = BatchedQuadEnv(n, env_version=2, precision="f32", seed=1)
    vn = DeviceVecNormalize(env, norm_obs=True, norm_reward=False, gamma=0.995)
    ppo = QuadPPO(env, vecnorm=vn, n_steps=32, batch_size=16384, n_epochs=3, policy_impl="fp32", seed=0)
    before = ppo.net.packed().clone()
    logs = []
    ppo.learn(2 * 32 * n, log=logs.append)
    assert len(logs) == 2 and all(math.isfinite(l[k]) for l in logs for k in ("loss", "policy_gradient_loss", "value_loss", "entropy_loss"))
    assert ppo.num_timesteps == 2 * 32 * n
    after = ppo.net.packed()
    assert float((after - before).abs().max()) > 1e-5 and torch.equal(ppo.policy.params, after)
    assert abs(float(vn.obs_rms.count) - (1e-4 + n * (1 + 2 * 32))) < 1e-3          # reset + every step
    # GAE bookkeeping: returns = advantages + values; episode_starts marks the step after a done
    assert torch.allclose(ppo.returns, ppo.advantages + ppo.values, atol=1e-3)
    obs = ppo.obs[5]
    with torch.no_grad():
        v_t, lp_t, _ = ppo.net.evaluate_actions(obs, ppo.actions[5])
    a_k, v_k, _ = ppo.policy.forward(obs.contiguous())
    assert float((v_k - v_t).abs().max()) < 5e-3
    env.close()


@pytest.mark.gpu
def test_graph_replayed_update_equals_eager_update():
    """The CUDA-graph replay of the minibatch update (gather, forward, loss, backward, clipping, Adam) must reproduce the eager
    update: same seeds, same rollouts -> same parameters after two iterations (same kernels in the same order)."""
    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
    from rl_aerial_manipulator_b200.ppo import QuadPPO
    params = {}
    for graph in (False, True):
        env = BatchedQuadEnv(64, env_version=1, precision="f32", seed=3)
        ppo = QuadPPO(env, n_steps=64, batch_size=128, n_epochs=3, policy_impl="fp32", seed=5, graph_update=graph)
        logs = []
        ppo.learn(2 * 64 * 64, log=logs.append)
        assert len(logs) == 2 and all(math.isfinite(v) for l in logs for k, v in l.items() if k.endswith("loss"))
        params[graph] = ppo.net.packed().clone()
        env.close()
    assert float((params[True] - params[False]).abs().max()) < 1e-5


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one node (run with gpurun --gpus 2)")
def test_ppo_two_ranks_stay_in_lockstep():
    """BASELINE configs[4] mechanics on 2 ranks (tools/ppo_multi_rank_check.py under torchrun): sharded envs, VecNormalize statistics
    through the peer-memory exchange, NCCL gradient all-reduce -> bit-identical parameters and statistics on every rank."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29523", os.path.join(root, "tools", "ppo_multi_rank_check.py")], cwd=root, capture_output=True,
                       text=True, timeout=400)
    assert r.returncode == 0 and "PPO_MULTI_RANK_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
