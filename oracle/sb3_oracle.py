"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the Stable-Baselines3 pieces on the hot path.

stable-baselines3 is a third-party dependency of the reference (requirements.txt:4, unpinned; the shipped
zips record 2.6.0).  It is NOT installed in this image and cannot be (no network), so these functions restate
the published SB3 2.6.0 algorithms:
    common/running_mean_std.py      RunningMeanStd.update / update_from_moments
    common/vec_env/vec_normalize.py VecNormalize.step_wait / normalize_obs / reset
    common/policies.py              ActorCriticPolicy.forward for MlpPolicy(net_arch=[128,64,64], Tanh)
    common/distributions.py         DiagGaussianDistribution.sample / log_prob
Parity status: PARTIALLY PINNED -- what the reference tree can pin is pinned: the parameter shapes / key names by
the state_dict of the shipped zips (tests/golden/policy_*.npz, forward outputs generated with torch), and the
VecNormalize field layout, dtypes and hyper-parameters by initial-implementation-v1/vec_normalize.pkl
(tests/golden/vecnorm_v1.npz).  The update rule itself has no golden vector in the reference ("parity unpinned"
for RunningMeanStd.update: restated from the published source).
"""
from __future__ import annotations

import numpy as np


class RunningMeanStd:
    """SB3 RunningMeanStd.  `dtype_batch=np.float32` reproduces SB3's arithmetic on float32 observations
    (NumPy computes batch mean/var in float32); np.float64 is the exact-arithmetic variant."""

    def __init__(self, shape=(), epsilon=1e-4, dtype_batch=np.float64):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon
        self.dtype_batch = dtype_batch

    def update(self, arr):
        arr = np.asarray(arr).astype(self.dtype_batch, copy=False)
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, m_2 / (self.count + batch_count), batch_count + self.count


class VecNormalizeOracle:
    """VecNormalize(norm_obs=True, norm_reward=False) bookkeeping around externally supplied (obs, reward, done)."""

    def __init__(self, n_envs, obs_dim, clip_obs=10.0, gamma=0.99, epsilon=1e-8, dtype_batch=np.float64):
        self.obs_rms = RunningMeanStd((obs_dim,), dtype_batch=dtype_batch)
        self.ret_rms = RunningMeanStd((), dtype_batch=dtype_batch)
        self.clip_obs, self.gamma, self.epsilon = clip_obs, gamma, epsilon
        self.returns = np.zeros(n_envs)

    def normalize_obs(self, obs):
        return np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs, self.clip_obs).astype(np.float32)

    def reset(self, obs):
        self.returns = np.zeros_like(self.returns)
        self.obs_rms.update(obs)
        return self.normalize_obs(obs)

    def step(self, obs, rewards, dones):
        self.obs_rms.update(obs)
        out = self.normalize_obs(obs)
        self.returns = self.returns * self.gamma + rewards
        self.ret_rms.update(self.returns)
        self.returns[dones] = 0
        return out


def mlp_policy_forward(sd, obs, noise=None, dtype=np.float64):
    """ActorCriticPolicy.forward: returns (actions, values, log_prob, mean)."""
    g = lambda k: np.asarray(sd[k], dtype=dtype)
    x = np.asarray(obs, dtype=dtype)

    def trunk(x, net):
        for i in (0, 2, 4):
            x = np.tanh(x @ g(f"mlp_extractor.{net}.{i}.weight").T + g(f"mlp_extractor.{net}.{i}.bias"))
        return x
    mean = trunk(x, "policy_net") @ g("action_net.weight").T + g("action_net.bias")
    value = (trunk(x, "value_net") @ g("value_net.weight").T + g("value_net.bias"))[:, 0]
    log_std = g("log_std")
    eps = np.zeros_like(mean) if noise is None else np.asarray(noise, dtype=dtype)
    actions = mean + np.exp(log_std) * eps
    logp = np.sum(-0.5 * eps ** 2 - log_std - 0.5 * np.log(2 * np.pi), axis=1)
    return actions, value, logp, mean


def merge_moments(stats, moments):
    """CPU restatement of qs_vecnorm_merge / qs_xchg_merge (RunningMeanStd.update_from_moments applied to k batch triplets in rank
    order) for float64 torch CPU tensors: stats (count, mean[d], var[d]) <- merge of triplets (n, mean[d], M2[d]).
    Test infrastructure: the world-size-2 gloo test of the moment exchange uses it as the checker."""
    import torch
    d = (stats.shape[0] - 1) // 2
    count, mean, var = stats[0].clone(), stats[1:1 + d].clone(), stats[1 + d:].clone()
    for m in moments.reshape(-1, 1 + 2 * d):
        bn = m[0]
        if bn <= 0:
            continue
        delta = m[1:1 + d] - mean
        tot = count + bn
        mean = mean + delta * bn / tot
        M2 = var * count + m[1 + d:] + delta * delta * count * bn / tot
        var = M2 / tot
        count = tot
    return torch.cat([count.reshape(1), mean, var])


def torch_policy_forward(state_dict, obs, dtype=None):
    """Plain torch forward of SB3's MlpPolicy (mlp_extractor policy_net / value_net: Linear-Tanh x 3, then action_net / value_net)
    on `obs`'s device -- the float32 (or float64) checker of the CUDA policy kernels.  `state_dict`: name -> numpy array, the
    names of the policy.pth inside an SB3 zip.  Returns (mean[n,4], value[n])."""
    import torch

    dtype = dtype or torch.float32
    sd = {k: torch.from_numpy(np.asarray(v)).to(obs.device, dtype) for k, v in state_dict.items()}
    x = obs.to(dtype)

    def mlp(x, net):
        for i in (0, 2, 4):
            x = torch.tanh(x @ sd[f"mlp_extractor.{net}.{i}.weight"].T + sd[f"mlp_extractor.{net}.{i}.bias"])
        return x
    mean = mlp(x, "policy_net") @ sd["action_net.weight"].T + sd["action_net.bias"]
    value = (mlp(x, "value_net") @ sd["value_net.weight"].T + sd["value_net.bias"])[:, 0]
    return mean, value
