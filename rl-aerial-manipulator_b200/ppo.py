"""QuadPPO -- on-device PPO training loop around the batched simulator.

Mirrors `stable_baselines3.PPO("MlpPolicy", env, ...).learn(total_timesteps)` as the reference calls it
(initial-implementation-v1/rl_train_vecN.py:13-36: n_steps=2048, batch_size=128, n_epochs=10, gamma=0.995, gae_lambda=0.9,
clip_range=0.2, ent_coef=0.01, learning_rate=2e-4, net_arch=[128,64,64], Tanh; initial-implementation-v2/rl_train.py:27-56)
with SB3 2.6.0 semantics restated: rollout collection with action clipping and time-limit bootstrapping
(OnPolicyAlgorithm.collect_rollouts), GAE (RolloutBuffer.compute_returns_and_advantage), the clipped surrogate / value /
entropy loss with advantage normalisation, gradient-norm clipping and Adam(eps=1e-5) (PPO.train).

What runs where:
  * rollout: hand-written kernels only -- env step (+ fused VecNormalize moments), tcgen05 policy forward, GAE (csrc/qs_gae.cu);
    everything stays in HBM, time-major buffers [T, N, ...];
  * update: the minibatch forward/backward of the 30k-parameter MLP uses torch autograd (library GEMMs -- the same role cuBLAS
    plays; no hand-written backward this round), gradients are all-reduced over NCCL when the env batch is sharded over ranks;
    the updated weights are re-packed on device into the blob the rollout kernels read.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from ._cabi import load_library
from .policy import MlpPolicyKernel, H1, H2, H3, NACT


def _bind(lib):
    if getattr(lib, "_gae_bound", False):
        return
    vp = C.c_void_p
    lib.qs_gae.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int64, C.c_float, C.c_float, vp, vp, vp]
    lib.qs_gae.restype = C.c_int
    lib.qs_gae_last_error.restype = C.c_char_p
    lib._gae_bound = True


def gae(rewards, values, episode_starts, last_values, last_dones, gamma: float, gae_lambda: float, out=None):
    """advantages, returns = GAE over time-major device buffers [T, N] (qs_gae); out = (advantages, returns) to reuse buffers."""
    lib = load_library()
    _bind(lib)
    T, n = rewards.shape
    adv, ret = out if out is not None else (torch.empty_like(rewards), torch.empty_like(rewards))
    p = lambda t: C.c_void_p(t.data_ptr())
    for t in (rewards, values, episode_starts, last_values, last_dones):
        assert t.is_contiguous() and t.is_cuda
    rc = lib.qs_gae(p(rewards), p(values), p(episode_starts), p(last_values), p(last_dones), T, n, gamma, gae_lambda, p(adv), p(ret),
                    C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"qs_gae failed ({rc}): {lib.qs_gae_last_error().decode()}")
    return adv, ret


def init_state_dict(obs_dim: int, seed: int = 0, log_std_init: float = 0.0) -> dict:
    """SB3 ActorCriticPolicy initialisation: orthogonal weights (gain sqrt(2) trunk, 0.01 action head, 1 value head), zero biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def ortho(shape, gain):
        w = torch.empty(shape)
        nn.init.orthogonal_(w, gain=gain, generator=g)
        return w.numpy()
    for net in ("policy_net", "value_net"):
        for i, (k_in, k_out) in zip((0, 2, 4), ((obs_dim, H1), (H1, H2), (H2, H3))):
            sd[f"mlp_extractor.{net}.{i}.weight"] = ortho((k_out, k_in), math.sqrt(2))
            sd[f"mlp_extractor.{net}.{i}.bias"] = np.zeros(k_out, np.float32)
    sd["action_net.weight"], sd["action_net.bias"] = ortho((NACT, H3), 0.01), np.zeros(NACT, np.float32)
    sd["value_net.weight"], sd["value_net.bias"] = ortho((1, H3), 1.0), np.zeros(1, np.float32)
    sd["log_std"] = np.full(NACT, log_std_init, np.float32)
    return sd


class TorchActorCritic(nn.Module):
    """The same network as the rollout kernels, as torch modules, for the update step (parameter names == SB3's)."""

    def __init__(self, sd: dict, obs_dim: int):
        super().__init__()
        mk = lambda: nn.Sequential(nn.Linear(obs_dim, H1), nn.Tanh(), nn.Linear(H1, H2), nn.Tanh(), nn.Linear(H2, H3), nn.Tanh())
        self.mlp_extractor = nn.ModuleDict({"policy_net": mk(), "value_net": mk()})
        self.action_net, self.value_net = nn.Linear(H3, NACT), nn.Linear(H3, 1)
        self.log_std = nn.Parameter(torch.zeros(NACT))
        self.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})

    def evaluate_actions(self, obs, actions):
        """ActorCriticPolicy.evaluate_actions: values, log_prob(actions), entropy."""
        mean = self.action_net(self.mlp_extractor["policy_net"](obs))
        values = self.value_net(self.mlp_extractor["value_net"](obs))[:, 0]
        logp = (-0.5 * ((actions - mean) / self.log_std.exp()) ** 2 - self.log_std - 0.5 * math.log(2 * math.pi)).sum(1)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum().expand_as(logp)
        return values, logp, entropy

    def packed(self) -> torch.Tensor:
        """Blob layout of include/quadsim.h, built on device (same order as policy.pack_params)."""
        parts = []
        for net, head in (("policy_net", self.action_net), ("value_net", self.value_net)):
            seq = self.mlp_extractor[net]
            for i in (0, 2, 4):
                parts += [seq[i].weight.t().reshape(-1), seq[i].bias]
            wpad = torch.zeros((H3, NACT), device=head.weight.device)
            bpad = torch.zeros(NACT, device=head.weight.device)
            wpad[:, : head.weight.shape[0]] = head.weight.t()
            bpad[: head.bias.shape[0]] = head.bias
            parts += [wpad.reshape(-1), bpad]
        parts.append(self.log_std)
        return torch.cat([p.detach().reshape(-1) for p in parts]).float()


def ppo_loss(values, logp, entropy, old_logp, advantages, returns, clip_range: float, ent_coef: float, vf_coef: float,
             normalize_advantage: bool = True):
    """SB3 PPO.train loss for one minibatch (clip_range_vf=None)."""
    if normalize_advantage and advantages.numel() > 1:
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    ratio = torch.exp(logp - old_logp)
    pg = -torch.min(advantages * ratio, advantages * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    vf = torch.nn.functional.mse_loss(returns, values)
    ent = -entropy.mean()
    return pg + ent_coef * ent + vf_coef * vf, pg, vf, ent


class QuadPPO:
    def __init__(self, env, vecnorm=None, state_dict: dict | None = None, n_steps: int = 64, batch_size: int = 65536, n_epochs: int = 10,
                 gamma: float = 0.995, gae_lambda: float = 0.9, clip_range: float = 0.2, ent_coef: float = 0.01, vf_coef: float = 0.5,
                 max_grad_norm: float = 0.5, learning_rate: float = 2e-4, normalize_advantage: bool = True, seed: int = 0,
                 policy_impl: str = "auto", graph_update: bool = True):
        """graph_update: replay every minibatch update (gather, forward, loss, backward, gradient clipping, Adam) from one CUDA graph
        -- at SB3-size minibatches (128) the ~60 launches of an eager update are pure launch latency.  Single rank only; with
        several ranks the update runs eagerly around the NCCL gradient all-reduce."""
        self.env, self.vecnorm = env, vecnorm
        self.n_steps, self.batch_size, self.n_epochs = n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda, self.clip_range = gamma, gae_lambda, clip_range
        self.ent_coef, self.vf_coef, self.max_grad_norm, self.normalize_advantage = ent_coef, vf_coef, max_grad_norm, normalize_advantage
        dev, n, d = env.device, env.n_envs, env.obs_dim
        sd = state_dict or init_state_dict(d, seed)
        self.net = TorchActorCritic(sd, d).to(dev)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.graph_update = bool(graph_update) and self.world == 1
        self.opt = torch.optim.Adam(self.net.parameters(), lr=learning_rate, eps=1e-5, capturable=self.graph_update)
        self._graph = None
        self.policy = MlpPolicyKernel(sd, d, dev, impl=policy_impl)
        self.gen = torch.Generator(device=dev).manual_seed(seed + 1000 * (dist.get_rank() if dist.is_initialized() else 0))
        T = n_steps
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = torch.empty((T, n, d), **f32)
        self.actions = torch.empty((T, n, NACT), **f32)
        self.values, self.logp, self.rewards = torch.empty((T, n), **f32), torch.empty((T, n), **f32), torch.empty((T, n), **f32)
        self.episode_starts = torch.zeros((T, n), dtype=torch.uint8, device=dev)
        self._noise = torch.empty((n, NACT), **f32)
        self._last_dones = torch.ones(n, dtype=torch.uint8, device=dev)
        self._last_obs = None
        self.num_timesteps = 0
        self._ep_stats = torch.zeros(2, dtype=torch.float64, device=dev)
        self._zero = torch.zeros((), dtype=torch.float64, device=dev)
        self.ep_rew_mean, self.ep_count = float("nan"), 0

    # ---- rollout -------------------------------------------------------------------------------------
    def _forward(self, obs_raw, noise, obs_norm_out=None):
        stats = self.vecnorm.obs_rms.stats if (self.vecnorm is not None and self.vecnorm.norm_obs) else None
        eps, clip = (self.vecnorm.epsilon, self.vecnorm.clip_obs) if self.vecnorm is not None else (1e-8, 10.0)
        return self.policy.forward(obs_raw, noise, norm_stats=stats, norm_eps=eps, norm_clip=clip, obs_norm_out=obs_norm_out)

    def collect_rollouts(self):
        env = self.env
        if self._last_obs is None:
            self._last_obs = env.reset()
            if self.vecnorm is not None and self.vecnorm.norm_obs and self.vecnorm.training:
                self.vecnorm.obs_rms.update(self._last_obs)      # VecNormalize.reset() updates the statistics too
                self.vecnorm.obs_rms.attach(env)                 # from here on the step kernel reduces its own observations
        self._ep_stats.zero_()
        for t in range(self.n_steps):
            self._noise.normal_(generator=self.gen)
            raw = self._last_obs
            norm_out = self.obs[t] if self.vecnorm is not None and self.vecnorm.norm_obs else None
            a, v, lp = self._forward(raw, self._noise, norm_out)
            if norm_out is None:
                self.obs[t].copy_(raw)
            self.actions[t].copy_(a)
            self.values[t].copy_(v)
            self.logp[t].copy_(lp)
            self.episode_starts[t].copy_(self._last_dones)
            out = env.step(self.policy.actions_clipped)            # SB3 clips to the action box before env.step, stores the unclipped
            self.rewards[t].copy_(out.reward)
            if self.vecnorm is not None and self.vecnorm.training:
                if self.vecnorm.norm_obs:
                    self.vecnorm.obs_rms.update_from_moments()
                self.vecnorm.update_returns(out)
            trunc_only = (out.flags & 3) == 2                        # TimeLimit.truncated: bootstrap with gamma * V(terminal_obs)
            if bool(trunc_only.any()):
                idx = trunc_only.nonzero(as_tuple=True)[0]
                tv = self._forward(out.terminal_obs.index_select(0, idx).contiguous(), None)[1].clone()
                self.rewards[t].index_add_(0, idx, self.gamma * tv)
            done = (out.flags & 3) != 0
            self._last_dones = done.to(torch.uint8)
            self._ep_stats[0] += torch.where(done, out.ep_return.double(), self._zero).sum()     # Monitor-style episode returns,
            self._ep_stats[1] += done.sum()                                                     # reduced on device: no per-step sync
            self._last_obs = out.obs
        last_values = self._forward(self._last_obs, None)[1].clone()
        if getattr(self, "advantages", None) is None:            # persistent: the captured update graph reads these addresses
            self.advantages, self.returns = torch.empty_like(self.rewards), torch.empty_like(self.rewards)
        gae(self.rewards, self.values, self.episode_starts, last_values, self._last_dones, self.gamma, self.gae_lambda,
            out=(self.advantages, self.returns))
        self.num_timesteps += self.n_steps * env.n_envs * self.world
        s, c = self._ep_stats.tolist()
        self.ep_rew_mean = s / c if c > 0 else float("nan")     # mean return of the episodes that finished in this rollout
        self.ep_count = int(c)

    # ---- update --------------------------------------------------------------------------------------
    def _minibatch_update(self, obs, act, oldlp, adv, ret):
        values, logp, entropy = self.net.evaluate_actions(obs, act)
        loss, pg, vf, ent = ppo_loss(values, logp, entropy, oldlp, adv, ret, self.clip_range, self.ent_coef, self.vf_coef,
                                     self.normalize_advantage)
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        if self.world > 1:                                   # data-parallel: average the 30,537 gradients over NVLink
            flat_g = torch.cat([p.grad.reshape(-1) for p in self.net.parameters()])
            dist.all_reduce(flat_g)
            flat_g /= self.world
            o = 0
            for p in self.net.parameters():
                p.grad.copy_(flat_g[o:o + p.numel()].view_as(p))
                o += p.numel()
        nn.utils.clip_grad_norm_(self.net.parameters(), self.max_grad_norm)
        self.opt.step()
        return loss.detach(), pg.detach(), vf.detach(), ent.detach()

    def _build_update_graph(self, flat, bs):
        """Capture one minibatch update over static buffers: idx -> gathers -> forward/backward -> clip -> Adam."""
        import copy
        obs, act, oldlp, adv, ret = flat
        dev = obs.device
        self._idx = torch.zeros(bs, dtype=torch.int64, device=dev)
        net_sd = copy.deepcopy(self.net.state_dict())
        opt_saved = {q: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for q, st in self.opt.state.items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                          # warm-up off the capture stream (allocations, autotuning, Adam state)
            for _ in range(3):
                self._minibatch_update(obs[self._idx], act[self._idx], oldlp[self._idx], adv[self._idx], ret[self._idx])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.net.load_state_dict(net_sd)                       # the warm-up steps must not count as training: parameters and Adam
        for q, st in self.opt.state.items():                   # state go back IN PLACE (the graph captures their addresses; state
            for k, v in st.items():                            # created lazily inside the capture would be re-zeroed by every replay)
                if torch.is_tensor(v):
                    if q in opt_saved:
                        v.copy_(opt_saved[q][k])
                    else:
                        v.zero_()
        self._graph_src = flat
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._graph_stats = self._minibatch_update(obs[self._idx], act[self._idx], oldlp[self._idx], adv[self._idx], ret[self._idx])
        self._graph = g

    def train(self) -> dict:
        T, n = self.rewards.shape
        total = T * n
        flat = lambda x: x.reshape(total, *x.shape[2:])
        obs, act, oldv, oldlp, adv, ret = map(flat, (self.obs, self.actions, self.values, self.logp, self.advantages, self.returns))
        bs = min(self.batch_size, total)
        stats = None
        if self.graph_update and self._graph is None:
            # the rollout buffers (and GAE outputs) keep their addresses between iterations, so one capture serves the whole run
            self._build_update_graph((obs, act, oldlp, adv, ret), bs)
        if self.graph_update and any(a.data_ptr() != b.data_ptr() for a, b in zip(self._graph_src, (obs, act, oldlp, adv, ret))):
            self._build_update_graph((obs, act, oldlp, adv, ret), bs)
        for epoch in range(self.n_epochs):
            perm = torch.randperm(total, device=obs.device, generator=self.gen)
            for start in range(0, total - bs + 1, bs):
                if self.graph_update:
                    self._idx.copy_(perm[start:start + bs])
                    self._graph.replay()
                    stats = self._graph_stats
                else:
                    idx = perm[start:start + bs]
                    stats = self._minibatch_update(obs[idx], act[idx], oldlp[idx], adv[idx], ret[idx])
        self.policy.params.copy_(self.net.packed())                  # the rollout kernels read the updated weights
        keys = ("loss", "policy_gradient_loss", "value_loss", "entropy_loss")
        return {k: float(v) for k, v in zip(keys, stats)}

    def learn(self, total_timesteps: int, log=None):
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            info = self.train()
            if log:
                log(dict(info, timesteps=self.num_timesteps, ep_rew_mean=self.ep_rew_mean, episodes=self.ep_count))
        return self

    def state_dict(self) -> dict:
        """SB3-named policy parameters (feed to sb3_compat.save_policy_zip)."""
        return {k: v.detach().cpu().numpy() for k, v in self.net.state_dict().items()}
