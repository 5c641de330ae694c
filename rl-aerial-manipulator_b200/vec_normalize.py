"""VecNormalize running statistics on the device (csrc/qs_vecnorm.cu), sharded over GPUs with one NCCL
all-gather of 2D+1 doubles per update.

Replaces stable_baselines3 `VecNormalize(env, norm_obs=True, norm_reward=False)` as used by the reference
(initial-implementation-v1/rl_train_vecN.py:11, rl_checkpoint_train_vecN.py:23-28; saved state
initial-implementation-v1/vec_normalize.pkl: obs_rms / ret_rms with mean, var, count; clip_obs=10,
gamma=0.99, epsilon=1e-8).

  DeviceRunningMeanStd   RunningMeanStd: stats tensor f64[1+2d] = (count, mean, var) living on the GPU
  DeviceVecNormalize     tensor-API wrapper around BatchedQuadEnv: step() -> normalised obs, statistics
                         updated on device, merged across ranks when torch.distributed is initialised
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._cabi import load_library


def _bind(lib):
    if getattr(lib, "_vn_bound", False):
        return
    vp, i64 = C.c_void_p, C.c_int64
    lib.qs_moments_scratch_len.argtypes = [C.c_int]
    lib.qs_moments_scratch_len.restype = i64
    lib.qs_batch_moments.argtypes = [vp, i64, C.c_int, vp, vp, vp]
    lib.qs_vecnorm_merge.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.qs_vecnorm_apply.argtypes = [vp, vp, i64, C.c_int, vp, C.c_double, C.c_double, vp]
    lib.qs_returns_update.argtypes = [vp, vp, C.c_int, vp, C.c_float, i64, vp, vp]
    for f in ("qs_batch_moments", "qs_vecnorm_merge", "qs_vecnorm_apply", "qs_returns_update"):
        getattr(lib, f).restype = C.c_int
    lib.qs_vecnorm_last_error.restype = C.c_char_p
    lib.qs_xchg_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), vp]
    lib.qs_xchg_connect.argtypes = [vp, C.c_char_p]
    lib.qs_xchg_merge.argtypes = [vp, vp, vp, vp]
    lib.qs_xchg_failed.argtypes = [vp]
    lib.qs_xchg_destroy.argtypes = [vp]
    for f in ("qs_xchg_create", "qs_xchg_connect", "qs_xchg_merge", "qs_xchg_failed", "qs_xchg_destroy"):
        getattr(lib, f).restype = C.c_int
    lib.qs_xchg_last_error.restype = C.c_char_p
    lib._vn_bound = True


class DeviceRunningMeanStd:
    """stable_baselines3.common.running_mean_std.RunningMeanStd on the device."""

    def __init__(self, dim: int, device, epsilon: float = 1e-4, group=None, exchange: str = "auto"):
        """exchange (ranks > 1): how the per-rank batch moments meet.  "peer": qs_xchg_merge, the all-gather fused with the merge
        in one kernel over NVLink peer memory; "nccl": all_gather_into_tensor + qs_vecnorm_merge; "auto": peer when every rank
        could map its peers (CUDA IPC), else nccl."""
        self.lib = load_library()
        _bind(self.lib)
        self.dim, self.device, self.group = int(dim), torch.device(device), group
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.stats = torch.zeros(1 + 2 * dim, dtype=torch.float64, device=self.device)
        self.stats[0] = epsilon
        self.stats[1 + dim:] = 1.0
        self._scratch = torch.empty(int(self.lib.qs_moments_scratch_len(dim)), dtype=torch.float64, device=self.device)
        self._moments = torch.zeros(1 + 2 * dim, dtype=torch.float64, device=self.device)
        self._gathered = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:   # pre-allocate: update() may run under graph capture
            self._gathered = torch.empty((dist.get_world_size(group), 1 + 2 * dim), dtype=torch.float64, device=self.device)
        self._xchg = None
        self.exchange = "none"
        if self._gathered is not None:
            if exchange not in ("auto", "peer", "nccl"):
                raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
            self.exchange = "nccl"
            if exchange != "nccl" and self.device.type == "cuda":
                self._connect_peers(exchange == "peer")

    def _connect_peers(self, required: bool) -> None:
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        x, handle = C.c_void_p(), (C.c_ubyte * 64)()
        rc = self.lib.qs_xchg_create(self.device.index, rank, world, self.dim, C.byref(x), handle)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if rc == 0 else None, group=self.group)
        ok = rc == 0 and all(h is not None for h in handles)
        if ok:
            ok = self.lib.qs_xchg_connect(x, b"".join(handles)) == 0
        on_host = dist.get_backend(self.group) == "gloo"                       # (two processes on one device rendezvous over gloo)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cpu" if on_host else self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)          # every rank takes the same path
        if int(flag.item()) == 1:
            self._xchg, self.exchange = x, "peer"
            return
        why = self.lib.qs_xchg_last_error().decode()
        if x:
            self.lib.qs_xchg_destroy(x)
        if required:
            raise RuntimeError(f"peer-memory exchange unavailable on some rank (this rank: {why or 'ok'})")
        import sys
        print(f"[rank {rank}] peer-memory moment exchange unavailable ({why or 'another rank failed'}); using NCCL all-gather", file=sys.stderr)

    def exchange_failed(self) -> bool:
        """True if a peer-memory merge timed out waiting for a rank (synchronises the device)."""
        return bool(self._xchg) and self.lib.qs_xchg_failed(self._xchg) != 0

    def close(self) -> None:
        if self._xchg:
            torch.cuda.synchronize(self.device)
            env = getattr(self, "_env", None)
            if env is not None and getattr(self, "env_merges", False) and getattr(env, "_h", None):
                env.fuse_obs_moments(None)          # the env handle must not keep a pointer to the exchange
                self._env = None
            dist.barrier(group=self.group)          # nobody unmaps while a peer may still store into the buffer
            self.lib.qs_xchg_destroy(self._xchg)
            self._xchg = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.qs_vecnorm_last_error().decode()}")

    @property
    def count(self):
        return self.stats[0]

    @property
    def mean(self):
        return self.stats[1:1 + self.dim]

    @property
    def var(self):
        return self.stats[1 + self.dim:]

    def batch_moments(self, x: torch.Tensor) -> torch.Tensor:
        x2 = x.reshape(x.shape[0], -1)
        assert x2.dtype == torch.float32 and x2.is_contiguous() and x2.shape[1] == self.dim
        self._check(self.lib.qs_batch_moments(C.c_void_p(x2.data_ptr()), x2.shape[0], self.dim, C.c_void_p(self._moments.data_ptr()),
                                              C.c_void_p(self._scratch.data_ptr()), self._stream()), "qs_batch_moments")
        return self._moments

    def update(self, x: torch.Tensor) -> None:
        """RunningMeanStd.update(x) for the batch sharded over all ranks of `group` (or this GPU alone)."""
        self.update_from_moments(self.batch_moments(x))

    def attach(self, env, merge: bool = False) -> None:
        """Let `env`'s step kernel produce the batch moments of the observations it returns (no separate read pass):
        after every env.step(), call update_from_moments().  merge=True: the step also merges them into these running statistics
        itself -- qs_step_moments_merge on one GPU, qs_step_moments_exchange (peer-memory all-gather + merge in the kernel that
        finishes the moments) with several ranks -- and update_from_moments() without argument becomes a no-op."""
        if merge and self._gathered is not None and not self._xchg:
            raise ValueError("merge=True with several ranks needs the peer-memory exchange (exchange='peer'): over NCCL the all-gather "
                             "is a separate call between the step and the merge")
        self.batch_moments(env.obs)                     # seeds the summation offset with the current observations' mean
        # several ranks + merge: the kernel that finishes the step's moments also runs the peer exchange and merges every rank's
        # triplet (qs_step_moments_exchange) -- one launch for what update_from_moments() does after the step otherwise
        env.fuse_obs_moments(self._moments, self.stats, merge_stats=self.stats if merge else None,
                             exchange=self._xchg if (merge and self._xchg) else None)
        self.env_merges = bool(merge)
        self._env = env

    def update_from_moments(self, m: torch.Tensor | None = None) -> None:
        """Merge a batch triplet (n, mean, M2) -- by default `self._moments`, e.g. filled by the env-step kernel
        (BatchedQuadEnv.fuse_obs_moments) -- into the running statistics, all-gathering over the ranks first."""
        if m is None:
            if getattr(self, "env_merges", False):
                return                                  # the env step already merged its batch (attach(merge=True))
            m = self._moments
        k = 1
        if self._xchg:
            rc = self.lib.qs_xchg_merge(self._xchg, C.c_void_p(self.stats.data_ptr()), C.c_void_p(m.data_ptr()), self._stream())
            if rc != 0:
                raise RuntimeError(f"qs_xchg_merge failed ({rc}): {self.lib.qs_xchg_last_error().decode()}")
            return
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            k = dist.get_world_size(self.group)
            if self._gathered is None:
                self._gathered = torch.empty((k, 1 + 2 * self.dim), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(self._gathered, m, group=self.group)   # 2d+1 doubles per rank over NVLink
            m = self._gathered
        self._check(self.lib.qs_vecnorm_merge(C.c_void_p(self.stats.data_ptr()), C.c_void_p(m.data_ptr()), k, self.dim, self._stream()),
                    "qs_vecnorm_merge")

    def normalize(self, x: torch.Tensor, out: torch.Tensor | None = None, epsilon: float = 1e-8, clip: float = 10.0) -> torch.Tensor:
        x2 = x.reshape(x.shape[0], -1)
        if out is None:
            out = torch.empty_like(x2)
        self._check(self.lib.qs_vecnorm_apply(C.c_void_p(x2.data_ptr()), C.c_void_p(out.data_ptr()), x2.shape[0], self.dim,
                                              C.c_void_p(self.stats.data_ptr()), epsilon, clip, self._stream()), "qs_vecnorm_apply")
        return out.reshape(x.shape)


class DeviceVecNormalize:
    """VecNormalize over a BatchedQuadEnv, everything on the device.

    step(actions) -> StepOut whose `.obs` is the normalised observation; `raw_obs` keeps the env's own.
    `training=False` freezes the statistics (evaluation, runsim).
    """

    def __init__(self, env, norm_obs: bool = True, norm_reward: bool = False, clip_obs: float = 10.0, clip_reward: float = 10.0,
                 gamma: float = 0.99, epsilon: float = 1e-8, training: bool = True, group=None, exchange: str = "auto"):
        self.env, self.norm_obs, self.norm_reward = env, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon, self.training = clip_obs, clip_reward, gamma, epsilon, training
        dev, n, d = env.device, env.n_envs, env.obs_dim
        self.obs_rms = DeviceRunningMeanStd(d, dev, group=group, exchange=exchange)
        self.ret_rms = DeviceRunningMeanStd(1, dev, group=group, exchange=exchange)
        self.returns = torch.zeros(n, dtype=torch.float32, device=dev)
        self._ret_snapshot = torch.zeros(n, dtype=torch.float32, device=dev)
        self.norm_obs_buf = torch.empty((n, d), dtype=torch.float32, device=dev)
        self.norm_terminal_obs = torch.empty((n, d), dtype=torch.float32, device=dev)
        self.lib = self.obs_rms.lib

    def reset(self) -> torch.Tensor:
        obs = self.env.reset()
        self.returns.zero_()
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.obs_rms.normalize(obs, self.norm_obs_buf, self.epsilon, self.clip_obs) if self.norm_obs else obs

    def step(self, actions: torch.Tensor):
        out = self.env.step(actions)
        self.raw_obs = out.obs
        if self.training:
            if self.norm_obs:
                self.obs_rms.update(out.obs)
            self.update_returns(out)
        if self.norm_obs:
            out.obs = self.obs_rms.normalize(out.obs, self.norm_obs_buf, self.epsilon, self.clip_obs)
            out.terminal_obs = self.obs_rms.normalize(out.terminal_obs, self.norm_terminal_obs, self.epsilon, self.clip_obs)
        if self.norm_reward:
            out.reward = torch.clamp(out.reward / torch.sqrt(self.ret_rms.var[0] + self.epsilon), -self.clip_reward, self.clip_reward)
        return out

    def update_returns(self, out) -> None:
        """`returns = returns*gamma + reward; ret_rms.update(returns); returns[dones] = 0` (VecNormalize.step_wait; SB3 runs
        it even with norm_reward=False -- the reference's pkl has ret_rms.count = 2031616.0001)."""
        rc = self.lib.qs_returns_update(C.c_void_p(self.returns.data_ptr()), C.c_void_p(out.reward.data_ptr()),
                                        int(out.reward.dtype == torch.float64), C.c_void_p(out.flags.data_ptr()), self.gamma,
                                        self.env.n_envs, C.c_void_p(self._ret_snapshot.data_ptr()), self.obs_rms._stream())
        self.obs_rms._check(rc, "qs_returns_update")
        self.ret_rms.update(self._ret_snapshot)

    def state_dict(self) -> dict:
        """Field names of stable_baselines3's pickled VecNormalize (see tests/golden/vecnorm_v1.npz)."""
        g = lambda t: t.detach().cpu().numpy().copy()
        return {"obs_mean": g(self.obs_rms.mean), "obs_var": g(self.obs_rms.var), "obs_count": float(self.obs_rms.count),
                "ret_mean": float(self.ret_rms.mean[0]), "ret_var": float(self.ret_rms.var[0]), "ret_count": float(self.ret_rms.count),
                "clip_obs": self.clip_obs, "clip_reward": self.clip_reward, "gamma": self.gamma, "epsilon": self.epsilon,
                "norm_obs": self.norm_obs, "norm_reward": self.norm_reward}

    def load_state_dict(self, sd: dict) -> None:
        d = self.env.obs_dim
        dev = self.env.device
        self.obs_rms.stats[0] = float(sd["obs_count"])
        self.obs_rms.stats[1:1 + d] = torch.as_tensor(sd["obs_mean"], dtype=torch.float64, device=dev)
        self.obs_rms.stats[1 + d:] = torch.as_tensor(sd["obs_var"], dtype=torch.float64, device=dev)
        self.ret_rms.stats[0] = float(sd["ret_count"])
        self.ret_rms.stats[1] = float(sd["ret_mean"])
        self.ret_rms.stats[2] = float(sd["ret_var"])
        for k in ("clip_obs", "clip_reward", "gamma", "epsilon"):
            if k in sd:
                setattr(self, k, float(sd[k]))
