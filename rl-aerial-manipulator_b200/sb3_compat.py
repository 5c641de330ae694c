"""Stable-Baselines3 file formats and wrapper surface, without importing SB3.

  QuadVecNormalize          stable_baselines3.common.vec_env.VecNormalize for a QuadVecEnv: same constructor arguments,
                            step_wait / reset / normalize_obs / unnormalize_obs / get_original_obs / save / load; the
                            running statistics live on the GPU (vec_normalize.DeviceVecNormalize)
  load_vecnormalize_pkl     read an SB3 `VecNormalize.save()` pickle (e.g. the reference's
                            initial-implementation-v1/vec_normalize.pkl) into a plain dict
  save_vecnormalize_pkl     write a pickle that real SB3 `VecNormalize.load(path, venv)` accepts: same class paths
                            (stable_baselines3.common.vec_env.vec_normalize.VecNormalize, ...running_mean_std.RunningMeanStd,
                            gymnasium.spaces.box.Box) and the same attribute dictionary as the reference's pkl
  save_policy_zip           write an SB3 model zip by replacing `policy.pth` of a template zip (PPO.load-compatible)

Reference call sites: initial-implementation-v1/rl_train_vecN.py:11,39 (VecNormalize(...), env.save("vec_normalize.pkl")),
rl_checkpoint_train_vecN.py:22-28,63-64 (VecNormalize.load), runsim_vecN.py:24-26; model zips: v2/rl_train.py:33-35,
v2/runsim_scaledObs.py:15.
"""
from __future__ import annotations

import contextlib
import io
import pickle
import sys
import types
import zipfile

import numpy as np
import torch

from .vec_env import LazyInfos, QuadVecEnv
from .vec_normalize import DeviceVecNormalize

_SB3_VN = "stable_baselines3.common.vec_env.vec_normalize"
_SB3_RMS = "stable_baselines3.common.running_mean_std"
_GYM_BOX = "gymnasium.spaces.box"


class _Bag:
    """Attribute bag standing in for an SB3 / gymnasium class while (un)pickling."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.__dict__.update(st)


# What a VecNormalize pickle may reference besides SB3 / gymnasium classes (which become attribute bags): numpy's array /
# scalar / dtype reconstructors and a few plain containers.  Anything else (os.system, builtins.eval, ...) is refused: a
# vec_normalize.pkl from a model zoo is untrusted input.
_NUMPY_OK = {"_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer", "float64", "float32", "int64", "int32", "bool_", "uint8"}
_PLAIN_OK = {("collections", "OrderedDict"), ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "slice"),
             ("builtins", "complex"), ("builtins", "bytearray"), ("_codecs", "encode")}


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        top = module.split(".")[0]
        if top in ("stable_baselines3", "gymnasium", "gym"):
            return type(name, (_Bag,), {"__module__": module})
        if (top == "numpy" and name in _NUMPY_OK) or (module, name) in _PLAIN_OK:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"vec_normalize pickle references {module}.{name}, which is not on the allowlist")


def load_vecnormalize_pkl(path: str) -> dict:
    """obs_rms / ret_rms statistics and hyper-parameters of an SB3 VecNormalize pickle."""
    with open(path, "rb") as f:
        vn = _TolerantUnpickler(f).load()
    d = vn.__dict__
    o, r = d["obs_rms"].__dict__, d["ret_rms"].__dict__
    return {"obs_mean": np.asarray(o["mean"], np.float64), "obs_var": np.asarray(o["var"], np.float64), "obs_count": float(o["count"]),
            "ret_mean": float(r["mean"]), "ret_var": float(r["var"]), "ret_count": float(r["count"]),
            "clip_obs": float(d["clip_obs"]), "clip_reward": float(d["clip_reward"]), "gamma": float(d["gamma"]),
            "epsilon": float(d["epsilon"]), "norm_obs": bool(d["norm_obs"]), "norm_reward": bool(d["norm_reward"]),
            "training": bool(d.get("training", True)), "num_envs": int(d["num_envs"])}


@contextlib.contextmanager
def _spoofed_classes():
    """Classes whose pickled path is SB3's / gymnasium's.  When the real packages are importable they are used as is."""
    made = []

    def get(module, name):
        try:
            mod = __import__(module, fromlist=[name])
            return getattr(mod, name)
        except Exception:  # noqa: BLE001 - package absent: register a stand-in under the same dotted path
            parts = module.split(".")
            for i in range(1, len(parts) + 1):
                mn = ".".join(parts[:i])
                if mn not in sys.modules:
                    sys.modules[mn] = types.ModuleType(mn)
                    made.append(mn)
            cls = type(name, (_Bag,), {"__module__": module, "__qualname__": name})
            setattr(sys.modules[module], name, cls)
            return cls
    try:
        yield get(_SB3_VN, "VecNormalize"), get(_SB3_RMS, "RunningMeanStd"), get(_GYM_BOX, "Box")
    finally:
        for mn in reversed(made):
            sys.modules.pop(mn, None)


def _box_state(low, high, dtype=np.float32) -> dict:
    low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
    rep = lambda a: str(a.flat[0]) if np.all(a == a.flat[0]) else str(a)   # gymnasium's short repr
    return {"dtype": np.dtype(dtype), "_shape": low.shape, "low": low, "bounded_below": np.isfinite(low) & (low > -np.inf),
            "high": high, "bounded_above": np.isfinite(high) & (high < np.inf), "low_repr": rep(low), "high_repr": rep(high),
            "_np_random": None}


def save_vecnormalize_pkl(path: str, state: dict, num_envs: int, obs_dim: int) -> None:
    """Pickle with SB3's VecNormalize layout (attribute set of the reference's vec_normalize.pkl)."""
    with _spoofed_classes() as (VN, RMS, Box):
        def bag(cls, d):
            o = cls.__new__(cls)
            o.__dict__.update(d)
            return o
        obs_rms = bag(RMS, {"mean": np.asarray(state["obs_mean"], np.float64), "var": np.asarray(state["obs_var"], np.float64),
                            "count": float(state["obs_count"])})
        ret_rms = bag(RMS, {"mean": np.float64(state["ret_mean"]), "var": np.float64(state["ret_var"]), "count": float(state["ret_count"])})
        inf = np.full(obs_dim, np.inf, dtype=np.float32)
        d = {"num_envs": int(num_envs), "observation_space": bag(Box, _box_state(-inf, inf)),
             "action_space": bag(Box, _box_state([0, -1, -1, -1], [2.0, 1, 1, 1])),
             "reset_infos": [{} for _ in range(num_envs)], "_seeds": [None] * num_envs, "_options": [{} for _ in range(num_envs)],
             "render_mode": None, "metadata": {"render_modes": []}, "norm_obs": bool(state.get("norm_obs", True)), "norm_obs_keys": None,
             "obs_rms": obs_rms, "ret_rms": ret_rms, "clip_obs": float(state.get("clip_obs", 10.0)),
             "clip_reward": float(state.get("clip_reward", 10.0)), "gamma": float(state.get("gamma", 0.99)),
             "epsilon": float(state.get("epsilon", 1e-8)), "training": bool(state.get("training", True)),
             "norm_reward": bool(state.get("norm_reward", False)), "old_reward": np.zeros(num_envs), "old_obs": np.zeros((num_envs, obs_dim), np.float32)}
        if not issubclass(VN, _Bag):
            # the real SB3 class: its __getstate__ deletes these three from a copy of __dict__ before pickling
            d.update(venv=None, class_attributes={}, returns=np.zeros(num_envs))
        with open(path, "wb") as f:
            pickle.dump(bag(VN, d), f, protocol=4)


def save_policy_zip(template_zip: str, out_path: str, state_dict: dict) -> None:
    """SB3 model zip = template (data json, optimizer, versions) with `policy.pth` replaced by `state_dict`."""
    buf = io.BytesIO()
    torch.save({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}, buf)
    with zipfile.ZipFile(template_zip) as zin, zipfile.ZipFile(out_path, "w", zipfile.ZIP_DEFLATED) as zout:
        for item in zin.infolist():
            zout.writestr(item, buf.getvalue() if item.filename == "policy.pth" else zin.read(item.filename))


class QuadVecNormalize:
    """SB3 `VecNormalize` semantics over a QuadVecEnv, statistics updated and applied on the device."""

    def __init__(self, venv: QuadVecEnv, training: bool = True, norm_obs: bool = True, norm_reward: bool = True,
                 clip_obs: float = 10.0, clip_reward: float = 10.0, gamma: float = 0.99, epsilon: float = 1e-8):
        self.venv = venv
        self.num_envs, self.observation_space, self.action_space = venv.num_envs, venv.observation_space, venv.action_space
        self._dev = DeviceVecNormalize(venv.sim, norm_obs=norm_obs, norm_reward=norm_reward, clip_obs=clip_obs, clip_reward=clip_reward,
                                       gamma=gamma, epsilon=epsilon, training=training)
        self.old_obs = None
        self.old_reward = None
        n, d = venv.num_envs, venv.sim.obs_dim
        self._h_raw = [torch.empty((n, d), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._h_raw_rew = [torch.empty((n,), dtype=venv.sim.real_dtype).pin_memory() for _ in range(2)]
        venv._transform = self._device_transform

    # properties mirroring SB3's attributes
    training = property(lambda self: self._dev.training, lambda self, v: setattr(self._dev, "training", bool(v)))
    norm_obs = property(lambda self: self._dev.norm_obs)
    norm_reward = property(lambda self: self._dev.norm_reward, lambda self, v: setattr(self._dev, "norm_reward", bool(v)))
    clip_obs = property(lambda self: self._dev.clip_obs)
    gamma = property(lambda self: self._dev.gamma)
    epsilon = property(lambda self: self._dev.epsilon)

    @property
    def obs_rms(self):
        st = self._dev.state_dict()
        return types.SimpleNamespace(mean=st["obs_mean"], var=st["obs_var"], count=st["obs_count"])

    @property
    def ret_rms(self):
        st = self._dev.state_dict()
        return types.SimpleNamespace(mean=st["ret_mean"], var=st["ret_var"], count=st["ret_count"])

    def _device_transform(self, out):
        """Runs on the env's stream right after the step kernel: keep the raw obs/reward for get_original_*, update the
        running statistics (device reduction; NCCL all-gather when sharded), normalise."""
        flip = self.venv._flip
        self._h_raw[flip].copy_(out.obs, non_blocking=True)
        self._h_raw_rew[flip].copy_(out.reward, non_blocking=True)
        dv = self._dev
        if dv.training:
            if dv.norm_obs:
                dv.obs_rms.update(out.obs)
            dv.update_returns(out)
        obs = dv.obs_rms.normalize(out.obs, dv.norm_obs_buf, dv.epsilon, dv.clip_obs) if dv.norm_obs else out.obs
        rew = out.reward
        if dv.norm_reward:
            rew = torch.clamp(rew / torch.sqrt(dv.ret_rms.var[0] + dv.epsilon), -dv.clip_reward, dv.clip_reward).to(out.reward.dtype)
        return obs, rew

    def reset(self):
        raw = self.venv.reset()
        self.old_obs = raw
        self._dev.returns.zero_()
        dev_obs = self.venv.sim.obs
        if self._dev.training and self._dev.norm_obs:
            self._dev.obs_rms.update(dev_obs)
        return self.normalize_obs(raw)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rewards, dones, infos = self.venv.step_wait()
        flip = self.venv._flip
        self.old_obs, self.old_reward = self._h_raw[flip].numpy(), self._h_raw_rew[flip].numpy()
        if self._dev.norm_obs:   # SB3 normalises info["terminal_observation"] as well
            ids = infos.done_indices if isinstance(infos, LazyInfos) else np.nonzero(dones)[0]
            for i in ids:
                info = infos[int(i)]
                if "terminal_observation" in info:
                    info["terminal_observation"] = self.normalize_obs(info["terminal_observation"])
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def normalize_obs(self, obs: np.ndarray) -> np.ndarray:
        if not self._dev.norm_obs:
            return obs
        st = self._dev.state_dict()
        return np.clip((obs - st["obs_mean"]) / np.sqrt(st["obs_var"] + self._dev.epsilon), -self._dev.clip_obs, self._dev.clip_obs).astype(np.float32)

    def unnormalize_obs(self, obs: np.ndarray) -> np.ndarray:
        if not self._dev.norm_obs:
            return obs
        st = self._dev.state_dict()
        return obs * np.sqrt(st["obs_var"] + self._dev.epsilon) + st["obs_mean"]

    def get_original_obs(self) -> np.ndarray:
        return np.array(self.old_obs)

    def get_original_reward(self) -> np.ndarray:
        return np.array(self.old_reward)

    def save(self, path: str) -> None:
        save_vecnormalize_pkl(path, dict(self._dev.state_dict(), training=self._dev.training), self.num_envs, self.venv.sim.obs_dim)

    @staticmethod
    def load(path: str, venv: QuadVecEnv) -> "QuadVecNormalize":
        st = load_vecnormalize_pkl(path)
        vn = QuadVecNormalize(venv, training=st["training"], norm_obs=st["norm_obs"], norm_reward=st["norm_reward"], clip_obs=st["clip_obs"],
                              clip_reward=st["clip_reward"], gamma=st["gamma"], epsilon=st["epsilon"])
        vn._dev.load_state_dict(st)
        return vn

    def close(self):
        self.venv.close()

    def __getattr__(self, name):   # VecEnvWrapper forwards everything else
        return getattr(self.venv, name)
