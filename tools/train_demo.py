"""On-device PPO run (tools, not a test): prints ep_rew_mean per iteration.
python tools/train_demo.py [n_envs] [iters] [n_steps] [batch] [epochs] [env_version]      (QS_PPO_ENT=<ent_coef>, default 0.01)
reference settings: v1/rl_train_vecN.py: 8 .. 2048 128 10 1 (ent 0.01); v2/rl_train.py: 8 .. 2048 128 12 2 with QS_PPO_ENT=0.0005"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
from rl_aerial_manipulator_b200.ppo import QuadPPO
arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
n, iters, n_steps, batch, epochs, ver = arg(1, 16384), arg(2, 30), arg(3, 128), arg(4, 0), arg(5, 4), arg(6, 2)
env = BatchedQuadEnv(n, env_version=ver, precision="f32", seed=0)
ppo = QuadPPO(env, n_steps=n_steps, batch_size=batch or n * n_steps // 32, n_epochs=epochs, learning_rate=2e-4, ent_coef=float(os.environ.get("QS_PPO_ENT", "0.01")), seed=0)
t0 = time.time()
k = [0]
def log(d):
    k[0] += 1
    if k[0] % max(1, iters // 25) == 0 or k[0] == 1:
        print(f"steps {d['timesteps']:.3e}  ep_rew_mean {d['ep_rew_mean']:9.2f} ({d['episodes']} eps)  loss {d['loss']:10.3f}  vf {d['value_loss']:10.3f}  t {time.time()-t0:6.1f}s", flush=True)
if os.environ.get("QS_PPO_BREAKDOWN"):
    # wall-clock split of an iteration: rollout collection vs the minibatch updates (device-synchronised)
    for it in range(iters):
        torch.cuda.synchronize(); t1 = time.time()
        ppo.collect_rollouts()
        torch.cuda.synchronize(); t2 = time.time()
        info = ppo.train()
        torch.cuda.synchronize(); t3 = time.time()
        nb = epochs * -(-n_steps * n // (batch or n * n_steps // 32))
        print(f"iter {it}: collect {t2 - t1:.3f} s ({n_steps} steps), train {t3 - t2:.3f} s ({nb} minibatch updates = {1e6 * (t3 - t2) / nb:.1f} us each), "
              f"loss {info['loss']:.3f} grad_norm {info['grad_norm']:.3f} clip_fraction {info['clip_fraction']:.3f}", flush=True)
else:
    ppo.learn(iters * n_steps * n, log=log)
