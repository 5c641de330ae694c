"""Multi-rank check of the on-device PPO loop (BASELINE configs[4] mechanics): env shards with global env ids, VecNormalize
statistics through the peer-memory exchange, NCCL all-reduce of the 30,537 gradients.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/ppo_multi_rank_check.py
After two collect/train iterations every rank must hold bit-identical parameters and VecNormalize statistics.  Rank 0 prints a
line starting with PPO_MULTI_RANK_OK."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
# QS_ONE_DEVICE=1: every rank on GPU 0 (a single-GPU box): NCCL refuses two ranks on one device, so the collectives go through gloo
# (the gradient all-reduce is staged through the host); the VecNormalize moments still meet in qs_xchg_merge over CUDA IPC
ONE_DEVICE = os.environ.get("QS_ONE_DEVICE", "0") == "1"
if ONE_DEVICE:
    local = 0
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if ONE_DEVICE:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
from rl_aerial_manipulator_b200.ppo import QuadPPO
from rl_aerial_manipulator_b200.vec_normalize import DeviceVecNormalize

n = 8192 if ONE_DEVICE else 65536
env = BatchedQuadEnv(n, env_version=2, precision="f32", seed=0, env_id_offset=rank * n, device=local)
vn = DeviceVecNormalize(env, norm_obs=True, norm_reward=False, gamma=0.995)
ppo = QuadPPO(env, vecnorm=vn, n_steps=32, batch_size=n, n_epochs=2, seed=0)
t0 = time.time()
logs = []
ppo.learn(2 * 32 * n * world, log=logs.append)
torch.cuda.synchronize()
el = time.time() - t0
for name, t in (("params", ppo.policy.params), ("obs_rms", vn.obs_rms.stats)):
    t = t.cpu() if ONE_DEVICE else t
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t.contiguous())
    assert all(torch.equal(allt[0], x) for x in allt), f"{name} differ across ranks"
assert vn.obs_rms.exchange in ("peer", "nccl") and not vn.obs_rms.exchange_failed()
assert abs(float(vn.obs_rms.count) - (1e-4 + n * world * (1 + 2 * 32))) < 1e-2, float(vn.obs_rms.count)
if rank == 0:
    print(f"PPO_MULTI_RANK_OK world={world} timesteps={ppo.num_timesteps} exchange={vn.obs_rms.exchange} "
          f"loss={logs[-1]['loss']:.3f} wall={el:.1f}s", flush=True)
sys.stdout.flush()
os._exit(0)
