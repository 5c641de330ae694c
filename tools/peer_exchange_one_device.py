"""Two PROCESSES on ONE GPU: the peer-memory moment exchange (qs_xchg_merge: CUDA IPC mapping, P2P stores, release/acquire flags,
bounded spin, Chan merge) against the CPU restatement oracle/sb3_oracle.merge_moments.  CUDA IPC works between processes on the
same device, NCCL does not allow two ranks on one device -- so the rendezvous is gloo and the checker is the oracle.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/peer_exchange_one_device.py
Rank 0 prints a line starting with PEER_ONE_DEVICE_OK.  (Both kernels must be co-resident while one spins on the other's flag:
two single-CTA kernels of two processes on one B200 are, under the default time-slicing as well as under MPS.)"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("gloo")
from oracle import sb3_oracle as so
from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd

D, N, STEPS = 20, 4096, 40
peer = DeviceRunningMeanStd(D, dev, exchange="peer")
assert peer.exchange == "peer", peer.exchange
g = torch.Generator(device=dev).manual_seed(100 + rank)
batches = [torch.randn((N + 64 * rank, D), device=dev, generator=g) * (1 + rank) + rank for _ in range(8)]
want = peer.stats.cpu().clone()
for i in range(STEPS):
    x = batches[i % 8]
    m = peer.batch_moments(x).clone()
    peer.update_from_moments(m)
    torch.cuda.synchronize()
    allm = [torch.empty(1 + 2 * D, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allm, m.cpu())
    want = so.merge_moments(want, torch.stack(allm))
    got = peer.stats.cpu()
    # the device kernel may contract a*b+c into FMAs, the CPU restatement does not: equal to rounding; the carried `want` follows the
    # device so that the comparison stays a per-step one
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-300), (i, (got - want).abs().max())
    want = got.clone()
assert not peer.exchange_failed()
# every rank holds the same statistics
alls = [torch.empty(1 + 2 * D, dtype=torch.float64) for _ in range(world)]
dist.all_gather(alls, peer.stats.cpu())
assert all(torch.equal(alls[0], s) for s in alls)
# back-to-back exchanges without host synchronisation in between (sequence numbers, two parity slots)
m = peer.batch_moments(batches[0]).clone()
allm = [torch.empty(1 + 2 * D, dtype=torch.float64) for _ in range(world)]
dist.all_gather(allm, m.cpu())
for _ in range(64):
    peer.update_from_moments(m)
    want = so.merge_moments(want, torch.stack(allm))
torch.cuda.synchronize()
assert torch.allclose(peer.stats.cpu(), want, rtol=1e-10, atol=1e-300) and not peer.exchange_failed()
dist.barrier()
# ---- the exchange inside the kernel that finishes the env step's observation moments (qs_step_moments_exchange): every step of a
# sharded env leaves the statistics of the WHOLE batch on every rank, one launch after the step kernel -- against the oracle's merge of
# the per-rank batch moments of the returned observations, step by step
from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv
n_env = 3000 + 96 * rank
env = BatchedQuadEnv(n_env, env_version=2, precision="f32", seed=5, env_id_offset=4096 * rank)
env.reset()
rms = DeviceRunningMeanStd(D, dev, exchange="peer")
rms.update(env.obs)
rms.attach(env, merge=True)
assert rms.env_merges
ref = DeviceRunningMeanStd(D, dev, exchange="nccl")      # only its reduction pass is used (the checker side)
ref.stats.copy_(rms.stats)
want = rms.stats.cpu().clone()
ga = torch.Generator(device=dev).manual_seed(7 + rank)
lo, span = torch.tensor([0.0, -1, -1, -1], device=dev), torch.tensor([2.0, 2, 2, 2], device=dev)
for i in range(24):
    a = lo + span * torch.rand((n_env, 4), device=dev, generator=ga)
    out = env.step(a)
    rms.update_from_moments()                        # no-op: the step has merged already
    m = ref.batch_moments(out.obs).clone()           # separate reduction pass over the same observations
    torch.cuda.synchronize()
    allm = [torch.empty(1 + 2 * D, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allm, m.cpu())
    want = so.merge_moments(want, torch.stack(allm))
    got = rms.stats.cpu()
    # the fused moments are float32 group sums re-centred in float64 (1e-6 class against the float64 reduction pass, as on one GPU)
    assert torch.allclose(got, want, rtol=2e-6, atol=1e-9), (i, (got - want).abs().max())
    want = got.clone()
assert not rms.exchange_failed()
alls = [torch.empty(1 + 2 * D, dtype=torch.float64) for _ in range(world)]
dist.all_gather(alls, rms.stats.cpu())
assert all(torch.equal(alls[0], s_) for s_ in alls)
assert float(rms.count) > 24 * (3000 + 3096) - 1
dist.barrier()
rms.close()
ref.close()
env.close()
if rank == 0:
    print(f"PEER_ONE_DEVICE_OK world={world} steps={STEPS + 64} count={float(peer.count):.1f}", flush=True)
peer.close()
sys.stdout.flush()
os._exit(0)
