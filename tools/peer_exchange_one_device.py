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
if rank == 0:
    print(f"PEER_ONE_DEVICE_OK world={world} steps={STEPS + 64} count={float(peer.count):.1f}", flush=True)
peer.close()
sys.stdout.flush()
os._exit(0)
