"""2+ GPU check of the fused peer-memory moment exchange (qs_xchg_merge) against the NCCL path.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_exchange_check.py
Every rank feeds the same sequence of rank-dependent batches to two DeviceRunningMeanStd objects (exchange="peer" / "nccl"):
statistics must be bit-identical between the two paths and across ranks, eagerly and when replayed from a CUDA graph; then both
are timed with CUDA events.  Rank 0 prints one summary line starting with PEER_EXCHANGE_OK."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd

D, N = 20, 4096
peer = DeviceRunningMeanStd(D, dev, exchange="peer")
nccl = DeviceRunningMeanStd(D, dev, exchange="nccl")
assert peer.exchange == "peer" and nccl.exchange == "nccl", (peer.exchange, nccl.exchange)
g = torch.Generator(device=dev).manual_seed(100 + rank)
batches = [torch.randn((N + 64 * rank, D), device=dev, generator=g) * (1 + rank) + rank for _ in range(8)]
for i in range(50):
    x = batches[i % 8]
    peer.update(x)
    nccl.update(x)
torch.cuda.synchronize()
assert torch.equal(peer.stats, nccl.stats), (peer.stats - nccl.stats).abs().max()
allstats = [torch.empty_like(peer.stats) for _ in range(world)]
dist.all_gather(allstats, peer.stats)
assert all(torch.equal(allstats[0], s) for s in allstats), "ranks disagree"
total = sum(N + 64 * r for r in range(world)) * 50
assert abs(float(peer.count) - (1e-4 + total)) < 1e-3, float(peer.count)

# CUDA-graph replay: the sequence number lives on the device
m = peer.batch_moments(batches[0]).clone()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    peer.update_from_moments(m)          # warm-up on the capture stream
torch.cuda.current_stream().wait_stream(side)
nccl.update_from_moments(m)
torch.cuda.synchronize()
dist.barrier()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(4):
        peer.update_from_moments(m)
for _ in range(25):
    graph.replay()
for _ in range(100):
    nccl.update_from_moments(m)
torch.cuda.synchronize()
assert torch.equal(peer.stats, nccl.stats), (peer.stats - nccl.stats).abs().max()
assert not peer.exchange_failed()

def timed(fn, iters):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3

ngraph = torch.cuda.CUDAGraph()
with torch.cuda.graph(ngraph):
    for _ in range(4):
        nccl.update_from_moments(m)
us_peer = timed(graph.replay, 200) / 4
us_nccl = timed(ngraph.replay, 200) / 4
us_peer_eager = timed(lambda: peer.update_from_moments(m), 400)
us_nccl_eager = timed(lambda: nccl.update_from_moments(m), 400)
if rank == 0:
    print(f"PEER_EXCHANGE_OK world={world} us_per_exchange graph: peer {us_peer:.2f} nccl {us_nccl:.2f}  eager: peer {us_peer_eager:.2f} nccl {us_nccl_eager:.2f}", flush=True)
sys.stdout.flush()
os._exit(0)
