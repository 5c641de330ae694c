"""Phase trace of the warp-specialised pipeline kernel (development aid; needs the -DQS_RO_TRACE=1 build):
    QS_NVCC_DEFINES="-DQS_RO_TRACE=1" QS_LIB_TAG=trace python -m rl_aerial_manipulator_b200._build
    QS_LIB_PATH=rl-aerial-manipulator_b200/lib/libquadsim_trace.so python tools/rollout_trace.py policy|fused [tiles_per_cta] [out.npy]
Lane 0 of every warp of CTA 0 logs (event, clock) pairs; this prints the merged timeline of TMEM lane quadrant 0 (one warp of every
role) in the steady state and the distribution of every hand-over latency."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv  # noqa: E402
from rl_aerial_manipulator_b200.policy import MlpPolicyKernel  # noqa: E402
from rl_aerial_manipulator_b200.rollout import FusedRollout  # noqa: E402
from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "policy"
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out_path = sys.argv[3] if len(sys.argv) > 3 else None
n = 148 * 128 * tiles
env = BatchedQuadEnv(n, env_version=2, precision="f32", seed=0)
env.reset()
rms = DeviceRunningMeanStd(20, "cuda")
rms.update(env.obs)
rms.attach(env, merge=True)
pol = MlpPolicyKernel.from_npz(os.path.join(ROOT, "tests", "golden", "policy_v2.npz"), device="cuda", impl="tensor_pipeline")
noise = torch.randn((n, 4), device="cuda")
if mode == "policy":
    for _ in range(3):
        pol.forward(env.obs, noise, norm_stats=rms.stats)
else:
    fused = FusedRollout(env, pol, vecnorm=rms, sample="philox")
    for _ in range(3):
        fused.step()
    assert fused.status() == 0
torch.cuda.synchronize()
TRACE_LEN = 2048
buf = np.zeros((32, TRACE_LEN, 2), dtype=np.uint32)
fn = getattr(pol.lib, "qs_trace_dump_policy" if mode == "policy" else "qs_trace_dump_fused")
fn.argtypes = [C.c_void_p, C.c_int]
assert fn(buf.ctypes.data, buf.nbytes) == 0
if out_path:
    np.save(out_path, buf)

NAMES = {0x10: "epi chunk begin", 0x11: "epi chunk done ", 0x13: "epi D3 ready   ", 0x14: "epi head done  ", 0x20: "mma L1 issued  ",
         0x21: "mma H1 arrived ", 0x22: "mma L2 commit  ", 0x23: "mma L1next     ", 0x24: "mma D3 free    ", 0x25: "mma H2 arrived ",
         0x26: "mma L3 commit  ", 0x30: "env stage begin", 0x31: "env stage end  ", 0x32: "env outputs rdy"}
events = []
for w in range(32):
    cnt = int(buf[w, TRACE_LEN - 1, 0])
    for k in range(min(cnt, TRACE_LEN - 1)):
        code, clk = int(buf[w, k, 0]), int(buf[w, k, 1])
        events.append((clk, w, code >> 16, (code >> 12) & 0xF, code & 0xFFF))
if not events:
    raise SystemExit("empty trace: is this the QS_RO_TRACE build (QS_LIB_PATH)?")
t_first = min(e[0] for e in events)
events = sorted(((e[0] - t_first) & 0xFFFFFFFF, *e[1:]) for e in events)
print(f"mode {mode}: {tiles} tiles per CTA, {len(events)} events, span {events[-1][0]} cycles = {events[-1][0] / tiles:.0f} cycles per tile")
warps_q0 = sorted({e[1] for e in events if e[1] % 4 == 0})
mid = events[-1][0] // 2
print(f"--- timeline, quadrant-0 warps {warps_q0}, window of 14k cycles from t = {mid}")
for t, w, ev, ch, arg in events:
    if w % 4 == 0 and mid <= t < mid + 14000:
        print(f"{t - mid:7d}  w{w:02d}  {NAMES.get(ev, hex(ev))}  ch {ch}  job/tile {arg}")


def latencies(src_ev, dst_ev, same_key, what):
    """for every dst event: time since the latest matching src event"""
    src = {}
    out = []
    for t, w, ev, ch, arg in events:
        if ev == src_ev:
            src[same_key(w, ch, arg, True)] = t
        elif ev == dst_ev:
            k = same_key(w, ch, arg, False)
            if k in src:
                out.append(t - src[k])
    if out:
        a = np.array(out)
        print(f"{what:60s} n={len(a):5d}  median {np.median(a):7.0f}  p10 {np.percentile(a, 10):7.0f}  p90 {np.percentile(a, 90):7.0f} cycles")


# slot of a warp: epilogue warps w < EPI_WARPS: policy build 16 (slot = w // 8), fused build 8 (slot = w // 4); mma warps are the last two
epi_warps = 16 if mode == "policy" else 8
slot_of = lambda w: (w // (epi_warps // 2)) if w < epi_warps else (w - (epi_warps + (4 if mode == "policy" else 8)))
quad0 = lambda w: w % 4 == 0
latencies(0x10, 0x11, lambda w, ch, a, s: (w, ch, a), "epilogue: one chunk (ld -> tanh -> split -> st -> arrive)")
latencies(0x11, 0x21, lambda w, ch, a, s: (slot_of(w), ch, a) if quad0(w) or not s else None, "hand-over: chunk done (quad 0) -> MMA warp sees H1[ch]")
latencies(0x22, 0x10, lambda w, ch, a, s: (slot_of(w), a) if s else ((slot_of(w), a) if ch in (4, 5) else None), "L2 commit issued -> epilogue starts chunk 4/5 (MMA tail + hand-over)")
latencies(0x26, 0x13, lambda w, ch, a, s: (slot_of(w), a), "L3 commit issued -> epilogue sees D3 (MMA tail + hand-over)")
latencies(0x13, 0x14, lambda w, ch, a, s: (w, a), "epilogue: head chunk(s)")
latencies(0x30, 0x31, lambda w, ch, a, s: (w, a), "env: stage one X tile")
