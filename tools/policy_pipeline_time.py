"""Time the pipeline policy kernel alone at 1M envs (CUDA events around CUDA-graph replays); A/B of experimental builds:
    QS_LIB_PATH=.../libquadsim_<tag>.so python tools/policy_pipeline_time.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv  # noqa: E402
from rl_aerial_manipulator_b200.policy import MlpPolicyKernel  # noqa: E402
from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd  # noqa: E402

n = 1 << 20
env = BatchedQuadEnv(n, env_version=2, precision="f32", seed=0)
env.reset()
rms = DeviceRunningMeanStd(20, "cuda")
rms.update(env.obs)
pol = MlpPolicyKernel.from_npz(os.path.join(ROOT, "tests", "golden", "policy_v2.npz"), device="cuda", impl="tensor_pipeline")
noise = torch.randn((n, 4), device="cuda")
fn = lambda: pol.forward(env.obs, noise, norm_stats=rms.stats)
for _ in range(8):
    fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(4):
        fn()
g.replay()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(25):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100)
print(f"{os.path.basename(os.environ.get('QS_LIB_PATH', 'libquadsim.so')):32s} policy pipeline + vecnorm  {best * 1e3:8.1f} us", flush=True)
