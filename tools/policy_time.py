"""Time the policy forward implementations at 1M envs (CUDA events)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
n = 1 << 20
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn((n, 20), device="cuda", generator=g) * 0.7
noise = torch.randn((n, 4), device="cuda", generator=g)
for impl in sys.argv[1:] or ["tensor", "tensor_fast"]:
    pol = MlpPolicyKernel.from_npz("tests/golden/policy_v2.npz", device="cuda", impl=impl)
    for _ in range(5): pol.forward(obs, noise)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): pol.forward(obs, noise)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("QS_LIB_PATH", "default")[-24:], impl, "ms/forward", e0.elapsed_time(e1) / 50, flush=True)
