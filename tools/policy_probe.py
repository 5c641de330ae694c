import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
from oracle import sb3_oracle as so
path = "tests/golden/policy_v2.npz"
IMPLS = sys.argv[1:] or ["tensor", "tensor_fast"]
for n in (128, 256, 1000, 65536, 1 << 20):
    g = torch.Generator(device="cuda").manual_seed(n)
    obs = torch.randn((n, 20), device="cuda", generator=g) * 0.7
    noise = torch.randn((n, 4), device="cuda", generator=g)
    pf = MlpPolicyKernel.from_npz(path, device="cuda", impl="fp32")
    a0, v0, l0 = [x.clone() for x in pf.forward(obs, noise)]
    m64, v64 = so.torch_policy_forward(pf.state_dict, obs[:4096], torch.float64)
    a64 = m64 + torch.exp(torch.from_numpy(pf.state_dict["log_std"]).cuda().double()) * noise[:4096].double()
    print(n, "fp32 vs f64: da", float((a0[:4096] - a64).abs().max()), "dv", float((v0[:4096] - v64).abs().max()), flush=True)
    pols = {"fp32": pf}
    for impl in IMPLS:
        pt = MlpPolicyKernel.from_npz(path, device="cuda", impl=impl)
        pols[impl] = pt
        a1, v1, l1 = [x.clone() for x in pt.forward(obs, noise)]
        torch.cuda.synchronize()
        print(n, impl, "vs fp32: max |da|", float((a0 - a1).abs().max()), "max |dv|", float((v0 - v1).abs().max()),
              "| vs f64: da", float((a1[:4096] - a64).abs().max()), "dv", float((v1[:4096] - v64).abs().max()),
              "dlogp", float((l0 - l1).abs().max()), "nan", bool(torch.isnan(a1).any() or torch.isnan(v1).any()), flush=True)
    if n >= 65536:
        for name, pol in pols.items():
            for _ in range(3): pol.forward(obs, noise)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): pol.forward(obs, noise)
            e1.record(); torch.cuda.synchronize()
            print("   ", name, n, "ms/forward", e0.elapsed_time(e1) / 20, flush=True)
