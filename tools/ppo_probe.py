"""A few launches of the PPO update kernel for ncu / timing: python tools/ppo_probe.py [B] [launches]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_aerial_manipulator_b200.policy import pack_params
from rl_aerial_manipulator_b200.ppo import PpoUpdateKernel, init_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 6
D, total = 20, max(4 * B, 16384)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn((total, D), device="cuda", generator=g)
act = torch.randn((total, 4), device="cuda", generator=g)
oldlp = torch.randn(total, device="cuda", generator=g) - 5
adv = torch.randn(total, device="cuda", generator=g)
ret = torch.randn(total, device="cuda", generator=g) * 10
params = torch.from_numpy(pack_params(init_state_dict(D, 0), D)).cuda()
opt = PpoUpdateKernel(params, D, learning_rate=2e-4, ent_coef=0.01)
idx = torch.randperm(total, device="cuda", generator=g)[:B].contiguous()
for _ in range(launches):
    opt.update(obs, act, oldlp, adv, ret, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 200
e0.record()
for _ in range(reps):
    opt.update(obs, act, oldlp, adv, ret, idx)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print(f"qs_ppo_update B={B}: {us:.1f} us per minibatch update (stream of {reps} launches), {B * 90.0e3 * 2 / us / 1e6:.2f} TFLOP/s (forward + backward ~ 3 x 30k MAC per row)", opt.read_stats())
