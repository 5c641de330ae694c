"""A few launches of the pipeline kernel for ncu: python tools/rollout_probe.py policy|fused [n_envs] [launches]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv  # noqa: E402
from rl_aerial_manipulator_b200.policy import MlpPolicyKernel  # noqa: E402
from rl_aerial_manipulator_b200.rollout import FusedRollout  # noqa: E402
from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "policy"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 6
env = BatchedQuadEnv(n, env_version=2, precision="f32", seed=0)
env.reset()
rms = DeviceRunningMeanStd(20, "cuda")
rms.update(env.obs)
rms.attach(env, merge=True)
pol = MlpPolicyKernel.from_npz(os.path.join(ROOT, "tests", "golden", "policy_v2.npz"), device="cuda",
                               impl=os.environ.get("QS_PROBE_IMPL", "tensor_pipeline"))
noise = torch.randn((n, 4), device="cuda")
if mode == "policy":
    for _ in range(launches):
        pol.forward(env.obs, noise, norm_stats=rms.stats)
else:
    fused = FusedRollout(env, pol, vecnorm=rms, sample="philox")
    for _ in range(launches):
        fused.step()
    assert fused.status() == 0
torch.cuda.synchronize()
print("ok")
