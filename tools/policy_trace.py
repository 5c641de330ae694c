"""Phase timeline of the tcgen05 policy kernel (needs a library built with -DQS_TC_TRACE, see _build.py QS_LIB_TAG):
    QS_LIB_TAG=trace QS_NVCC_DEFINES="-DQS_TC_TRACE" python -m rl_aerial_manipulator_b200._build
    QS_LIB_PATH=.../lib/libquadsim_trace.so python tools/policy_trace.py [tensor|tensor_fast]
Prints, for threads 0 and 32 of both 128-env groups of CTA 0, the clock64 stamps of the fourth tile relative to the earliest."""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
from rl_aerial_manipulator_b200._cabi import load_library
impl = sys.argv[1] if len(sys.argv) > 1 else "tensor"
n = 1 << 20
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn((n, 20), device="cuda", generator=g) * 0.7
noise = torch.randn((n, 4), device="cuda", generator=g)
pol = MlpPolicyKernel.from_npz("tests/golden/policy_v2.npz", device="cuda", impl=impl)
for _ in range(3): pol.forward(obs, noise)
torch.cuda.synchronize()
lib = load_library()
buf = np.zeros((2, 2, 64), np.int64)
rc = lib.qs_policy_debug_trace(buf.ctypes.data_as(C.c_void_p))
assert rc == 0, rc
names = ["start"]
for net in range(2):
    for l in (1, 2, 3):
        names += [f"n{net}L{l} pre-issue", f"n{net}L{l} issued", f"n{net}L{l} mma-done", f"n{net}L{l} turn", f"n{net}L{l} epi-done", f"n{net}L{l} grp-bar"]
t0 = buf[buf > 0].min()
print(f"{'event':18s} " + " ".join(f"g{g}t{t*128:<3d}      " for g in range(2) for t in range(2)))
for i, nm in enumerate(names):
    print(f"{nm:18s} " + " ".join(f"{int(buf[g, t, i] - t0):6d} ({int(buf[g, t, i] - buf[g, t, i - 1]) if i else 0:5d})" for g in range(2) for t in range(2)))
