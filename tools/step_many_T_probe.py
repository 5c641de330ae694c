import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for T in (16, 32, 64, 128):
    os.environ["QS_BENCH_T"] = str(T)
    for prec in ("f32", "f64"):
        r = bench.step_extra(65536, prec, "rk4", dev, 0, 1024)
        print(T, prec, f"{r['ms_per_step']*1e3:.3f} us  {r['value']:.3e}", flush=True)
