"""bench.py -- env-steps/s of the batched WaypointQuadEnv step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rollout|step] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over the whole env batch.  Workloads (BASELINE.json `configs`):
  rollout  configs[3]/[4]: v2 env, 1,048,576 envs per GPU, float32: VecNormalize statistics (moments reduced inside the step
           kernel) -> MlpPolicy rollout forward on tcgen05 (weights of the reference's
           checkpoints_from_8_6M/ppo_model_2300000_steps.zip; normalise, sample, clip fused) -> env step (physics + reward +
           termination + obs) -> auto-reset.  Default.
  step     configs[2]-style: the env step alone on pre-generated uniform-random actions.
Envs are independent, so N GPUs run N shards with no data-path collective (weak scaling); the only exchange is the
VecNormalize moment triplet (41 doubles per rank and step: one fused all-gather+merge kernel over NVLink peer memory, NCCL
as fallback) and the max-over-ranks of the timing.

Printed JSON (one line, rank 0): the base contract + `roofline` (dominant kernel: the policy forward against the measured
tensor peak; `roofline_other`: the HBM-bound env step), `cpu_baseline` (the CPU oracle port of the reference step on the host
cores), `e2e` (same metric through the SB3-style VecEnv / predict calls with pinned host buffers, copies inside the timed
region), `gpu_launches`, `clocks`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ALGO_BYTES = {("v2", "f32"): 269, ("v2", "f64"): 425, ("v1", "f32"): 249}
POLICY_FLOPS = 60032                                                         # actor + critic MACs x 2, v2


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step on the host cores
# --------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(n_env):
    """Pool initializer: one v2 VecOracle per worker process (float64, SciPy LSODA like quadcopter.py:113, auto-reset)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np

    from oracle import quad_oracle as qo

    rng = np.random.default_rng(os.getpid())
    vec = qo.VecOracle("v2", n_env, lambda ids, eps: rng.random((len(ids), qo.N_UNIFORMS)), integrator="lsoda")
    vec.reset()
    _CPU.update(np=np, rng=rng, vec=vec, n=n_env, lo=np.array([0, -1, -1, -1.0]), hi=np.array([2, 1, 1, 1.0]))


def _cpu_chunk(seconds):
    """Step this worker's envs with uniform-random float32 actions for `seconds`; returns (env-steps, elapsed)."""
    np, rng, vec, n = _CPU["np"], _CPU["rng"], _CPU["vec"], _CPU["n"]
    done, t0 = 0, time.perf_counter()
    while True:
        a = (_CPU["lo"] + (_CPU["hi"] - _CPU["lo"]) * rng.random((n, 4))).astype(np.float32)
        with np.errstate(all="ignore"):
            vec.step(a)
        done += n
        if time.perf_counter() - t0 >= seconds:
            return done, time.perf_counter() - t0


class CpuArm:
    """The CPU arm: oracle port of the reference step, one process per host core (the SubprocVecEnv-style layout)."""

    def __init__(self, n_env: int = 8, procs: int | None = None):
        import multiprocessing as mp

        self.cores = procs or (os.cpu_count() or 1)
        self.n_env = n_env
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(n_env,))

    def sample(self, seconds: float):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_chunk, [seconds] * self.cores, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, total, wall, what):
        return (f"{total} env-steps of v2 (float64, scipy LSODA like quadcopter.py:113, auto-reset, uniform-random float32 actions) by "
                f"oracle/quad_oracle.py, {self.cores} processes x {self.n_env} envs, {what} (wall {wall:.1f} s)")


def cpu_baseline(seconds: float = 12.0) -> dict:
    arm = CpuArm()
    arm.sample(0.5)
    total, wall = arm.sample(seconds)
    arm.close()
    return {"value": total / wall, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": arm.describe(total, wall, f"one {seconds:.0f} s sample")}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" is one bounded sample on all host cores; the whole K + W run is sized to ~75 s
    per_step = max(0.02, min(10.0, 75.0 / max(1, args.steps + args.warmup)))
    arm = CpuArm()
    for _ in range(args.warmup):
        arm.sample(per_step)
    total, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        total += arm.sample(per_step)[0]
    wall = time.perf_counter() - t0
    arm.close()
    value = total / wall
    base = {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": arm.describe(total, wall, f"{args.steps} samples of {per_step * 1e3:.0f} ms")}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "note": "CPU oracle port of the reference step (the reference tree is not on the GPU box); "
                       "each step = one bounded sample on all host cores"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak(key: str, fallback: float) -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))[key]), f"measured (MEASURED_PEAKS.json {key})"
        except Exception:  # noqa: BLE001
            pass
    return fallback, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/roofline_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel_key)
        except Exception:  # noqa: BLE001
            return None
    return None


def workload_name(args) -> str:
    n = args.n_envs
    tag = f"{n // (1 << 20)}M" if n % (1 << 20) == 0 else str(n)
    return f"v2_{args.workload}_{tag}_{args.precision}"


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from rl_aerial_manipulator_b200.batched_env import BatchedQuadEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    n = args.n_envs                     # per GPU: weak scaling, one shard per rank
    env = BatchedQuadEnv(n, env_version=2, precision=args.precision, integrator="rk4", substeps=args.substeps,
                         device=local, env_id_offset=rank * n, seed=args.seed)
    env.reset()
    launches_per_step = 1
    policy = None
    vn = None
    if args.workload == "rollout":
        from rl_aerial_manipulator_b200.policy import MlpPolicyKernel
        from rl_aerial_manipulator_b200.vec_normalize import DeviceRunningMeanStd
        policy = MlpPolicyKernel.from_npz(os.path.join(ROOT, "tests", "golden", "policy_v2.npz"), device=dev)
        launches_per_step = 2
        if args.vecnorm:
            # VecNormalize(norm_obs=True): moments of this rank's shard -> (N>1) all-gather of 2D+1 doubles fused with the Chan
            # merge in one kernel over NVLink peer memory (qs_xchg_merge; NCCL all-gather + merge kernel if peers cannot be
            # mapped) -> normalisation fused into the policy kernel's obs load
            vn = DeviceRunningMeanStd(env.obs_dim, dev, exchange=args.vecnorm_exchange)
            # the step kernel reduces the obs it returns (no separate read pass); on one GPU the kernel that finishes the
            # moments also merges them into the running statistics, with several ranks the exchange kernel does
            vn.attach(env, merge=(world == 1))
            launches_per_step = 3 if world == 1 else 4       # env step + moments_final(+merge) [+ exchange/merge] + policy forward
    # uniform-random actions over the action box, pre-generated ring (step workload) / sampling noise (rollout)
    g = torch.Generator(device=dev).manual_seed(args.seed + rank)
    lo = torch.tensor([0.0, -1, -1, -1], device=dev)
    hi = torch.tensor([2.0, 1, 1, 1], device=dev)
    ring = [(lo + (hi - lo) * torch.rand((n, 4), device=dev, generator=g)).contiguous() for _ in range(4)]
    noise = [torch.randn((n, 4), device=dev, generator=g) for _ in range(4)] if policy else None
    act_lo, act_hi = lo, hi

    def policy_step(i):
        if vn is not None:
            vn.update_from_moments()                         # all-gather over ranks (N > 1) + Chan merge on device
        policy.forward(env.obs, noise[i & 3], norm_stats=vn.stats if vn is not None else None)   # -> policy.actions_clipped

    def one_step(i):
        if policy is None:
            env.step(ring[i & 3])
        else:
            policy_step(i)
            env.step(policy.actions_clipped)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()                                                # ranks enter the first exchange together
    for i in range(args.warmup):
        one_step(i)
    barrier()
    # CUDA-graph replay of the step loop (4 steps per graph, so the 4 action / noise buffers rotate exactly as in eager
    # mode): removes the host launch path (Python -> ctypes -> cudaLaunch, ~5 launches per step) from the critical path
    unroll = (4 if args.steps % 4 == 0 else 2 if args.steps % 2 == 0 else 1) if args.graph else 0   # exactly K steps are timed
    graph = None
    if unroll:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(unroll):
                one_step(i)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(unroll):
                one_step(i)
        graph.replay()
        barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    if graph is not None:
        for _ in range(args.steps // unroll):
            graph.replay()
    else:
        for i in range(args.steps):
            one_step(i)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n * world * args.steps / (ms_total / 1e3)

    # ---- per-kernel durations for the roofline: a separate eager pass with CUDA events around each launch (untimed) ----
    probe = 40
    step_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(probe)]
    policy_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(probe)]
    for i in range(probe):
        if policy is not None:
            if vn is not None:
                vn.update_from_moments()
            policy_events[i][0].record()
            policy.forward(env.obs, noise[i & 3], norm_stats=vn.stats if vn is not None else None)
            policy_events[i][1].record()
            step_events[i][0].record()
            env.step(policy.actions_clipped)
            step_events[i][1].record()
        else:
            step_events[i][0].record()
            env.step(ring[i & 3])
            step_events[i][1].record()
    torch.cuda.synchronize()
    step_kernel_ms = statistics.median(a.elapsed_time(b) for a, b in step_events)
    if policy is None:
        # bare-step workload: a step IS one launch of this kernel, so the graph-replayed step time is its duration without
        # the host launch gap the event-bracketed eager pass includes
        step_kernel_ms = min(step_kernel_ms, ms_total / args.steps)
    policy_kernel_ms = statistics.median(a.elapsed_time(b) for a, b in policy_events) if policy is not None else None
    if vn is not None and vn.exchange_failed():
        raise RuntimeError("peer-memory moment exchange timed out waiting for a rank; the timed region is invalid")

    # ---- end to end through the SB3-style VecEnv call: pinned host actions in, obs/reward/done out --------
    from rl_aerial_manipulator_b200.vec_env import QuadVecEnv
    env.close()
    del ring
    venv = QuadVecEnv(n, env_version=2, precision=args.precision, substeps=args.substeps, device=local, seed=args.seed,
                      env_id_offset=rank * n, info_mode="lazy")
    obs = venv.reset()
    rng = np.random.default_rng(args.seed + rank)
    host_actions = [(np.array([0, -1, -1, -1.0]) + np.array([2, 2, 2, 2.0]) * rng.random((n, 4))).astype(np.float32) for _ in range(2)]
    e2e_steps = max(3, min(args.steps, 30))

    def e2e_step(i):
        nonlocal obs
        if policy is not None:
            a = policy.predict_host(obs, stochastic=True, norm_stats=vn.stats if vn is not None else None)  # obs H2D -> forward -> actions D2H
        else:
            a = host_actions[i & 1]
        obs, rew, dones, infos = venv.step(a)
        return float(rew[0])

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = n * world * e2e_steps / e2e_s
    h2d = venv.h2d_bytes_per_step + (n * venv.sim.obs_dim * 4 if policy is not None else 0)
    d2h = venv.d2h_bytes_per_step + (n * 4 * 4 if policy is not None else 0)
    venv.close()

    if rank == 0:
        peak, peak_src = measured_peak("hbm_gbs", 6650.0)
        algo = ALGO_BYTES[("v2", args.precision)]
        achieved = algo * n / (step_kernel_ms / 1e3) / 1e9
        kname = f"env_step_kernel<{'float' if args.precision == 'f32' else 'double'},v2,rk4>"
        step_roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(kname), "peak_source": peak_src, "algorithmic_bytes_per_env_step": algo,
                         "kernel_ms": step_kernel_ms}
        roofline, roofline_other = step_roofline, []
        if policy is not None and policy_kernel_ms > step_kernel_ms:
            # the dominant kernel of this workload is the tcgen05 policy forward: algorithmic FLOPs (one float32 pass of the
            # two MLPs; the split-float16 mode issues 3x that on the tensor pipe) against the measured dense bf16 GEMM rate
            tpeak, tsrc = measured_peak("bf16_tflops_sustained", 1400.0)
            tach = POLICY_FLOPS * n / (policy_kernel_ms / 1e3) / 1e12
            roofline = {"bound": "tensor", "kernel": "policy_forward_tc3_kernel<20,split-f16>", "achieved": tach, "peak": tpeak, "unit": "TFLOP/s",
                        "frac": tach / tpeak, "traffic": ncu_traffic("policy_forward_tc_kernel"), "peak_source": tsrc,
                        "algorithmic_flops_per_env_step": POLICY_FLOPS, "kernel_ms": policy_kernel_ms,
                        "note": "epilogue-bound: 512 tanh per env at 1.25 MUFU and 9 warp instructions each; ncu: issue slots 57 %, XU pipe 53 %, tensor pipe 34 % busy (three MMA->epilogue chains per SM)"}
            roofline_other = [step_roofline]
        base = None
        if world == 1 and not args.no_cpu_baseline:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only"], capture_output=True, text=True)
            try:
                base = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:  # noqa: BLE001
                base = {"error": (r.stderr or r.stdout)[-300:]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": workload_name(args), "envs_per_gpu": n, "envs_total": n * world, "integrator": f"rk4x{args.substeps}",
                           "vecnormalize": bool(vn is not None), "cuda_graph": bool(graph is not None),
                           "actions": "policy (ppo_model_2300000_steps weights, stochastic, clipped)" if policy else "uniform-random over the action box, 4 pre-generated device buffers",
                           "l2": "working set per step (state pool + obs + actions) exceeds the 126 MB L2" if n * 185 > 126e6 else "working set fits L2; no flush between steps",
                           "parallelism": f"env-shard x{world}, no data-path collective",
                           "moment_exchange": vn.exchange if vn is not None else "none"},
                "roofline": roofline, "roofline_other": roofline_other, "cpu_baseline": base,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
                        "api": "QuadVecEnv.step(actions: np.ndarray) -> obs, rewards, dones, infos (pinned staging, info_mode=lazy)"},
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks}
        emit(line)
    if world > 1:
        # no destroy_process_group(): tearing down a communicator that a captured CUDA graph still references can block;
        # every rank has passed the final all_reduce, so leave without running the destructors
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["rollout", "step"])
    ap.add_argument("--n-envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--substeps", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--vecnorm", type=int, default=1, help="rollout workload: update VecNormalize statistics every step (NCCL all-gather when N>1)")
    ap.add_argument("--vecnorm-exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1: how ranks exchange the VecNormalize moments (peer = fused all-gather+merge kernel over NVLink peer memory)")
    ap.add_argument("--graph", type=int, default=1, help="replay the step loop from a CUDA graph (4 steps per graph)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "rollout" if os.path.exists(os.path.join(ROOT, "rl-aerial-manipulator_b200", "policy.py")) else "step"
    if args.cpu_baseline_only:
        print(json.dumps(cpu_baseline()), flush=True)
        return
    # rank 0 must print exactly ONE line on stdout: park the real stdout and send everything else (NCCL's version banner,
    # library chatter) to stderr
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
