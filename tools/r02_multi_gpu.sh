#!/bin/bash
# round-2 multi-GPU check (gpurun --gpus N): the bench line on all visible GPUs + the two-rank tests
set -u
OUT=gpurun_out/r02_multi
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 400 --warmup 20 > $OUT/bench_rollout_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench $N rc=$?"; cut -c1-330 $OUT/bench_rollout_${N}gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 400 --warmup 20 --split-exchange --no-cpu-baseline > $OUT/bench_rollout_${N}gpu_split_exchange.json 2> $OUT/bench_${N}gpu_split.err; echo "bench (exchange as its own launch) $N rc=$?"; cut -c1-330 $OUT/bench_rollout_${N}gpu_split_exchange.json
timeout 300 python -m pytest tests -m gpu -q -k "two_ranks" > $OUT/pytest_two_ranks.log 2>&1; tail -3 $OUT/pytest_two_ranks.log
