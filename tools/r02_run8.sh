#!/bin/bash
# r02 run 8 (2 GPUs): the two-rank tests, the 2-GPU bench line in both launch structures
set -u
OUT=gpurun_out/r02g
mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q -k "two_ranks or two_processes or ppo_update_kernel" > $OUT/pytest_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_multi.log; tail -6 $OUT/pytest_multi.log
for mode in separate fused; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 400 --warmup 20 --rollout $mode > $OUT/bench_rollout_2gpu_$mode.json 2> $OUT/bench_2gpu_$mode.err; echo "bench $mode rc=$?"; cut -c1-330 $OUT/bench_rollout_2gpu_$mode.json
done
timeout 200 python tools/ppo_probe.py 128 > $OUT/ppo_time.txt 2>&1; timeout 100 python tools/ppo_probe.py 65536 >> $OUT/ppo_time.txt 2>&1; cut -c1-120 $OUT/ppo_time.txt
