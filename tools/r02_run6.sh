#!/bin/bash
set -u
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
timeout 500 python bench.py > $OUT/bench_rollout_1gpu.json 2> $OUT/bench_rollout.err; echo "bench rc=$?"; tail -3 $OUT/bench_rollout.err; cut -c1-400 $OUT/bench_rollout_1gpu.json
timeout 300 python tools/train_demo.py 8 3 2048 128 10 1 > $OUT/ppo_ref_hparams_timing.txt 2>&1; tail -4 $OUT/ppo_ref_hparams_timing.txt
