#!/bin/bash
set -u
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_policy_vecnorm.py -m gpu -x -q > $OUT/pytest_rollout.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pytest_rollout.log
timeout 200 python tools/rollout_time.py > $OUT/rollout_time.txt 2>&1; grep -E "pipeline|fused" $OUT/rollout_time.txt
