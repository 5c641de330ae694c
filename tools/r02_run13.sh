#!/bin/bash
set -u
OUT=gpurun_out/r02k
mkdir -p $OUT
L=rl-aerial-manipulator_b200/lib
for tag in ldsplit f16x4; do
  QS_LIB_PATH=$L/libquadsim_$tag.so timeout 300 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q > $OUT/pytest_$tag.log 2>&1; echo "$tag tests rc=$?"; tail -3 $OUT/pytest_$tag.log
  QS_LIB_PATH=$L/libquadsim_$tag.so timeout 200 python tools/rollout_time.py > $OUT/rollout_time_$tag.txt 2>&1; grep -E "pipeline|fused" $OUT/rollout_time_$tag.txt
done
