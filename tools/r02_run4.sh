#!/bin/bash
set -u
OUT=gpurun_out/r02d
mkdir -p $OUT
L=rl-aerial-manipulator_b200/lib
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_policy_vecnorm.py -m gpu -x -q > $OUT/pytest_rollout.log 2>&1; echo "rc=$?" >> $OUT/pytest_rollout.log; tail -15 $OUT/pytest_rollout.log
timeout 300 python tools/rollout_time.py > $OUT/rollout_time.txt 2>&1; cat $OUT/rollout_time.txt
QS_LIB_PATH=$L/libquadsim_mmalast.so timeout 300 python tools/rollout_time.py > $OUT/rollout_time_mmalast.txt 2>&1; echo "== mma last"; grep -E "pipeline  |fused rollout step \(philox\) " $OUT/rollout_time_mmalast.txt
QS_LIB_PATH=$L/libquadsim_trace.so timeout 120 python tools/rollout_trace.py policy 16 $OUT/trace_policy.npy > $OUT/trace_policy.txt 2>&1; tail -7 $OUT/trace_policy.txt; head -2 $OUT/trace_policy.txt
QS_LIB_PATH=$L/libquadsim_trace.so timeout 120 python tools/rollout_trace.py fused 16 $OUT/trace_fused.npy > $OUT/trace_fused.txt 2>&1; tail -7 $OUT/trace_fused.txt; head -2 $OUT/trace_fused.txt
