#!/bin/bash
set -u
OUT=gpurun_out/r02i
mkdir -p $OUT
L=rl-aerial-manipulator_b200/lib
timeout 300 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_policy_vecnorm.py -m gpu -x -q > $OUT/pytest_rollout.log 2>&1; echo "rc=$?" >> $OUT/pytest_rollout.log; tail -6 $OUT/pytest_rollout.log
timeout 200 python tools/rollout_time.py > $OUT/rollout_time.txt 2>&1; cat $OUT/rollout_time.txt
QS_LIB_PATH=$L/libquadsim_trace.so timeout 100 python tools/rollout_trace.py policy 16 $OUT/trace_policy.npy > $OUT/trace_policy.txt 2>&1; tail -7 $OUT/trace_policy.txt; head -2 $OUT/trace_policy.txt
QS_LIB_PATH=$L/libquadsim_trace.so timeout 100 python tools/rollout_trace.py fused 16 $OUT/trace_fused.npy > $OUT/trace_fused.txt 2>&1; tail -7 $OUT/trace_fused.txt; head -2 $OUT/trace_fused.txt
