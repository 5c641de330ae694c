#!/bin/bash
# r02 run 2: PPO kernel tests, full GPU suite, pipeline trace, wait-hint A/B
set -u
OUT=gpurun_out/r02b
mkdir -p $OUT
L=rl-aerial-manipulator_b200/lib
timeout 600 python -m pytest tests/test_ppo.py -m gpu -x -q > $OUT/pytest_ppo.log 2>&1; echo "rc=$?" >> $OUT/pytest_ppo.log; tail -15 $OUT/pytest_ppo.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
QS_LIB_PATH=$L/libquadsim_trace.so timeout 120 python tools/rollout_trace.py policy 16 $OUT/trace_policy.npy > $OUT/trace_policy.txt 2>&1; tail -12 $OUT/trace_policy.txt
QS_LIB_PATH=$L/libquadsim_trace.so timeout 120 python tools/rollout_trace.py fused 16 $OUT/trace_fused.npy > $OUT/trace_fused.txt 2>&1; tail -12 $OUT/trace_fused.txt
for tag in hint0 hint1k; do
  QS_LIB_PATH=$L/libquadsim_$tag.so timeout 200 python tools/rollout_time.py > $OUT/rollout_time_$tag.txt 2>&1; echo "== $tag"; grep -E "pipeline  |fused rollout step \(philox\) " $OUT/rollout_time_$tag.txt
done
timeout 300 python tools/train_demo.py 8 3 2048 128 10 1 > $OUT/ppo_ref_hparams_timing.txt 2>&1; tail -4 $OUT/ppo_ref_hparams_timing.txt
