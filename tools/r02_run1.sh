#!/bin/bash
# r02 run 1: GPU tests, rollout timings, ncu full captures of the pipeline kernels
set -u
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/rollout_time.py > $OUT/rollout_time.txt 2>&1; cat $OUT/rollout_time.txt
for mode in policy fused; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel --launch-skip 3 -c 1 -f -o $OUT/rollout_${mode} \
      python tools/rollout_probe.py $mode > $OUT/ncu_${mode}.log 2>&1
  tail -2 $OUT/ncu_${mode}.log
done
ls -la $OUT
