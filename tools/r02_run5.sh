#!/bin/bash
# r02 run 5: full GPU suite on the reverted kernel + new tests, bench line
set -u
OUT=gpurun_out/r02e
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_step_many.py -m gpu -x -q > $OUT/pytest_step_many.log 2>&1; echo "rc=$?" >> $OUT/pytest_step_many.log; tail -12 $OUT/pytest_step_many.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_step_many.py > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
timeout 500 python bench.py > $OUT/bench_rollout_1gpu.json 2> $OUT/bench_rollout.err; echo "bench rc=$?"; tail -3 $OUT/bench_rollout.err; cat $OUT/bench_rollout_1gpu.json
