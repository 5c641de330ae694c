#!/bin/bash
# round-2 validation on one B200 (run through gpurun): smoke(), the full GPU suite, the default bench line, the reference arm
set -u
OUT=gpurun_out/r02_validate
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log | cut -c1-100
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 500 python bench.py > $OUT/bench_rollout_1gpu.json 2> $OUT/bench_rollout.err; echo "bench rc=$?"; tail -3 $OUT/bench_rollout.err; cut -c1-330 $OUT/bench_rollout_1gpu.json
timeout 200 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/bench_reference_arm.json 2>$OUT/bench_ref.err; cut -c1-300 $OUT/bench_reference_arm.json
