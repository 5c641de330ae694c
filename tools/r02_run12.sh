#!/bin/bash
set -u
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 500 python bench.py > $OUT/bench_rollout_1gpu.json 2> $OUT/bench_rollout.err; echo "bench rc=$?"; tail -3 $OUT/bench_rollout.err; cut -c1-330 $OUT/bench_rollout_1gpu.json
timeout 240 ncu --clock-control none --set full --import-source on -k regex:rollout_kernel --launch-skip 3 -c 1 -f -o $OUT/rollout_policy_1M \
    python tools/rollout_probe.py policy > $OUT/ncu_policy.log 2>&1; echo "ncu policy rc=$?"
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/rollout_1M_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --graph 0 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
