#!/bin/bash
# r02 (final build): launch list of the bench command + ncu full capture of the pipeline policy kernel
set -u
OUT=gpurun_out/r02q
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/rollout_1M_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --graph 0 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:rollout_kernel --launch-skip 3 -c 1 -f -o $OUT/rollout_policy_1M \
    python tools/rollout_probe.py policy > $OUT/ncu_policy.log 2>&1; echo "policy rc=$?"
ls -la $OUT
