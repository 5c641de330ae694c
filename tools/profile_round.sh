#!/bin/bash
# Round profile on one B200 (run through gpurun from the repo root): tests, bench lines, ncu launch lists and full captures.
# usage: bash tools/profile_round.sh <outdir under gpurun_out/>
set -u
OUT=gpurun_out/${1:-profile}
mkdir -p $OUT
L=rl-aerial-manipulator_b200/lib
run() { echo "== $*" >> $OUT/log.txt; "$@" >> $OUT/log.txt 2>&1; echo "rc=$?" >> $OUT/log.txt; }

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log

timeout 400 python bench.py > $OUT/bench_rollout_1gpu.json 2> $OUT/bench_rollout.err
timeout 400 python bench.py --workload step --no-cpu-baseline > $OUT/bench_step_1M_f32_graph.json 2> $OUT/bench_step.err
timeout 400 python bench.py --workload step --no-cpu-baseline --graph 0 > $OUT/bench_step_1M_f32.json 2>> $OUT/bench_step.err
timeout 400 python bench.py --workload step --no-cpu-baseline --precision f64 > $OUT/bench_step_1M_f64.json 2>> $OUT/bench_step.err
timeout 400 python bench.py --workload step --no-cpu-baseline --n-envs 65536 > $OUT/bench_step_64k_f32_graph.json 2>> $OUT/bench_step.err
timeout 400 python bench.py --workload step --no-cpu-baseline --n-envs 65536 --precision f64 > $OUT/bench_step_64k_f64_graph.json 2>> $OUT/bench_step.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference.err
timeout 200 python tools/policy_time.py fp32 tensor tensor_fast > $OUT/policy_time.txt 2>&1
cat $OUT/bench_rollout_1gpu.json $OUT/bench_step_1M_f32_graph.json; cat $OUT/policy_time.txt | grep ms/forward

# launch lists (same commands as the bench lines above, which exited 0 without ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/rollout_1M_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --graph 0 > $OUT/ncu_rollout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/step_1M_f32_launches.csv \
    python bench.py --workload step --steps 40 --warmup 5 --no-cpu-baseline --graph 0 > $OUT/ncu_step.log 2>&1
# full captures of the two hot kernels
timeout 400 ncu --set full --clock-control none --import-source on -k regex:policy_forward_tc --launch-skip 3 -c 1 -o $OUT/policy_tc \
    python tools/policy_time.py tensor > $OUT/ncu_policy_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel --launch-skip 10 -c 1 -o $OUT/step_f32 \
    python bench.py --workload step --steps 40 --warmup 5 --no-cpu-baseline --graph 0 > $OUT/ncu_step_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel --launch-skip 10 -c 1 -o $OUT/step_f32_moments \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --graph 0 > $OUT/ncu_step_moments_full.log 2>&1
ls -la $OUT
