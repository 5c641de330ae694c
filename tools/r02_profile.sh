#!/bin/bash
# r02 profile run on one B200: launch list of the bench command, ncu captures of the kernels the bench line names
set -u
OUT=gpurun_out/r02p
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 400 python -m pytest tests/test_gpu_step_many.py tests/test_gpu_rollout.py -m gpu -q > $OUT/pytest_fixed.log 2>&1; echo "rc=$?" >> $OUT/pytest_fixed.log; tail -6 $OUT/pytest_fixed.log
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/rollout_1M_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --graph 0 > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 240 $NCU --set full --import-source on -k regex:rollout_kernel --launch-skip 3 -c 1 -f -o $OUT/rollout_policy_1M \
    python tools/rollout_probe.py policy > $OUT/ncu_policy.log 2>&1; echo "policy rc=$?"
timeout 200 $NCU --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats \
    --section SchedulerStats --section ComputeWorkloadAnalysis -k regex:rollout_kernel --launch-skip 3 -c 1 -f -o $OUT/rollout_fused_1M \
    python tools/rollout_probe.py fused > $OUT/ncu_fused.log 2>&1; echo "fused rc=$?"
timeout 240 $NCU --set full --import-source on -k regex:env_step_kernel --launch-skip 6 -c 1 -f -o $OUT/step_1M_f32_moments \
    python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-extras --graph 0 > $OUT/ncu_step.log 2>&1; echo "step rc=$?"
timeout 200 $NCU --set full --import-source on -k regex:ppo_update_kernel --launch-skip 3 -c 1 -f -o $OUT/ppo_update_B128 \
    python tools/ppo_probe.py 128 > $OUT/ncu_ppo128.log 2>&1; echo "ppo128 rc=$?"
timeout 200 $NCU --set full --import-source on -k regex:ppo_update_kernel --launch-skip 3 -c 1 -f -o $OUT/ppo_update_B65536 \
    python tools/ppo_probe.py 65536 > $OUT/ncu_ppo64k.log 2>&1; echo "ppo64k rc=$?"
timeout 120 python tools/ppo_probe.py 128 > $OUT/ppo_time.txt 2>&1; timeout 120 python tools/ppo_probe.py 1024 >> $OUT/ppo_time.txt 2>&1; timeout 120 python tools/ppo_probe.py 65536 >> $OUT/ppo_time.txt 2>&1; cat $OUT/ppo_time.txt
QS_PPO_BREAKDOWN=1 timeout 200 python tools/train_demo.py 8 3 2048 128 10 1 > $OUT/ppo_ref_hparams_breakdown.txt 2>&1; cat $OUT/ppo_ref_hparams_breakdown.txt
ls -la $OUT
