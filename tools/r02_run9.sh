#!/bin/bash
set -u
OUT=gpurun_out/r02h
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 600 python tools/train_demo.py 8 400 2048 128 10 1 > $OUT/ppo_demo_v1_ref_hparams.txt 2>&1; tail -5 $OUT/ppo_demo_v1_ref_hparams.txt
QS_PPO_BREAKDOWN=1 timeout 100 python tools/train_demo.py 8 3 2048 128 10 1 > $OUT/ppo_ref_hparams_breakdown.txt 2>&1; cat $OUT/ppo_ref_hparams_breakdown.txt
