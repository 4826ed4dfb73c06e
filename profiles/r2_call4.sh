#!/bin/bash
# round 2, call 4: cooperative row writer + bitmap rollout path: GPU tests, A/B bench, PPO workload
set -u
O=gpurun_out/r2c4
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $O/pytest_gpu.log
VARIANTS="shogidrl_b200/libkeisei_b200.so build/nocoop/libkeisei_b200.so shogidrl_b200/libkeisei_b200.so" STEPS=128 bash profiles/run_variants.sh 2>&1 | tee $O/variants.txt
timeout 600 python bench.py --workload ppo --steps 2 --warmup 1 > $O/ppo.json 2> $O/ppo.err; echo "ppo rc=$?"; cat $O/ppo.json; tail -5 $O/ppo.err
