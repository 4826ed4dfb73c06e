#!/bin/bash
# round 2, call 5: compose-once writers (default build) -- GPU tests, then A/B of the writer variants on one box
set -u
O=gpurun_out/r2c5
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 $O/pytest_gpu.log
KZ_LIB_PATH=$PWD/build/coop/libkeisei_b200.so timeout 600 python -m pytest tests/test_gpu_engine.py -q -x > $O/pytest_coop.log 2>&1; echo "coop engine tests rc=$?"; tail -3 $O/pytest_coop.log
VARIANTS="shogidrl_b200/libkeisei_b200.so build/base/libkeisei_b200.so build/once1/libkeisei_b200.so build/once2/libkeisei_b200.so build/coop/libkeisei_b200.so shogidrl_b200/libkeisei_b200.so build/base/libkeisei_b200.so" STEPS=128 bash profiles/run_variants.sh 2>&1 | tee $O/variants.txt
timeout 600 python bench.py --workload ppo --steps 2 --warmup 1 > $O/ppo.json 2> $O/ppo.err; echo "ppo rc=$?"; cat $O/ppo.json; tail -5 $O/ppo.err
