#!/bin/bash
# round 2, call 1: sanity of the round-1 tree on a fresh box + the sparse-expander split pipeline probe
set -u
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python bench.py --steps 64 --warmup 8 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python profiles/split_probe.py > $O/split_base.txt 2>&1
KZ_LIB_PATH=$PWD/build/sparse/libkeisei_b200.so timeout 300 python profiles/split_probe.py > $O/split_sparse.txt 2>&1
KZ_LIB_PATH=$PWD/build/sparse/libkeisei_b200.so QUICK=1 timeout 300 python profiles/split_probe.py > $O/split_sparse_quick.txt 2>&1
KZ_LIB_PATH=$PWD/build/sparse/libkeisei_b200.so timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k split > $O/split_sparse_test.txt 2>&1
tail -3 $O/pytest_gpu.log; cat $O/bench.json; cat $O/split_base.txt $O/split_sparse.txt $O/split_sparse_quick.txt; tail -3 $O/split_sparse_test.txt
