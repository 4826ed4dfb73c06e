set -u
O=gpurun_out/r2ab3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x > $O/pytest_engine.log 2>&1; echo "engine tests rc=$?"; tail -2 $O/pytest_engine.log
VARIANTS="shogidrl_b200/libkeisei_b200.so build/prev/libkeisei_b200.so shogidrl_b200/libkeisei_b200.so build/prev/libkeisei_b200.so" STEPS=128 bash profiles/run_variants.sh 2>&1 | tee $O/variants.txt
