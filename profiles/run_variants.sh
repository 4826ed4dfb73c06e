#!/bin/bash
# A/B of alternative builds of the native library (KZ_LIB_PATH): optional parity tests (TESTS=1), then the config-2 bench.
# usage: VARIANTS="build/a/libkeisei_b200.so build/b/libkeisei_b200.so" STEPS=128 TESTS=1 bash profiles/run_variants.sh
for f in ${VARIANTS:-build/*/*.so}; do
  echo "== $f"
  if [ "${TESTS:-0}" = "1" ]; then
    KZ_LIB_PATH=$PWD/$f timeout 300 python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -2
  fi
  KZ_LIB_PATH=$PWD/$f timeout 120 python bench.py --steps ${STEPS:-128} --warmup 8 --no-cpu-baseline --no-ppo 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('value %.1fM  kernel_ms %.4f  frac %.3f  e2e %.1fM' % (d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']/1e6))
"
done
