#!/bin/bash
for f in ${VARIANTS:-build/*/*.so}; do
  echo "== $f"
  KZ_LIB_PATH=$PWD/$f python bench.py --steps ${STEPS:-128} --warmup 8 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('value %.1fM  kernel_ms %.4f  frac %.3f  e2e %.1fM' % (d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']/1e6))
"
done
