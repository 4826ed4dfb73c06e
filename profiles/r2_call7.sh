#!/bin/bash
# round 2, call 7: new bench line (ppo record, consistent e2e, build info, real-reference legs), reference arm, obs-once A/B
set -u
O=gpurun_out/r2c7
mkdir -p $O
timeout 900 python bench.py --steps 64 --warmup 8 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
( time timeout 1200 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/ref.json 2> $O/ref.err; echo "ref rc=$?"; tail -5 $O/ref.err
run() { KZ_LIB_PATH=$PWD/$1 timeout 200 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --no-ppo ${2:-} 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('$1 ${2:-}: value %.1fM  kernel_ms %.4f  frac %.3f  e2e %.1fM' % (d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']/1e6))
"; }
( run shogidrl_b200/libkeisei_b200.so; run build/once3/libkeisei_b200.so; run build/base/libkeisei_b200.so; run shogidrl_b200/libkeisei_b200.so; run build/once3/libkeisei_b200.so ) | tee $O/variants.txt
cat $O/bench.json; cat $O/ref.json
