#!/bin/bash
# round 2, call 6: mask-once default + kz_step_range: GPU tests, A/B (streams 1 vs 2, obs-once variant), PPO workload
set -u
O=gpurun_out/r2c6
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 $O/pytest_gpu.log
run() { KZ_LIB_PATH=$PWD/$1 timeout 200 python bench.py --steps 128 --warmup 8 --no-cpu-baseline ${2:-} 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('$1 ${2:-}: value %.1fM  kernel_ms %.4f  frac %.3f  e2e %.1fM' % (d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']/1e6))
"; }
( run shogidrl_b200/libkeisei_b200.so "--step-streams 1"; run shogidrl_b200/libkeisei_b200.so "--step-streams 2"; run shogidrl_b200/libkeisei_b200.so "--step-streams 4"
  run build/once3/libkeisei_b200.so "--step-streams 1"; run build/once3/libkeisei_b200.so "--step-streams 2"
  run build/base/libkeisei_b200.so "--step-streams 1"; run shogidrl_b200/libkeisei_b200.so "--step-streams 2" ) | tee $O/variants.txt
timeout 600 python bench.py --workload ppo --steps 2 --warmup 1 > $O/ppo.json 2> $O/ppo.err; echo "ppo rc=$?"; cat $O/ppo.json; tail -5 $O/ppo.err
