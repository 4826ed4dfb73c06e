#!/bin/bash
# round 2, call 8: full GPU suite, config 5 at size, ncu evidence (env step + PPO kernels)
set -u
O=gpurun_out/r2c8
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q --durations=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -22 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py --workload cfg5 --steps 64 --warmup 8 > $O/cfg5.json 2> $O/cfg5.err; echo "cfg5 rc=$?"; cat $O/cfg5.json; tail -3 $O/cfg5.err
timeout 600 python profiles/ppo_kernels_probe.py > $O/ppo_kernels_probe.txt 2>&1; echo "probe rc=$?"; cat $O/ppo_kernels_probe.txt
# ncu: launch list of the env bench, full capture of two kz_step launches, full capture of the PPO kernels
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_env.csv \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppo > $O/ncu_env.log 2>&1; echo "ncu env rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kz_step_kernel -s 40 -c 2 -f -o $O/kz_step_r2 \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppo > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
T=4 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kz_(sample|eval|obs_conv|adam|sumsq|gae|step)" -s 30 -c 40 -f -o $O/ppo_kernels_r2 \
  python profiles/ppo_kernels_probe.py > $O/ncu_ppo.log 2>&1; echo "ncu ppo rc=$?"; tail -3 $O/ncu_ppo.log
ls -la $O
