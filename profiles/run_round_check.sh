#!/bin/bash
# One gpurun call that re-validates the tree on a fresh B200: GPU tests, smoke, both bench workloads, both reference
# arms, the PPO update breakdown and an ncu launch list of a shortened PPO epoch.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload ppo --steps 2 --warmup 1 > $O/ppo.json 2> $O/ppo.err; echo "ppo rc=$?"
timeout 300 python bench.py --impl reference --steps 64 --warmup 3 > $O/ref.json 2> $O/ref.err; echo "ref rc=$?"
timeout 600 python bench.py --impl reference --workload ppo --steps 2 --warmup 1 > $O/ref_ppo.json 2> $O/ref_ppo.err; echo "ref ppo rc=$?"
timeout 300 python profiles/ppo_update_probe.py > $O/ppo_update_breakdown.txt 2>&1; echo "probe rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_ppo.csv \
  python bench.py --workload ppo --steps 1 --warmup 1 --ppo-horizon 8 --ppo-epochs 1 > $O/ncu_ppo.log 2>&1; echo "ncu ppo rc=$?"
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; cat $O/bench.json $O/ppo.json $O/ref.json $O/ref_ppo.json
# env-step bench under ncu: launch list, then one full-set capture of two kz_step launches (read with ncu_summary.py)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_env.csv \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline > $O/ncu_env.log 2>&1; echo "ncu env rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kz_step_kernel -s 40 -c 2 -f -o $O/kz_step_final \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
