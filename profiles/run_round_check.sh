#!/bin/bash
# One gpurun call that re-validates the tree on a fresh B200 and collects the profiler evidence kept under profiles/:
# GPU tests, smoke, the default bench line (env + ppo record) and the reference arm, the kernel table of the PPO path,
# ncu launch lists (env bench, one shortened PPO epoch) and `ncu --set full` captures of kz_step_kernel<1> and of the
# hand-written kernels of the PPO path.  Outputs land in gpurun_out/check/.
set -u
O=gpurun_out/check
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q --durations=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref.json 2> $O/ref.err; echo "ref rc=$?"
timeout 600 python profiles/ppo_kernels_probe.py > $O/ppo_kernels_probe.txt 2>&1; echo "probe rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_env.csv \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppo > $O/ncu_env.log 2>&1; echo "ncu env rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kz_step_kernel -s 40 -c 2 -f -o $O/kz_step \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppo > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
T=4 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"kz_(sample|eval|obs_conv|cobs_conv|adam|sumsq|gae)" -s 12 -c 60 -f -o $O/ppo_kernels \
  python profiles/ppo_kernels_probe.py > $O/ncu_ppo.log 2>&1; echo "ncu ppo rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_ppo.csv \
  python bench.py --workload ppo --steps 1 --ppo-horizon 8 --ppo-epochs 1 --no-graph-rollout > $O/ncu_ppo_list.log 2>&1; echo "ncu ppo list rc=$?"
tail -4 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench.json $O/ref.json | cut -c1-600
