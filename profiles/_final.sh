set -u
O=gpurun_out/final
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_env.csv \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppo > $O/ncu_env.log 2>&1; echo "ncu env rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kz_step_kernel -s 40 -c 2 -f -o $O/kz_step \
  python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppo > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-400 $O/bench.json
