#!/bin/bash
# round 2, call 11: CUDA-graph rollout + compact-observation input layer: tests, kernel table, PPO record
set -u
O=gpurun_out/r2c12
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log
timeout 600 python profiles/ppo_kernels_probe.py > $O/ppo_kernels_probe.txt 2>&1; cat $O/ppo_kernels_probe.txt
timeout 600 python bench.py --workload ppo --steps 3 > $O/ppo1.json 2> $O/ppo1.err; echo "ppo1 rc=$?"; tail -3 $O/ppo1.err
python - <<'PY'
import json
for f in ("ppo1",):
    try:
        p = json.loads(open(f"gpurun_out/r2c12/{f}.json").read().strip().splitlines()[-1])
        print(f, "samples/s %.0f" % p["value"], "rollout %.2fM" % (p["rollout_samples_per_s"] / 1e6), "step_ms %.3f" % p["rollout_step_ms"],
              "update_mb_ms %.3f" % p["update_minibatch_ms"], p["clocks"], {k: round(v, 4) for k, v in p["last_metrics"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
