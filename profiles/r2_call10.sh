#!/bin/bash
# round 2, call 10 (2 GPUs): compact-observation input layer (tests + kernel times), PPO record at N=1 and N=2 with the
# side-stream bf16 gradient all-reduce (and the fp32 variant for comparison)
set -u
O=gpurun_out/r2c10
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log
timeout 600 python profiles/ppo_kernels_probe.py > $O/ppo_kernels_probe.txt 2>&1; cat $O/ppo_kernels_probe.txt
timeout 600 python bench.py --workload ppo --steps 2 > $O/ppo1.json 2> $O/ppo1.err; echo "ppo1 rc=$?"; tail -3 $O/ppo1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --workload ppo --steps 2 > $O/ppo2_bf16.json 2> $O/ppo2_bf16.err; echo "ppo2 bf16 rc=$?"; tail -3 $O/ppo2_bf16.err
KZ_GRAD_REDUCE_DTYPE=fp32 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --workload ppo --steps 2 > $O/ppo2_fp32.json 2> $O/ppo2_fp32.err; echo "ppo2 fp32 rc=$?"; tail -3 $O/ppo2_fp32.err
python - <<'PY'
import json
for f in ("ppo1", "ppo2_bf16", "ppo2_fp32"):
    try:
        p = json.loads(open(f"gpurun_out/r2c10/{f}.json").read().strip().splitlines()[-1])
        print(f, "samples/s %.0f" % p["value"], "rollout %.2fM" % (p["rollout_samples_per_s"] / 1e6), "step_ms %.3f" % p["rollout_step_ms"],
              "update_mb_ms %.3f" % p["update_minibatch_ms"], p["clocks"]["reasons"], {k: round(v, 4) for k, v in p["last_metrics"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
