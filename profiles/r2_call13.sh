#!/bin/bash
# round 2, call 13 (2 GPUs): default bench line at N=2 (graph rollout, overlapped gradient all-reduce) and a small config-4
# shaped run (ResTower + SE under DistributedDataParallel) as a sanity check before the 8-GPU call
set -u
O=gpurun_out/r2c13
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 2 --steps 64 --warmup 8 > $O/bench2.json 2> $O/bench2.err; echo "bench2 rc=$?"; tail -3 $O/bench2.err
timeout 900 $TR --master-port 29532 bench.py --gpus 2 --workload ppo --ppo-model resnet --ppo-envs 4096 --ppo-horizon 16 --ppo-epochs 1 --ppo-minibatch 2048 --steps 1 > $O/cfg4_small.json 2> $O/cfg4_small.err; echo "cfg4 small rc=$?"; tail -5 $O/cfg4_small.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c13/bench2.json").read().strip().splitlines()[-1])
print("env x2: %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6))
p = d["ppo"]
print("ppo x2: %.0f" % p["value"], "rollout %.2fM step_ms %.3f update_mb_ms %.3f" % (p["rollout_samples_per_s"] / 1e6, p["rollout_step_ms"], p["update_minibatch_ms"]), p["clocks"])
try:
    p = json.loads(open("gpurun_out/r2c13/cfg4_small.json").read().strip().splitlines()[-1])
    print("cfg4 small x2: %.0f" % p["value"], p["tower"], {k: round(v, 4) for k, v in p["last_metrics"].items()})
except Exception as e:
    print("cfg4 ERR", e)
PY
