#!/bin/bash
# round 2, call 14 (8 GPUs, one node): the default bench line at N=8 (env steps + PPO record), config 5 and config 4 at
# their stated shapes.  Lines kept under profiles/ (with clocks).
set -u
O=gpurun_out/r2c14
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29541 bench.py --gpus 8 --steps 64 --warmup 8 ) > $O/bench8.json 2> $O/bench8.err; echo "bench8 rc=$?"; tail -4 $O/bench8.err
( time timeout 600 $TR --master-port 29542 bench.py --gpus 8 --workload cfg5 --steps 64 --warmup 8 ) > $O/cfg5_8.json 2> $O/cfg5_8.err; echo "cfg5x8 rc=$?"; tail -4 $O/cfg5_8.err
( time timeout 1500 $TR --master-port 29543 bench.py --gpus 8 --workload ppo --ppo-model resnet --ppo-envs 32768 --ppo-horizon 128 --ppo-epochs 1 --ppo-minibatch 8192 --steps 1 --no-graph-rollout ) > $O/cfg4_8.json 2> $O/cfg4_8.err; echo "cfg4x8 rc=$?"; tail -6 $O/cfg4_8.err
python - <<'PY'
import json
def last(f):
    return json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
try:
    d = last("gpurun_out/r2c14/bench8.json")
    print("env x8: %.1fM" % (d["value"] / 1e6), "ms %.4f" % d["ms_per_step"], "e2e %.1fM" % (d["e2e"]["value"] / 1e6), d["clocks"])
    p = d["ppo"]
    print("ppo x8: %.0f" % p["value"], "rollout %.2fM step_ms %.3f update_mb_ms %.3f" % (p["rollout_samples_per_s"] / 1e6, p["rollout_step_ms"], p["update_minibatch_ms"]), p["clocks"])
except Exception as e:
    print("bench8 ERR", e)
try:
    d = last("gpurun_out/r2c14/cfg5_8.json")
    print("cfg5 x8: %.1fM" % (d["value"] / 1e6), d["config"]["finished_games_by_reason"], d["config"]["scripted_cycles_ended_by_sennichite_on_ply_13"], d["clocks"])
except Exception as e:
    print("cfg5 ERR", e)
try:
    p = last("gpurun_out/r2c14/cfg4_8.json")
    print("cfg4 x8: %.0f samples/s" % p["value"], "rollout %.2fM update %.2fM" % (p["rollout_samples_per_s"] / 1e6, p["update_samples_per_s"] / 1e6), p["tower"], p["clocks"], p["config"]["rollout_storage_gb"])
except Exception as e:
    print("cfg4 ERR", e)
PY
