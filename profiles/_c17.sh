set -u
O=gpurun_out/r2c17
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29543 bench.py --gpus 8 --workload ppo --ppo-model resnet --ppo-envs 32768 --ppo-horizon 128 --ppo-epochs 1 --ppo-minibatch 8192 --steps 1 --no-graph-rollout ) > $O/cfg4_8.json 2> $O/cfg4_8.err; echo "cfg4x8 rc=$?"; grep -v "OMP_NUM\|^\*\*\*\|^$" $O/cfg4_8.err | tail -6
python - <<'PY'
import json
p = json.loads([l for l in open("gpurun_out/r2c17/cfg4_8.json") if l.startswith("{")][-1])
print("cfg4 x8: %.0f samples/s" % p["value"], "rollout_ms %.0f update_ms %.0f" % (p["rollout_ms"], p["update_ms"]), p["tower"], p["clocks"])
PY
