set -u
O=gpurun_out/r2c16
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29532 bench.py --gpus 2 --workload ppo --ppo-model resnet --ppo-envs 4096 --ppo-horizon 16 --ppo-epochs 1 --ppo-minibatch 2048 --steps 1 > $O/cfg4_small.json 2> $O/cfg4_small.err; echo "cfg4 small rc=$?"; tail -5 $O/cfg4_small.err
python - <<'PY'
import json
p = json.loads([l for l in open("gpurun_out/r2c16/cfg4_small.json") if l.startswith("{")][-1])
print("cfg4 small x2: %.0f" % p["value"], p["tower"], {k: round(v, 4) for k, v in p["last_metrics"].items()})
PY
