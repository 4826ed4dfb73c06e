#!/bin/bash
# round 2, call 9 (2 GPUs): the default bench line under torchrun (env shards + PPO record with the overlapped gradient
# all-reduce), the reference arm under torchrun (rank 0 only), and the store probe's flat-fill variants
set -u
O=gpurun_out/r2c9
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > $O/bench2.json 2> $O/bench2.err; echo "bench2 rc=$?"; tail -5 $O/bench2.err; cat $O/bench2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --steps 32 --warmup 8 > $O/cfg5_2.json 2> $O/cfg5_2.err; echo "cfg5x2 rc=$?"; cat $O/cfg5_2.json
./build/probes/row_store_probe > $O/row_store_probe.txt 2>&1; cat $O/row_store_probe.txt
