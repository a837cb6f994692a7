#!/bin/bash
# usage (gpurun --gpus N): tools/gpu_multi.sh <tag> <N>  -- sharded parity check + bench at N GPUs
set -u
TAG=${1:-rX}; N=${2:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_multi.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== sharded parity (NCCL gather)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_sharded_check.py 2>&1 | grep -E "rank|Error|error" | head -20
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -E "^\{" | tee gpurun_out/${TAG}_bench_${N}gpu.json | cut -c1-700
echo "== bench --gpus $N, config 4: 1M options in total"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --options-per-gpu $((1048576 / N)) --seed 7 2>&1 | grep -E "^\{" | tee gpurun_out/${TAG}_bench_${N}gpu_1M.json | cut -c1-400
