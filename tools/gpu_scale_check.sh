#!/bin/bash
# The driver's scaling launch, both arms, at N GPUs (default flags otherwise).
N=${1:-2}
TAG=${2:-rX}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_scale_${N}gpu.log) 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== reference arm, N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep -v Warning | tail -3
echo "== our arm, N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v Warning | tail -3
