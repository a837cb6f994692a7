#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 1 --options-per-gpu $((1048576 / N)) --seed 7 2>&1 | tail -12 | cut -c1-500 | tee gpurun_out/multi2_${N}.log
