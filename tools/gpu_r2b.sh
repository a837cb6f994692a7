#!/bin/bash
# round 2 session b: SPLIT wide kernels vs round-1 wide kernels
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2b.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -x -q -k "336 or 436 or 236" 2>&1 | tail -5
echo "== probe 4096"; timeout 600 python tools/variant_probe.py 4096 4096 2368 431 436
echo "== probe 2048"; timeout 600 python tools/variant_probe.py 2048 2048 4736 331 336
