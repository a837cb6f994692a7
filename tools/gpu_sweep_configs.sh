#!/bin/bash
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== sweeps"; timeout 1500 python tools/sweep_configs.py ${2:-}
