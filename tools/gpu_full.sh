#!/bin/bash
# full GPU check: every -m gpu test, smoke, default bench
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_full.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py | tee gpurun_out/${TAG}_bench.json
