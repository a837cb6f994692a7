#!/bin/bash
# round 2 session a: new SPLIT variants against the round-1 defaults (parity subset + kernel times)
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2a.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -x -q -k "236 or 136 or fixture" 2>&1 | tail -5
echo "== probe 1024"; timeout 600 python tools/variant_probe.py 1024 1024 32768 233 236
echo "== probe 512"; timeout 300 python tools/variant_probe.py 512 512 32768 133 136
