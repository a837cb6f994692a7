#!/bin/bash
# FD1D-BS session: parity tests of the fused marches, then the timing probe.
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_bs_session.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest bs"; timeout 900 python -m pytest tests -m gpu -q -k "bs or all_1024_variants" 2>&1 | tail -15
echo "== probe"; timeout 600 python tools/bs_fused_probe.py 32768
