#!/bin/bash
# One GPU session: new parity tests, then the skew / fused-BS probe.
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_session.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader; nproc
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
echo "== pytest fused bs"; timeout 900 python -m pytest tests -m gpu -q -x -k "bs" 2>&1 | tail -15
echo "== probe"; timeout 900 python tools/skew_bs_probe.py 32768
