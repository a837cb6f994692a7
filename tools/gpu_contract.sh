#!/bin/bash
# Contract check: full gpu tests, smoke, default bench (with cpu baseline), reference arm, launch list.
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}_contract.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader; nproc
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || echo SMOKE_FAILED
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench default"; timeout 900 python bench.py | tee gpurun_out/${TAG}_bench_default.json | cut -c1-3000
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/${TAG}_bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1
grep -c "fd1d_" gpurun_out/${TAG}_launches.csv || true
