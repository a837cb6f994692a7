#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> "<variants>" [extra bench args]  -- one ncu --set full capture per variant
set -u
TAG=${1:-rX}; VNCU=${2:-}; EXTRA=${3:-}
mkdir -p gpurun_out
for v in $VNCU; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fd1d_[wi] -s 1 -c 1 -o gpurun_out/prof_${TAG}_v$v python bench.py --steps 1 --warmup 1 --variant $v --no-cpu-baseline $EXTRA > gpurun_out/ncu_${TAG}_v$v.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_v$v.log
done
ls -la gpurun_out | tail -5
