#!/bin/bash
# GPU session: parity tests, variant sweep, ncu of selected variants.
# usage: tools/gpu_sweep.sh <tag> "<1024 variants>" "<512 variants>" "<ncu variants>" [pytest-args]
set -u
TAG=${1:-rX}; V1024=${2:-}; V512=${3:-}; VNCU=${4:-}; PYT=${5:-tests -m gpu -x -q}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest"; timeout 1500 python -m pytest $PYT 2>&1 | tail -15
summ='import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; r=d["roofline"]; print("variant",c["variant"],"x",c["x"],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(r["frac"],4),"regs",c["regs_per_thread"],"ctas/sm",c["ctas_per_sm"],"kernel_ms",round(r["kernel_ms"],3),"clk",(d["clocks"] or {}).get("sm_mhz"),(d["clocks"] or {}).get("reasons"),"modes",c.get("carry_mode_histogram"))'
for v in $V1024; do
  timeout 300 python bench.py --steps 3 --warmup 3 --variant $v --no-cpu-baseline | tee gpurun_out/${TAG}_bench_v$v.json | python -c "$summ"
done
for v in $V512; do
  timeout 300 python bench.py --steps 3 --warmup 3 --x 512 --t 512 --variant $v --no-cpu-baseline | python -c "$summ"
done
for v in $VNCU; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fd1d_reg_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG}_v$v python bench.py --steps 1 --warmup 1 --variant $v --no-cpu-baseline > gpurun_out/ncu_${TAG}_v$v.log 2>&1
done
ls -la gpurun_out | tail -8
