#!/bin/bash
# usage: tools/gpu_quick.sh <tag> "<pytest -k expr or empty>" "<bench arg sets ';'-separated>"
set -u
TAG=${1:-rX}; KEXPR=${2:-}; EXTRA=${3:-}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}.log) 2>&1
if [ -n "$KEXPR" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -6; fi
summ='import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; r=d["roofline"]; print("variant",c["variant"],d["dtype"],"x",c["x"],"n",c["options_per_gpu"],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(r["frac"],4),"regs",c["regs_per_thread"],"ctas/sm",c["ctas_per_sm"],"kernel_ms",round(r["kernel_ms"],3),"clk",(d["clocks"] or {}).get("sm_mhz"),(d["clocks"] or {}).get("reasons"),"modes",c.get("carry_mode_histogram"))'
IFS=';' read -ra SETS <<< "$EXTRA"
for s in "${SETS[@]}"; do
  [ -z "$s" ] && continue
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $s | python -c "$summ"
done
