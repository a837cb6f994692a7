#!/bin/bash
# GPU session: device probes (FP64 peak, latencies, tensor-memory), parity tests, variant sweep.
# usage: tools/gpu_probe.sh <tag> "<1024 variants>" "<512 variants>" "<extra bench arg sets, ';'-separated>"
set -u
TAG=${1:-rX}; V1024=${2:-}; V512=${3:-}; EXTRA=${4:-}
mkdir -p gpurun_out
exec > >(tee gpurun_out/${TAG}.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== probes"; timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import kwfd1d
print("fp64_peak", kwfd1d.fp64_peak(0))
print("microbench", kwfd1d.microbench(0))
print("tmem_probe", kwfd1d.tmem_probe(0))
PY
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -40
summ='import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; r=d["roofline"]; print("variant",c["variant"],d["dtype"],"x",c["x"],"n",c["options_per_gpu"],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(r["frac"],4),"regs",c["regs_per_thread"],"ctas/sm",c["ctas_per_sm"],"kernel_ms",round(r["kernel_ms"],3),"clk",(d["clocks"] or {}).get("sm_mhz"),(d["clocks"] or {}).get("reasons"),"modes",c.get("carry_mode_histogram"))'
for v in $V1024; do
  timeout 300 python bench.py --steps 3 --warmup 3 --variant $v --no-cpu-baseline | tee gpurun_out/${TAG}_bench_v$v.json | python -c "$summ"
done
for v in $V512; do
  timeout 300 python bench.py --steps 3 --warmup 3 --x 512 --t 512 --variant $v --no-cpu-baseline | python -c "$summ"
done
IFS=';' read -ra SETS <<< "$EXTRA"
for s in "${SETS[@]}"; do
  [ -z "$s" ] && continue
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $s | python -c "$summ"
done
