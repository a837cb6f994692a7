#!/bin/bash
# First GPU session: smoke, parity tests, micro-probes, variant sweep, bench, ncu evidence.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/first.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
echo "== smoke"; timeout 300 python __graft_entry__.py smoke || echo SMOKE_FAILED
echo "== probes"; timeout 120 python - <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import kwfd1d
print("fp64_peak", kwfd1d.fp64_peak(0))
print("microbench", kwfd1d.microbench(0))
PY
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== variant sweep (1024^2, 32768 options)"
for v in 230 231 241 240; do
  timeout 300 python bench.py --steps 3 --warmup 3 --variant $v --no-cpu-baseline | tee gpurun_out/bench_v$v.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('variant',d['config']['variant'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],4),'regs',d['config']['regs_per_thread'],'ctas/sm',d['config']['ctas_per_sm'],'kernel_ms',round(d['roofline']['kernel_ms'],3),'clk',d['clocks'])"
done
echo "== 512^2 variants"
for v in 160 181 180; do
  timeout 300 python bench.py --steps 3 --warmup 3 --x 512 --t 512 --variant $v --no-cpu-baseline | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('variant',d['config']['variant'],'value',round(d['value']),'frac',round(d['roofline']['frac'],4),'regs',d['config']['regs_per_thread'],'ctas/sm',d['config']['ctas_per_sm'])"
done
echo "== layout A (soa) 1024^2"
timeout 600 python bench.py --steps 2 --warmup 1 --layout soa --no-cpu-baseline | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('soa value',round(d['value']),'frac',round(d['roofline']['frac'],4),'kernel_ms',d['roofline']['kernel_ms'])"
echo "== default bench"; timeout 900 python bench.py | tee gpurun_out/bench_default.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fd1d_reg_kernel -s 1 -c 1 -o gpurun_out/prof_r1_reg python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
