#!/bin/bash
# final source: ncu --set full of the packed-warp kernel at the reference's default grid (138, 512^2) and of the wide kernel (436, 4096^2)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fd1d_iw -s 1 -c 1 -o gpurun_out/prof_r2z_v138 python bench.py --steps 1 --warmup 1 --x 512 --t 512 --no-cpu-baseline --no-extras > gpurun_out/ncu_r2z_v138.log 2>&1
tail -1 gpurun_out/ncu_r2z_v138.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fd1d_wide_kernel -s 1 -c 1 -o gpurun_out/prof_r2z_v436 python bench.py --steps 1 --warmup 1 --x 4096 --t 4096 --n 2368 --no-cpu-baseline --no-extras > gpurun_out/ncu_r2z_v436.log 2>&1
tail -1 gpurun_out/ncu_r2z_v436.log | cut -c1-200
ls -la gpurun_out | tail -4
