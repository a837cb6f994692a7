#!/bin/bash
# final build: ncu --set full of the headline kernel + launch list of the default bench command
set -u
mkdir -p gpurun_out
bash tools/gpu_ncu.sh r2z "237"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2z_launches_bench.log 2>&1
tail -3 gpurun_out/r2z_launches.csv
