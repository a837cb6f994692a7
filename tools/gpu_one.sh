#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 1 --n 524288 --seed 7 --no-cpu-baseline 2>&1 | tail -15 | cut -c1-600
