#!/bin/bash
# the historic `kwinto bench` invocation (reference Makefile:5, log/z800_1024_32768.log) through kwinto-gpu-ref:
# the reference's CPU pricer and the GPU pricer side by side on the fixture portfolio
set -u
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, "tests")
from conftest import load_golden
from test_portfolio import write_csv
g = load_golden("portfolio_fd1d")
write_csv("/tmp/portfolio_fd1d.csv", g["options"], g["quantlib"])
PY
kwinto-cuda_b200/bin/kwinto-gpu-ref bench -v -p FD1D --cpu32 --cpu64 --gpu32 --gpu64 -n 4 -b 8192 -x 512 -t 512 /tmp/portfolio_fd1d.csv 2>&1 | tee gpurun_out/r2_cli_bench_512_8192.log
kwinto-cuda_b200/bin/kwinto-gpu-ref bench -p FD1D --cpu64 --gpu32 --gpu64 --put -n 2 -b 32768 -x 1024 -t 1024 /tmp/portfolio_fd1d.csv 2>&1 | tee gpurun_out/r2_cli_bench_1024_32768.log
