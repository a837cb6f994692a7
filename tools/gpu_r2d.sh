#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2d.log) 2>&1
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -x -q -k "237 or 137" 2>&1 | tail -5
echo "== probe 1024"; timeout 600 python tools/variant_probe.py 1024 1024 32768 236 237
timeout 600 python tools/variant_probe.py 1024 2 32768 236 237
echo "== probe 512"; timeout 300 python tools/variant_probe.py 512 512 32768 133 137
