#!/bin/bash
# fixed cost of a launch (set-up + epilogue + tail) from two step counts
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2c.log) 2>&1
for t in 2 257 513 1024 2047; do timeout 300 python tools/variant_probe.py 1024 $t 32768 236; done
for n in 37888 1184 2368 9472; do timeout 300 python tools/variant_probe.py 1024 1024 $n 236; done
