#!/bin/bash
# packed-warp variants (138: two PDEs per warp at x <= 512, 38: four at x <= 256): tests, fused BS, bench of configs[0]
python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== fused BS"
python tools/bs_probe.py 256 256 32768
python tools/bs_probe.py 128 128 65536
echo "== auto dispatch 512^2 / 256^2"
for n in 600 888 889 32768; do python tools/variant_probe.py 512 512 $n 0 | sed 's/regs.*kernel//'; done
for n in 1776 1778 32768; do python tools/variant_probe.py 256 256 $n 0 | sed 's/regs.*kernel//'; done
