#!/bin/bash
# fp32 march in the independent-warp kernel (1237; 1138 / 1038: two / four PDEs per warp)
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32" 2>&1 | tail -4
python tools/variant_probe.py 1024 1024 32768 1237 1201
python tools/variant_probe.py 512 512 32768 1138 1101
python tools/variant_probe.py 256 256 32768 1038 1001
python tools/variant_probe.py 1024 1024 1184 1237
