#!/bin/bash
python - <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import kwfd1d
print("fp64_peak", kwfd1d.fp64_peak(0))
print("dfma_probe", kwfd1d.dfma_probe(0))
PY
