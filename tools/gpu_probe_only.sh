#!/bin/bash
python - <<'PY'
import sys; sys.path.insert(0,'kwinto-cuda_b200')
import kwfd1d
print("dfma_probe", kwfd1d.dfma_probe(0))
PY
