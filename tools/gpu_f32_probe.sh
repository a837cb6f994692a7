#!/bin/bash
# fp32 march in the independent-warp kernel (1237; 1138 / 1038: two / four PDEs per warp): suite + auto dispatch + bench config 3
python -m pytest tests -q -m gpu 2>&1 | tail -5
for x in 256 512 1024; do python tools/variant_probe.py $x $x 32768 1000 | sed 's/regs.*kernel//'; done
python tools/variant_probe.py 1024 1024 32768 0 | sed 's/regs.*kernel//'
python bench.py --config 3 --no-extras 2>/dev/null | tail -1 > gpurun_out/r2u_bench_c3.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench_c3.json').read().strip().splitlines()[-1])
for r in d.get('sweep', d.get('config',{}).get('sweep',[])) or []: print(r)
print({k:d[k] for k in d if k not in ('sweep',)})
PY
