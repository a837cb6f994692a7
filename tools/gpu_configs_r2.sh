#!/bin/bash
# configs 3 and 5 through bench.py (one GPU), launch list of the default bench
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_configs.log) 2>&1
echo "== config 5"; timeout 1200 python bench.py --config 5 --no-cpu-baseline > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; tail -2 gpurun_out/r2_bench_c5.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/r2_bench_c5.json") if l.startswith("{")][-1]
print("config 5 value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 4), "kernel", d["roofline"]["kernel"], "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]))
PY
echo "== config 3"; timeout 1500 python bench.py --config 3 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -2 gpurun_out/r2_bench_c3.err
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/r2_bench_c3.json") if l.startswith("{")][-1]
print("| x | precision | n | variant | options/s (device) | march ms | e2e pinned | e2e pageable | % FP64 peak |")
for r in d["sweep"]:
    print("| %d | %s | %d | %d | %.0f | %.3f | %.0f | %.0f | %s |" % (r["x"], r["precision"], r["n"], r["variant"], r["value"], r["kernel_ms"], r["e2e"], r["e2e_pageable"], ("%.1f" % (100 * r["frac_fp64_peak"])) if r["frac_fp64_peak"] else "-"))
PY
echo "== launch list (ncu, default bench, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0][:60]
    try: v = float(r[-1].replace(",", ""))
    except ValueError: continue
    agg[name] += v; cnt[name] += 1
tot = sum(agg.values())
for k, v in agg.most_common(8): print("%-62s launches %3d  total %12.0f ns  share %.4f" % (k, cnt[k], v, v / tot))
PY
