#!/usr/bin/env python
"""Per-instruction stall samples of one loop of an `ncu --page source --csv` dump, selected by its
execution count: tools/ncu_hot2.py <csv> <exec_count> [min_cycles_to_print]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 6.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
keys = [k for k in hdr if k.startswith('stall_')]
loop = []
for r in rows[hi + 1:]:
    if len(r) <= ci['# Samples'] or int(r[ci['Instructions Executed']] or 0) != want:
        continue
    loop.append((r[ci['Source']], int(r[ci['# Samples']] or 0), {k: int(r[ci[k]] or 0) for k in keys}))
tot = sum(o[1] for o in loop)
sel = sum(o[2]['stall_selected'] for o in loop) / len(loop)
print('loop instrs', len(loop), 'samples', tot, 'samples per issue ~', round(sel), '=> cycles/iter ~', round(tot / sel))
agg = collections.Counter()
for o in loop:
    for k, v in o[2].items():
        agg[k] += v
print({k.replace('stall_', ''): round(v / sel, 1) for k, v in agg.most_common() if v})
for idx, o in enumerate(loop):
    if o[1] / sel >= thr:
        st = {k.replace('stall_', ''): round(v / sel, 1) for k, v in o[2].items() if v}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{idx:4d} {o[1] / sel:7.1f} {o[0][:52]:52s} {top}")
