#!/usr/bin/env python
"""Per-instruction stall samples of the hot loop from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
data = rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
keys = ['stall_wait', 'stall_short_sb', 'stall_barrier', 'stall_branch_resolving', 'stall_math', 'stall_dispatch',
        'stall_not_selected', 'stall_selected', 'stall_no_inst', 'stall_long_sb', 'stall_mio']
out = []
for r in data:
    if len(r) <= ci['# Samples']:
        continue
    out.append((r[ci['Source']], int(r[ci['# Samples']] or 0), int(r[ci['Instructions Executed']] or 0),
                {k: r[ci[k]] for k in keys}))
cnt = collections.Counter(o[2] for o in out)
heavy = [c for c, _ in cnt.most_common(12) if c > 5e7]
loop = [o for o in out if o[2] in heavy]
tot = sum(o[1] for o in loop)
sel = sum(int(o[3]['stall_selected'] or 0) for o in loop) / max(1, len(loop))
print('loop instrs', len(loop), 'samples', tot, 'samples per issue ~', round(sel), '=> cycles/iter ~', round(tot / sel))
agg = collections.Counter()
for o in loop:
    for k, v in o[3].items():
        agg[k] += int(v or 0)
print({k.replace('stall_', ''): round(v / sel, 1) for k, v in agg.most_common()})
if len(sys.argv) > 2:
    for o in loop:
        st = {k.replace('stall_', ''): int(v) for k, v in o[3].items() if v not in ('0', '')}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{o[1] / sel:7.1f} {o[0][:56]:56s} {[(k, round(v / sel, 1)) for k, v in top]}")
