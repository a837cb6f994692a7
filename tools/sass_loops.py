#!/usr/bin/env python
"""Instruction mix of every loop (backward branch) of the kernels matching a filter:
tools/sass_loops.py /tmp/kwfd1d.cubin <filter> [min_instrs]"""
import re
import subprocess
import sys
from collections import Counter

cubin, flt = sys.argv[1], sys.argv[2]
min_instrs = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
for fn in re.split(r"\n\s*Function : ", out)[1:]:
    name = fn.split("\n", 1)[0].strip()
    if flt not in name:
        continue
    lines = []
    for l in fn.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2)))
    addr = [a for a, _ in lines]
    print(name, "total instrs", len(lines))
    for i, (a, ins) in enumerate(lines):
        m = re.search(r"BRA(\.U)?\s+(!?U?P\d+,\s*)?0x([0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(3), 16)
            if tgt < a and tgt in addr:
                j = addr.index(tgt)
                if i - j + 1 < min_instrs:
                    continue
                ops = Counter()
                for _, b in lines[j:i + 1]:
                    t = b.split()
                    op = t[1] if t[0].startswith("@") else t[0]
                    ops[op.split(".")[0]] += 1
                print("  loop @%x..%x instrs %d" % (tgt, a, i - j + 1), dict(ops.most_common(24)))
