#!/usr/bin/env python
"""Instruction mix of the hot loop (the backward-branch body with the most SHFLs) of every
fd1d_reg_kernel instantiation in a cubin: tools/sass_loop.py /tmp/kwfd1d.cubin [filter]"""
import re
import subprocess
import sys
from collections import Counter

cubin = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else "fd1d_reg_kernel"
out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)[1:]
for fn in funcs:
    name = fn.split("\n", 1)[0].strip()
    if flt not in name:
        continue
    lines = []
    for l in fn.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2)))
    addr = [a for a, _ in lines]
    best = None
    for i, (a, ins) in enumerate(lines):
        m = re.search(r"BRA(\.U)?\s+(!?U?P\d+,\s*)?0x([0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(3), 16)
            if tgt < a and tgt in addr:
                j = addr.index(tgt)
                nsh = sum("SHFL" in x for _, x in lines[j:i + 1])
                if nsh >= 20 and (best is None or (i - j) < (best[2] - best[1])):
                    best = (nsh, j, i)
    if not best:
        continue
    _, j, i = best
    body = [x for _, x in lines[j:i + 1]]
    ops = Counter()
    for b in body:
        t = b.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    short = re.sub(r"_ZN6kwfd1d15fd1d_reg_kernelI|EEvNS_9Fd1dBatchE", "", name)
    print(short, "loop instrs", len(body), dict(ops.most_common(30)))
    if len(sys.argv) > 3:
        print("\n".join(body))
