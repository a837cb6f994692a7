#!/usr/bin/env python
"""Post-build SASS contract of the march kernels (no GPU needed): tools/sass_contract.py [lib.so] [--dump DIR]

The Layout W march loops consume the results of tcgen05.ld (SASS: LDTM) without a tcgen05.wait::ld statement in
between (tmem.cuh: KW_TMEM_HOT_WAIT = 0).  That is safe only as long as ptxas keeps doing what it does today: every
LDTM sets a write scoreboard and the first instruction that reads (or overwrites) one of its destination registers
carries a wait on that scoreboard.  This script re-checks exactly that on the built library, instruction by
instruction, over every control-flow path that leaves an LDTM:

  1. every LDTM names a write scoreboard (bits 110-112 of the 128-bit instruction word != 7) or is followed, before
     anything touches its destination registers, by an LDTM that does (tensor-memory loads of a warp complete in
     order -- ptxas relies on the same where the source does have a wait);
  2. on every path from an LDTM, an instruction whose wait mask (bits 116-121) contains that scoreboard comes
     before, or is, the first instruction that touches a destination register of the load;
  3. the march loops (backward-branch bodies with LDTM and >= 100 DFMA / FFMA) of the dispatched kernels contain no
     local-memory store and at most four loads (LDL / STL: spills) -- a spilled march state is a silent 30 %
     regression (measured: 24 STL + 24 LDL per step took the headline kernel from 23.9 to 40.4 ms);
  4. the kernels that are supposed to keep their coefficients in tensor memory do contain LDTM and STTM.

Control word layout (upper 64-bit word u of the instruction, as printed by cuobjdump -sass on the line below the
instruction): stall = u >> 41 & 15, yield = u >> 45 & 1, write scoreboard = u >> 46 & 7, read scoreboard =
u >> 49 & 7, wait mask = u >> 52 & 63."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "kwinto-cuda_b200", "lib", "libkwfd1d.so")
MARCH_KERNELS = ("fd1d_iw_kernel", "fd1d_warp_kernel", "fd1d_wide_kernel", "fd1d_warpf_kernel", "fd1d_warp2_kernel",
                 "fd1d_warp_bs_kernel", "fd1d_reg_kernel")
# rule 3 (no spills inside the march loop) is a hard rule for the kernels the dispatch picks for full devices;
# elsewhere it is reported as a warning (the fused FD1D-BS wide kernel re-reads up to 3 loop invariants in one of its ten
# loops; experiments build: the round-1 wide kernels 331 / 431 keep 1-2 LDL per step)
NO_SPILL_KERNELS = ("fd1d_iw_kernel", "fd1d_wide_kernelILi4ELi2ELb0ELb1ELb0", "fd1d_wide_kernelILi2ELi2ELb0ELb1ELb0",
                    "fd1d_warpf_kernel")
WIDE_OPS = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "F2F.F64", "I2F.F64", "MUFU.RCP64H", "MUFU.RSQ64H")


class Ins:
    __slots__ = ("addr", "text", "op", "base", "pred", "lo", "hi", "dst", "src", "stall", "wbar", "rbar", "wait", "target",
                 "cond")


def _regs(tok):
    return [int(r) for r in re.findall(r"(?<![A-Za-z])R(\d+)\b", tok)]


def parse_function(body):
    """list of Ins for one 'Function :' block of cuobjdump -sass output"""
    out = []
    lines = body.split("\n")
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if not m:
            i += 1
            continue
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1]) if i + 1 < len(lines) else None
        ins = Ins()
        ins.addr = int(m.group(1), 16)
        ins.text = m.group(2)
        ins.lo = int(m.group(3), 16)
        ins.hi = int(m2.group(1), 16) if m2 else 0
        t = ins.text.split()
        ins.pred = None
        if t[0].startswith("@"):
            ins.pred = t[0]
            t = t[1:]
        ins.op = t[0]
        ins.base = ins.op.split(".")[0]
        ops = [o.strip() for o in " ".join(t[1:]).split(",") if o.strip()]
        u = ins.hi
        ins.stall = (u >> 41) & 15
        ins.wbar = (u >> 46) & 7
        ins.rbar = (u >> 49) & 7
        ins.wait = (u >> 52) & 63
        wide = any(ins.op.startswith(w) for w in WIDE_OPS) or ".64" in ins.op
        quad = ".128" in ins.op
        ndst = 1
        if ins.base in ("DSETP", "ISETP", "FSETP", "UISETP", "PLOP3", "HSETP2"):
            ndst = 2  # predicates: no R registers written
        if ins.base in ("STS", "STG", "STL", "ST", "BAR", "BRA", "NOP", "WARPSYNC", "BSYNC", "BSSY", "EXIT", "RET", "CALL",
                        "STTM", "RED", "MEMBAR", "FENCE", "UTCBAR", "SYNCS", "R2UR"):
            ndst = 0
        ins.dst, ins.src = set(), set()
        for k, o in enumerate(ops):
            rs = _regs(o)
            tgt = ins.dst if k < ndst else ins.src
            for r in rs:
                tgt.add(r)
                is_addr64 = "[" in o and ".64" in o  # global / generic addresses are register pairs ([R2.64+...])
                if (wide and "[" not in o) or is_addr64:
                    tgt.add(r + 1)
                if quad and k < ndst:
                    tgt.update((r + 1, r + 2, r + 3))
        if ins.base == "LDTM":
            mm = re.search(r"\.x(\d+)", ins.op)
            n = int(mm.group(1)) if mm else 1
            r0 = _regs(ops[0])[0]
            ins.dst = set(range(r0, r0 + n))
        if ins.base == "STTM":
            mm = re.search(r"\.x(\d+)", ins.op)
            n = int(mm.group(1)) if mm else 1
            rs = _regs(ops[-1])
            if rs:
                ins.src.update(range(rs[0], rs[0] + n))
        ins.target = None
        ins.cond = ins.pred is not None
        if ins.base in ("BRA", "BRX", "JMP"):
            mm = re.search(r"0x([0-9a-f]+)\s*$", ins.text)
            if mm:
                ins.target = int(mm.group(1), 16)
            if re.search(r"\b!?U?P\d+\s*,", ins.text):
                ins.cond = True
        out.append(ins)
        i += 2 if m2 else 1
    return out


def split_functions(sass_text):
    fns = {}
    for fn in re.split(r"\n\s*Function : ", sass_text)[1:]:
        name, body = fn.split("\n", 1)
        fns[name.strip()] = body
    return fns


def loops_of(code):
    """(first, last) instruction indices of every backward-branch body"""
    idx = {ins.addr: k for k, ins in enumerate(code)}
    out = []
    for k, ins in enumerate(code):
        if ins.base == "BRA" and ins.target is not None and ins.target <= ins.addr and (ins.target & 0xfffff) in idx:
            out.append((idx[ins.target & 0xfffff], k))
    return out


def check_function(name, code):
    """returns (n_ldtm, n_sttm, problems[], march_loops[(first, last, mix)])"""
    problems = []
    warnings = check_function.warnings
    idx = {ins.addr: k for k, ins in enumerate(code)}
    n_ldtm = sum(i.base == "LDTM" for i in code)
    n_sttm = sum(i.base == "STTM" for i in code)
    for k, ld in enumerate(code):
        if ld.base != "LDTM":
            continue
        # An LDTM without a scoreboard of its own is covered by the next LDTM that has one: tensor-memory loads of
        # a warp complete in order, which is also what ptxas assumes where the source has a tcgen05.wait::ld (there
        # only the last LDTM before the wait carries the scoreboard).  DFS state: (instruction index, scoreboard bit).
        stack, seen = [(k + 1, None if ld.wbar == 7 else 1 << ld.wbar)], set()
        while stack:
            j, bit = stack.pop()
            while j < len(code):
                if (j, bit) in seen:
                    break
                seen.add((j, bit))
                c = code[j]
                if bit is not None and c.wait & bit:
                    break  # this path waits before touching the destination
                if c.base != "LDTM" and ((c.src | c.dst) & ld.dst):
                    problems.append("%s @%x: %s touches a destination of the LDTM @%x (%s) without waiting for it"
                                    % (name, c.addr, c.text, ld.addr, "no scoreboard" if bit is None else "SB mask %x" % bit))
                    break
                if c.base == "LDTM":
                    if c.dst & ld.dst:
                        problems.append("%s @%x: LDTM overwrites the block of the LDTM @%x before anything waited for it"
                                        % (name, c.addr, ld.addr))
                        break
                    if bit is None and c.wbar != 7:
                        bit = 1 << c.wbar
                if c.base in ("EXIT", "RET"):
                    break
                if c.base in ("BRA", "JMP") and c.target is not None:
                    t = idx.get(c.target & 0xfffff)
                    if t is None:
                        problems.append("%s @%x: branch target outside the function while an LDTM is pending" % (name, c.addr))
                        break
                    if c.cond:
                        stack.append((t, bit))
                    else:
                        j = t
                        continue
                if c.base in ("CALL", "BRX"):
                    problems.append("%s @%x: %s while the LDTM @%x is pending" % (name, c.addr, c.base, ld.addr))
                    break
                j += 1
    march = []
    for a, b in loops_of(code):
        body = code[a:b + 1]
        mix = Counter(i.base for i in body)
        if mix["LDTM"] and mix["DFMA"] + mix["FFMA"] >= 100 and len(body) < 1500:
            march.append((a, b, mix))
            spills = [i for i in body if i.base in ("LDL", "STL")]
            if spills:
                msg = "%s: march loop @%x..%x has %d local-memory accesses (spills)" % (name, code[a].addr, code[b].addr,
                                                                                        len(spills))
                # a few LDLs of loop-invariant values are tolerated (the rare four- and five-level loops of the fused
                # FD1D-BS kernel re-read two / four scan multipliers); a store, or more than four loads, is a real spill
                # of the march's state
                stores = [i for i in spills if i.base == "STL"]
                if any(k in name for k in NO_SPILL_KERNELS) and (stores or len(spills) > 4):
                    problems.append(msg)
                else:
                    warnings.append(msg)
    return n_ldtm, n_sttm, problems, march


check_function.warnings = []


def run(lib=DEFAULT_LIB, dump_dir=None, verbose=True):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    report, problems = [], []
    for name, body in split_functions(sass).items():
        if not any(k in name for k in MARCH_KERNELS):
            continue
        code = parse_function(body)
        n_ldtm, n_sttm, probs, march = check_function(name, code)
        problems += probs
        report.append((name, len(code), n_ldtm, n_sttm, march))
        if dump_dir and march and "fd1d_iw_kernelILi4ELi2ELb0ELb0ELi1Ed" in name:
            os.makedirs(dump_dir, exist_ok=True)
            a, b, mix = min(march, key=lambda m: abs(m[2]["SHFL"] - 8))  # the two-level loop
            with open(os.path.join(dump_dir, "r2_sass_hotloop_v237_level2.txt"), "w") as f:
                f.write("# %s\n# march loop (2 scan levels) @%x..%x, %d instructions: %s\n"
                        "# columns: address, instruction, [stall, write SB, read SB, wait mask]\n"
                        % (name, code[a].addr, code[b].addr, b - a + 1, dict(mix.most_common())))
                for i in code[a:b + 1]:
                    f.write("%05x  %-58s [st %2d wr %s rd %s wait %s]\n"
                            % (i.addr, i.text, i.stall, "-" if i.wbar == 7 else i.wbar, "-" if i.rbar == 7 else i.rbar,
                               format(i.wait, "06b")))
    if verbose:
        for name, n, n_ldtm, n_sttm, march in report:
            short = re.sub(r"^_ZN6kwfd1d\d+|EvNS_9Fd1dBatchE.*$", "", name)
            print("%-70s %6d instrs  LDTM %4d STTM %3d  march loops %s"
                  % (short[:70], n, n_ldtm, n_sttm, [(b - a + 1, m["IMAD"] + m["MOV"]) for a, b, m in march]))
        for w in check_function.warnings:
            print("warning: " + w)
        print("%d problems" % len(problems))
        for p in problems[:40]:
            print("  " + p)
    return report, problems


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    dump = None
    if "--dump" in sys.argv:
        dump = sys.argv[sys.argv.index("--dump") + 1]
        args = [a for a in args if a != dump]
    _, probs = run(args[0] if args else DEFAULT_LIB, dump)
    sys.exit(1 if probs else 0)
