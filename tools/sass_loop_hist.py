#!/usr/bin/env python
"""Static opcode histogram of the hottest loop of a kernel in a .so: finds the backward branch spanning the most
STG.E.EF.128 (embed) stores and prints per-opcode counts with the dispatch-cost model
(IMAD.WIDE = 4, packed fp32x2 = 2, everything else = 1 cycle).  Rarely-taken tail blocks are listed separately
(instructions between a forward `@P BRA` that skips them and its target are counted as 'skippable').
Usage: python tools/sass_loop_hist.py lib.so '<mangled-kernel-substring>' [n_float4_per_iteration]"""
import re
import subprocess
import sys
from collections import Counter

so, pat = sys.argv[1], sys.argv[2]
per = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = sass.split("Function : ")
fn = [b for b in blocks if b.startswith(pat) or pat in b.split("\n")[0]][0]
ins = []
for ln in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
# backward branches
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            j = addr_index[tgt]
            n_st = sum(1 for _, x in ins[j:i + 1] if "STG.E" in x and ".128" in x)
            if best is None or n_st > best[0] or (n_st == best[0] and (i - j) < (best[2] - best[1])):
                best = (n_st, j, i)
n_st, j, i = best
body = ins[j:i + 1]
# skippable regions: forward predicated branch inside the body
skip = [False] * len(body)
for k, (a, t) in enumerate(body):
    m = re.match(r"@!?U?P\d\s+BRA\s+0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt > a and tgt in addr_index and addr_index[tgt] - j <= len(body):
            for q in range(k + 1, addr_index[tgt] - j):
                skip[q] = True


def opname(t):
    p = t.split()
    op = p[1] if p[0].startswith("@") else p[0]
    return op


hot, cold = Counter(), Counter()
for k, (a, t) in enumerate(body):
    (cold if skip[k] else hot)[opname(t)] += 1


def cost(op):
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return 4
    if op in ("FFMA2", "FMUL2", "FADD2"):
        return 2
    return 1


tot = sum(hot.values())
cyc = sum(cost(o) * c for o, c in hot.items())
print(f"loop 0x{body[0][0]:x}..0x{body[-1][0]:x}: {len(body)} instr, {n_st} x STG.128, hot {tot} ({tot / per:.1f}/float4), "
      f"skippable {sum(cold.values())}; model cycles {cyc} ({cyc / per:.1f}/float4)")
for o, c in hot.most_common():
    print(f"  {o:24s} {c:5d}  {c / per:6.2f}/float4  cost {cost(o) * c / per:6.2f}")
