#!/usr/bin/env python3
"""Dynamic instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv`: executed warp instructions, shared
wavefronts and stall samples per opcode (and optionally per address range)."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
S = W = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0] if src else "?"
    n = float(r[ix["Instructions Executed"]] or 0)
    wf = float(r[ix["L1 Wavefronts Shared"]] or 0)
    st = float(r[ix["# Samples"]] or 0)
    a = tot.setdefault(op, [0.0, 0.0, 0.0, 0])
    a[0] += n; a[1] += wf; a[2] += st; a[3] += 1
    S += n; W += wf
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp instructions {S:.0f}, shared wavefronts {W:.0f}")
print("| opcode | static | executed (warp) | share | per cell (thread instr) | shared wavefronts | stall samples |\n|---|---|---|---|---|---|---|")
for k, (n, wf, st, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
    pc = f"{n * 32 / cells:.1f}" if cells else "-"
    print(f"| {k} | {c} | {n:.0f} | {100 * n / S:.1f}% | {pc} | {wf:.0f} | {st:.0f} |")
