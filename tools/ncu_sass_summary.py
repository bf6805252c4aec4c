#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instructions executed and stall samples per SASS opcode."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # one section per profiled launch (ncu repeats each twice)
hdr_i = his[sec]
end_i = his[sec + 1] if sec + 1 < len(his) else len(rows)
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
ops = defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
stall_cols = [h for h in hdr if h.startswith("stall_")]
stalls = defaultdict(int)
for r in rows[hdr_i + 1:end_i]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    op = sass.split()[0] if sass else "?"
    if op.startswith("@"):
        op = sass.split()[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "MUFU")) and "." in op else "")
    n = int(float(r[ix["Instructions Executed"]] or 0))
    s = int(float(r[ix["# Samples"]] or 0))
    ops[op][0] += n
    ops[op][1] += s
    ops[op][2] += 1
    tot[0] += n
    tot[1] += s
    for c in stall_cols:
        v = r[ix[c]]
        if v:
            stalls[c] += int(float(v))
print("total warp-instructions %d, samples %d" % tuple(tot))
for op, (n, s, c) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print("  %-14s inst %12d (%5.1f%%)  samples %8d (%5.1f%%)  static %d" % (op, n, 100 * n / max(tot[0], 1), s, 100 * s / max(tot[1], 1), c))
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(sum(stalls.values()), 1)) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
