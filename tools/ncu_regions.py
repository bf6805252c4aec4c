#!/usr/bin/env python
"""Sample and instruction shares per code region of an `ncu --page source --csv` dump (regions = runs of SASS
instructions with the same execution count: prologue, unrolled loop bodies, tails, epilogue):
   python tools/ncu_regions.py <src.csv> [min_samples] [section]      (one section per profiled launch)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = his[sec]
end = his[sec + 1] if sec + 1 < len(his) else len(rows)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:end] if len(r) >= len(hdr)]
S = [int(float(r[ix["# Samples"]] or 0)) for r in body]
E = [int(float(r[ix["Instructions Executed"]] or 0)) for r in body]
tot, tote = sum(S), sum(E)
segs, cur = [], None
for i, (s, e) in enumerate(zip(S, E)):
    if cur is None or abs(e - cur[2]) > 0.02 * max(e, cur[2], 1):
        cur = [i, i, e, 0, 0]
        segs.append(cur)
    cur[1] = i
    cur[3] += s
    cur[4] += e
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 200
print("total samples %d, warp-instructions %d, static %d" % (tot, tote, len(body)))
for g in segs:
    if g[3] >= lim:
        print("  instr %5d-%5d  exec/instr %8d  samples %6d (%4.1f%%)  executed %10d (%4.1f%%)  samples/kinstr %.2f" % (
            g[0], g[1], g[2], g[3], 100 * g[3] / tot, g[4], 100 * g[4] / tote, 1000.0 * g[3] / max(g[4], 1)))
