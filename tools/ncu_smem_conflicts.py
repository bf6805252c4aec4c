#!/usr/bin/env python
"""Shared-memory wavefronts per SASS instruction from `ncu --page source --csv` output: actual vs ideal."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # one section per profiled launch (ncu repeats each twice)
hi = his[sec]
end = his[sec + 1] if sec + 1 < len(his) else len(rows)
data = [r for r in rows[hi + 1:end] if len(r) >= len(hdr)]


def f(r, k):
    v = r[ix[k]]
    return float(v) if v else 0.0


def opcode(src):
    s = src.split()
    return s[1] if s[0].startswith("@") else s[0]


agg = {}
for r in data:
    w = f(r, "L1 Wavefronts Shared")
    if w > 0:
        a = agg.setdefault(opcode(r[ix["Source"]]), [0, 0, 0, 0])
        a[0] += w
        a[1] += f(r, "L1 Wavefronts Shared Ideal")
        a[2] += f(r, "Instructions Executed")
        a[3] += 1
print("total shared wavefronts", sum(a[0] for a in agg.values()))
print("opcode: wavefronts ideal warp-instr static")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(" ", k, a)
c = collections.Counter()
for r in data:
    w = f(r, "L1 Wavefronts Shared")
    if w > 0:
        n = f(r, "Instructions Executed")
        c[(opcode(r[ix["Source"]]), round(w / n, 2), round(f(r, "L1 Wavefronts Shared Ideal") / n, 2))] += 1
print("(opcode, wavefronts/instr, ideal/instr): static count")
for k, v in sorted(c.items(), key=lambda kv: -(kv[0][1] - kv[0][2]) * kv[1])[:14]:
    print(" ", k, v)
