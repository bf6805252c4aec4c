#!/usr/bin/env python
"""Stall-reason samples per kernel phase (segments between BAR.SYNC) from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # one section per profiled launch (ncu repeats each twice)
hi = his[sec]
end = his[sec + 1] if sec + 1 < len(his) else len(rows)
data = [r for r in rows[hi + 1:end] if len(r) >= len(hdr)]


def num(r, k):
    v = r[ix[k]]
    return float(v) if v else 0.0


tot = collections.Counter()
for r in data:
    for h in stall:
        tot[h[6:]] += int(num(r, h))
T = sum(tot.values())
print("all:", {k: round(100 * v / T, 1) for k, v in tot.most_common(9)})
bars = [i for i, r in enumerate(data) if "BAR" in r[ix["Source"]]]
prev = 0
for b in bars + [len(data) - 1]:
    seg = data[prev:b + 1]
    st = collections.Counter()
    for r in seg:
        for h in stall:
            st[h[6:]] += int(num(r, h))
    n = sum(int(num(r, "# Samples")) for r in seg)
    ie = sum(num(r, "Instructions Executed") for r in seg)
    if n:
        print("sass %4d..%4d  samples %6d (%4.1f%%)  instr %6.1fM  %s" % (prev, b, n, 100 * n / max(T, 1), ie / 1e6,
                                                                      dict((k, v) for k, v in st.most_common(5))))
    prev = b + 1
