#!/usr/bin/env python
"""decode_slots from a PAGEABLE NumPy array (what a NumPy-only caller of the reference passes) against the number of
staging threads and the store kind of ft8b_decode_slots_host's pinned-staging gather:
    python tools/stage_sweep.py [n_slots] [reps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ft8_demodulator_b200 import _lib, engine, synth  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = engine.require_cuda()
waves, _ = synth.synth_batch_torch(n, dev, seed=0)
host_np = waves.cpu().numpy()                 # pageable
pinned = waves.cpu().pin_memory()
kw = dict(max_candidates=20, min_score=1, max_iterations=20)
L = _lib.lib()


def run(x, piece):
    best = 1e9
    for _ in range(reps + 1):
        t = time.perf_counter()
        fd.decode_slots(x, 12000, piece_slots=piece, **kw)
        best = min(best, time.perf_counter() - t)
    return round(n / best)


out = {"n_slots": n, "host_threads": os.cpu_count(), "pinned": {p: run(pinned, p) for p in (128, 256)}, "pageable": {}}
for thr in (1, 2, 4, 6, 8, 10, 12, 14, 16):
    for nt in (1, 0):
        eff = L.ft8b_set_host_staging(thr, nt)
        out["pageable"]["t%d_nt%d" % (eff, nt)] = {p: run(host_np, p) for p in (128, 256)}
print(json.dumps(out))
