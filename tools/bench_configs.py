#!/usr/bin/env python
"""Latency / throughput of the BASELINE.json configs that bench.py does not time (bench.py is C3; C5: tools/bench_ldpc.py;
C4 extension: tools/bench_drift.py).  One JSON line.

  C1  one 15-s slot, one signal at -10 dB, decode_ft8_message(float64 NumPy array) - host array in, Python list out,
      through ft8b_decode_slots_host; median wall-clock latency (the reference: 15.6 s per slot)
  C2  one dense-band slot (50 signals, SNR -24..0 dB) with K = 2000 / min_score 3 and with K = 20: latency of the public
      call, and slots/s of a device-resident batch of 64 such slots (each decodes up to 2000 candidates)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ft8_demodulator_b200 import engine, synth  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd  # noqa: E402

dev = engine.require_cuda()


def latency(fn, n=20):
    fn()
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e3, float(np.min(ts)) * 1e3, r


c1 = synth.c1_slot()
kw1 = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
c1_med, c1_min, r1 = latency(lambda: fd.decode_ft8_message(c1, 12000, **kw1))
# where the C1 latency goes: device-side span of the host-ABI call (first copy enqueued .. results on the host, CUDA
# events inside the library) against the wall clock of the Python call around it
spans = []
for _ in range(20):
    fd.decode_ft8_message(c1, 12000, **kw1)
    spans.append(engine._lib.lib().ft8b_last_host_call_ms())
c1_span = float(np.median(spans))

c2, payloads, *_ = synth.c2_slot()
kw2 = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=2000, min_score=3, max_iterations=20, freq_max=3100.0)
c2_med, c2_min, r2 = latency(lambda: fd.decode_ft8_message(c2, 12000, **kw2), 10)
kw2b = dict(kw2, max_candidates=20)
c2b_med, c2b_min, r2b = latency(lambda: fd.decode_ft8_message(c2, 12000, **kw2b), 10)
found = {bytes(m.payload) for m, *_ in r2}
# opt-in strict top-K (not the reference's list): K = 20 and K = 200 best sync scores of the dense band
c2s_med, c2s_min, r2s = latency(lambda: fd.decode_ft8_message(c2, 12000, strict_topk=True, **kw2b), 10)
c2t_med, c2t_min, r2t = latency(lambda: fd.decode_ft8_message(c2, 12000, strict_topk=True, **dict(kw2, max_candidates=200)), 10)
found_s = {bytes(m.payload) for m, *_ in r2s}
found_t = {bytes(m.payload) for m, *_ in r2t}

nb = 64
batch = torch.from_numpy(c2.astype(np.float32)).to(dev).repeat(nb, 1).contiguous()
out = fd.decode_slots(batch, 12000, return_device=True, **kw2)
torch.cuda.synchronize()
steps_ms = []
for _ in range(5):                  # every step timed on its own; the median is immune to a one-off allocator stall
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fd.decode_slots(batch, 12000, return_device=True, **kw2)
    e1.record()
    torch.cuda.synchronize()
    steps_ms.append(e0.elapsed_time(e1))
ms = float(np.median(steps_ms))
n_cand = int(out.n_cand.sum())
print(json.dumps({
    "C1_single_slot": {"latency_ms_median": c1_med, "latency_ms_min": c1_min, "decodes": len(r1), "device_span_ms_median": c1_span,
                       "payload": r1[0][0].payload.hex() if r1 else None, "reference_python_s_per_slot": 15.6},
    "C2_dense_band_K2000": {"latency_ms_median": c2_med, "latency_ms_min": c2_min, "decodes": len(r2),
                            "distinct_payloads": len(found), "transmitted_found": sum(p in found for p in payloads),
                            "transmitted": len(payloads)},
    "C2_dense_band_K20": {"latency_ms_median": c2b_med, "latency_ms_min": c2b_min, "decodes": len(r2b)},
    "C2_dense_band_K20_strict_topk": {"latency_ms_median": c2s_med, "latency_ms_min": c2s_min, "decodes": len(r2s),
                                      "transmitted_found": sum(p in found_s for p in payloads)},
    "C2_dense_band_K200_strict_topk": {"latency_ms_median": c2t_med, "latency_ms_min": c2t_min, "decodes": len(r2t),
                                       "transmitted_found": sum(p in found_t for p in payloads)},
    "C2_batch64_K2000_device_resident": {"ms_per_step": ms, "ms_per_step_all": steps_ms, "slots_per_s": nb / (ms * 1e-3), "candidates_decoded_per_step": n_cand,
                                         "ldpc_codewords_per_s": n_cand / (ms * 1e-3)}}))
