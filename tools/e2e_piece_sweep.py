#!/usr/bin/env python
"""End-to-end (pinned host -> results on host) throughput of decode_slots against the piece size of the H2D/compute
overlap: python tools/e2e_piece_sweep.py [slots]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ft8_demodulator_b200 import engine, synth  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd  # noqa: E402

slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = engine.require_cuda()
waves, _ = synth.synth_batch_torch(slots, dev, seed=3)
host = waves.cpu().pin_memory()
host16 = (waves * 32767).round().clamp(-32768, 32767).to(torch.int16).cpu().pin_memory()
kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
for name, h, extra in (("float32", host, {}), ("pcm16", host16, {"pcm16": True})):
    for piece in (64, 128, 256, 512, 1024):
        fd.decode_slots(h, 12000, piece_slots=piece, **kw, **extra)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            fd.decode_slots(h, 12000, piece_slots=piece, **kw, **extra)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 3
        print("%-8s piece %5d: %8.0f slots/s  (%.2f ms, H2D %.1f GB/s)" % (name, piece, slots / dt, 1e3 * dt,
                                                                          h.numel() * h.element_size() / dt / 1e9))
