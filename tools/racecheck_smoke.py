#!/usr/bin/env python
"""One slot through every shared-memory kernel for compute-sanitizer --tool racecheck (slow tool, tiny input):
    compute-sanitizer --tool racecheck python tools/racecheck_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ft8_demodulator_b200 import _lib, engine, synth  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd, ldpc_decoder as ld  # noqa: E402

w, _ = synth.synth_batch_torch(1, engine.require_cuda(), seed=4)
b = fd.decode_slots(w, 12000, 2, 2, max_candidates=20, min_score=1, max_iterations=20)          # stft1920, sync pair, select, decode
print("decode", int(b.arrays.ok.sum()))
b = fd.decode_slots(w[:, :60000].contiguous(), 6000, 3, 1, max_candidates=40, min_score=1, max_iterations=5)   # generic stft, generic sync
print("decode generic", int(b.arrays.n_cand[0]))
mag = engine.stft_db(w, 1920, 960, 3840, 0, 1920, 0, 186)
for variant in (_lib.DRIFT_WALK, _lib.DRIFT_BLOCKSUM):
    sc, hy, _ = engine.sync_scores_drift(mag, 2, 2, [0.0, 0.17, -0.17, 0.05], variant=variant)
    print("drift sync variant", variant, float(sc[torch.isfinite(sc)].max()))
llr, _ = synth.c5_llr(64, seed=2)
print("bp", ld.bp_decode_batch(llr, 20)[1][:5])
torch.cuda.synchronize()
print("racecheck smoke done")
