#!/usr/bin/env python
"""Small pass over every kernel for compute-sanitizer (memcheck):  compute-sanitizer python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ft8_demodulator_b200 import engine, synth  # noqa: E402
from ft8_demodulator_b200.ft8_beacon_receiver import frequency_correction as fc  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd, ldpc_decoder as ld, spectrogram_analyse as sa  # noqa: E402

rng = np.random.default_rng(0)
w, pl = synth.synth_batch_torch(3, engine.require_cuda(), seed=1)
for kw in (dict(), dict(freq_min=200.0, freq_max=2000.0, time_min=0.3, time_max=14.0), dict(max_candidates=300, min_score=2)):
    b = fd.decode_slots(w, 12000, 2, 2, **{**dict(max_candidates=20, min_score=1, max_iterations=20), **kw})
    print("decode", kw, int(b.arrays.ok.sum()))
x16 = (w[:2].cpu().numpy() * 3000).astype(np.int16)
print("pcm16", int(fd.decode_slots(x16, 12000, pcm16=True, min_score=1).arrays.ok.sum()))
for fs, b_, s_ in ((6000, 2, 2), (3500, 2, 2), (10000, 4, 4), (12000, 3, 5), (32768, 2, 2), (12000, 1, 1)):
    x = rng.standard_normal(int(fs * 9.3))
    S, f, t = sa.calculate_spectrogram(x, fs, b_, s_)
    r = fd.decode_ft8_message(x, fs, b_, s_, 10, 0.5, 3)
    print("geom", fs, b_, s_, S.shape, len(r))
S, _, _ = sa.calculate_spectrogram(rng.standard_normal(30001) + 1j * rng.standard_normal(30001), 12000, 2, 2)
llr, _ = synth.c5_llr(300, seed=2)
print("bp", ld.bp_decode_batch(llr, 20)[1][:5], ld.ldpc_check(np.zeros(174)))
xc = synth.c4_capture(seed=5, fs=12000, drift_hz_s=3.3, esn0_db=30.0)
print("drift", fc.correct_frequency_drift(xc, 12000, 6.25, 0.16, {'steps_per_symbol': 8, 'debug_plots': False})[1] * 12000)
from ft8_demodulator_b200 import _lib  # noqa: E402
mag = engine.stft_db(w[:2], 1920, 960, 3840, 0, 1920, 0, 186)
for variant, hyps in ((_lib.DRIFT_WALK, [0.0, 0.17, -0.17, 0.05]), (_lib.DRIFT_BLOCKSUM, [0.0, 0.17, -0.17, 0.05]),
                      (_lib.DRIFT_DIRECT, [0.0, 1.5]), (_lib.DRIFT_AUTO, [0.0] + list(np.linspace(-0.2, 0.2, 127)))):
    sc, hy, _ = engine.sync_scores_drift(mag, 2, 2, hyps, variant=variant)
    print("drift sync variant", variant, float(sc[torch.isfinite(sc)].max()), int(hy.max()))
ds = np.stack([synth.drifting_slot(bytes([7] * 9 + [0x50]), seed=3, drift_hz_s=3.3, f0=900.0)])
found, dr = fd.decode_slots_drift(ds, 12000, np.arange(-4.0, 4.01, 0.5), max_candidates=20, min_score=1)
print("drift decode", int(found.arrays.ok.sum()), float(np.abs(np.asarray(dr)).max()))
host = fd.decode_slots(np.ascontiguousarray(w[:3].cpu().numpy()), 12000, 2, 2, max_candidates=20, min_score=1)   # ft8b_decode_slots_host, pageable
print("host abi", int(host.arrays.ok.sum()))
torch.cuda.synchronize()
print("sanitize smoke done")
