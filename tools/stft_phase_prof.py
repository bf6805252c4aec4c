#!/usr/bin/env python
"""Warp-cycles per phase of stft1920_kernel (experiment build: tools/build_variant.sh prof -DFT8B_STFT_PROF):
   FT8B_LIB_PATH=variants/libprof.so python tools/stft_phase_prof.py [slots]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ft8_demodulator_b200 import _lib, engine, synth  # noqa: E402

slots = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = engine.require_cuda()
waves, _ = synth.synth_batch_torch(slots, dev, seed=7)
mag = torch.empty((slots, 186, 1920), dtype=torch.float64, device=dev)
L = C.CDLL(_lib.LIB_PATH)
run = lambda: _lib.check(_lib.lib().ft8b_stft_db(waves.data_ptr(), _lib.DTYPE_F32, slots, 180000, 180000, 1920, 960, 3840, 0, 1920,
                                                  0, 186, mag.data_ptr(), None), "stft")
run(); run()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
L.ft8b_debug_stft_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
L.ft8b_debug_stft_prof(buf, 0)
names = ["wait audio", "P1 load+bf15+store+stage", "barriers P1", "P2 load+tw+bf16+store", "barriers P2", "P3 load+tw+2xbf8",
         "untangle+dB+store", "-"]
tot = sum(buf[:8])
frames = slots * 186
print("ms", e0.elapsed_time(e1), "frames", frames)
for n, v in zip(names, buf[:8]):
    print("%-28s %6.1f %%   %8.0f warp-cycles per frame-warp" % (n, 100.0 * v / tot, v / (frames * 4.0)))
print("total warp-cycles per frame-warp %.0f" % (tot / (frames * 4.0)))
