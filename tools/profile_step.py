#!/usr/bin/env python
"""One small decode step for ncu captures: python tools/profile_step.py [slots] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ft8_demodulator_b200 import engine, synth  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd  # noqa: E402

slots = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = engine.require_cuda()
waves, _ = synth.synth_batch_torch(slots, dev, seed=7)
g = fd.decode_geometry(180000, 12000, 2, 2)
params = engine.make_params(12000, 2, 2, 20, 1, 20, g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"], 0)
dec = engine.SlotDecoder(slots, 180000, params, device=dev)
for _ in range(reps):
    dec.run(waves)
torch.cuda.synchronize()
print("decoded", int(dec.out.ok.sum()), "of", slots * 20, "candidates")
