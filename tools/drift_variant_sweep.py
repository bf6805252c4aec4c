#!/usr/bin/env python
"""Walk vs block-sum drift-hypothesis sync kernel against the number of hypotheses (+-3.33 Hz/s span):
    python tools/drift_variant_sweep.py [slots]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ft8_demodulator_b200 import _lib, engine  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(1)
mag = (torch.randn((m, 186, 1920), generator=g, device="cuda", dtype=torch.float64) * 12.0 - 60.0).contiguous()


def t(hyps, variant):
    try:
        engine.sync_scores_drift(mag, 2, 2, hyps, variant=variant)
    except ValueError:
        return None
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        engine.sync_scores_drift(mag, 2, 2, hyps, variant=variant)
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / 3 / m * 1e3, 2)


out = {}
for span in (3.33, 10.0):
    for n in (4, 8, 16, 28, 56, 96, 127):
        hyps = [0.0] + [float(r) * 0.16 / 3.125 for r in np.linspace(-span, span, n)]
        out["span%g_n%d" % (span, len(hyps))] = {"walk_us": t(hyps, _lib.DRIFT_WALK), "blocksum_us": t(hyps, _lib.DRIFT_BLOCKSUM),
                                                 "auto_us": t(hyps, _lib.DRIFT_AUTO)}
print(json.dumps(out))
