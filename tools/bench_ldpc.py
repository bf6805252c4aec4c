#!/usr/bin/env python
"""Config C5 of BASELINE.json: LDPC codewords/s on low-SNR LLR vectors (standalone ft8b_ldpc_crc).

    python tools/bench_ldpc.py [n_words] [steps]        (default 1e6 words, 5 steps)
Prints one JSON line: codewords/s on the GPU (CUDA events, inputs resident), the CPU oracle port on all host
threads on a bounded sample, and bit-exact parity of that sample.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from ft8_demodulator_b200 import engine, synth  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = engine.require_cuda()
base, bits = synth.c5_llr(min(n, 100_000), seed=0)            # SNR_2500 in [-26,-20] dB
reps = (n + len(base) - 1) // len(base)
llr = torch.from_numpy(base).to(dev).repeat(reps, 1)[:n].contiguous()
# decorrelate the repeats: permuting the 58 data symbols of a vector keeps it a valid LLR vector of another word only
# approximately, so instead scale each repeat slightly (does not change hard decisions' statistics much)
scale = 1.0 + 0.01 * torch.arange(reps, device=dev, dtype=torch.float64).repeat_interleave(len(base))[:n]
llr *= scale[:, None]
for _ in range(3):
    out = engine.ldpc_crc(llr, 20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = engine.ldpc_crc(llr, 20)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
import ft8_oracle as orc  # noqa: E402
threads = len(os.sched_getaffinity(0))
m = min(n, 4096 * 4)
t0 = time.perf_counter()
plain, err = orc.bp_batch(llr[:m].cpu().numpy(), 20, n_threads=threads)
cpu_s = time.perf_counter() - t0
same = bool(np.array_equal(out.plain[:m].cpu().numpy(), plain) and np.array_equal(out.ldpc_errors[:m].cpu().numpy(), err))
print(json.dumps({"metric": "LDPC codewords/sec (C5, max_iterations=20)", "n_words": n, "ms_per_step": ms,
                  "value": n / (ms * 1e-3), "unit": "codewords/s", "converged_frac": float((out.ldpc_errors == 0).float().mean()),
                  "crc_ok_frac": float(out.ok.float().mean()),
                  "cpu_port": {"value": m / cpu_s, "unit": "codewords/s", "cores": threads, "sample": m},
                  "sample_bit_exact_vs_oracle": same}))
