#!/usr/bin/env python
"""Host-to-device bandwidth of the box (pinned memory, one copy engine), the bound of bench.py's e2e:
   python tools/h2d_bw.py"""
import time
import torch

dev = torch.device("cuda", 0)
for mb in (92, 737, 2949):
    n = mb * (1 << 20) // 4
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h.normal_()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("H2D %5d MiB: %.2f ms  %.1f GB/s" % (mb, ms, n * 4 / ms / 1e6))
    # device -> host for completeness
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("D2H %5d MiB: %.2f ms  %.1f GB/s" % (mb, ms, n * 4 / ms / 1e6))
