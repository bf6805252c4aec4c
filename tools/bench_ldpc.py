#!/usr/bin/env python
"""Config C5 of BASELINE.json: LDPC codewords/s on low-SNR LLR vectors (standalone ft8b_ldpc_crc).

    python tools/bench_ldpc.py [n_words] [steps]        (default 1e6 words, 5 steps)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ldpc.py 1e6 5
Prints one JSON line: codewords/s on the GPU (CUDA events, inputs resident), the CPU oracle port on all host
threads on a bounded sample, and bit-exact parity of that sample.  Under torchrun the n_words are sharded across
the ranks (dist.shard_range: strong scaling, no collective on the data path); the time is the max over ranks.
"""
import json
import os
import sys
import time

import numpy as np

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":     # keep stdout to the one JSON line
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import torch.distributed as dist  # noqa: E402

from ft8_demodulator_b200 import dist as fdist, engine, synth  # noqa: E402

n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, local_rank, world = fdist.init_process_group()
torch.cuda.set_device(local_rank)
dev = engine.require_cuda()
lo, hi = fdist.shard_range(n_total, rank, world)
n = hi - lo                                                   # this rank's share of the words
base, bits = synth.c5_llr(min(n, 100_000), seed=0)            # SNR_2500 in [-26,-20] dB
reps = (n + len(base) - 1) // len(base)
llr = torch.from_numpy(base).to(dev).repeat(reps, 1)[:n].contiguous()
# decorrelate the repeats: permuting the 58 data symbols of a vector keeps it a valid LLR vector of another word only
# approximately, so instead scale each repeat slightly (does not change hard decisions' statistics much)
scale = 1.0 + 0.01 * (rank * reps + torch.arange(reps, device=dev, dtype=torch.float64)).repeat_interleave(len(base))[:n]
llr *= scale[:, None]
for _ in range(3):
    out = engine.ldpc_crc(llr, 20)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = engine.ldpc_crc(llr, 20)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
stats = torch.tensor([ms, float((out.ldpc_errors == 0).sum()), float(out.ok.sum())], dtype=torch.float64, device=dev)
if world > 1:
    t = stats[:1].clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = stats[1:].clone()
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    stats = torch.cat([t, c])
ms, n_conv, n_ok = (float(v) for v in stats)
if rank != 0:
    dist.destroy_process_group()
    sys.exit(0)
import ft8_oracle as orc  # noqa: E402
threads = len(os.sched_getaffinity(0))
m = min(n, 4096 * 4)
t0 = time.perf_counter()
plain, err = orc.bp_batch(llr[:m].cpu().numpy(), 20, n_threads=threads)
cpu_s = time.perf_counter() - t0
same = bool(np.array_equal(out.plain[:m].cpu().numpy(), plain) and np.array_equal(out.ldpc_errors[:m].cpu().numpy(), err))
print(json.dumps({"metric": "LDPC codewords/sec (C5, max_iterations=20)", "n_words": n_total, "n_gpus": world,
                  "scaling": "strong", "ms_per_step": ms,
                  "value": n_total / (ms * 1e-3), "unit": "codewords/s", "converged_frac": n_conv / n_total,
                  "crc_ok_frac": n_ok / n_total,
                  "cpu_port": {"value": m / cpu_s, "unit": "codewords/s", "cores": threads, "sample": m},
                  "sample_bit_exact_vs_oracle": same}))
if world > 1:
    dist.destroy_process_group()
