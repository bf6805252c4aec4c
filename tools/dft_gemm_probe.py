"""Evidence for DESIGN.md section 4 ("no tensor cores on this path"), measured rather than argued (GPU box):
   python tools/dft_gemm_probe.py [slots]
The STFT as ONE dense contraction - frames [slots*186, 1920] x (window folded into the DFT matrix) [1920, 2*1920]
(real and imaginary parts of bins 0..1919), 2.74 GFLOP per slot - run by cuBLAS through torch.matmul in bf16, tf32,
fp32 and fp64, against the float64 FFT of the same frames: time per slot (CUDA events, best of 5) and the largest dB
error over bins that carry signal.  This is library code used as a yardstick for the rejected variant, not a product
path; the shipped kernel is csrc/stft1920.cuh (1.86 us per slot, <= 1e-11 dB)."""
import json
import sys

import numpy as np
import torch


def main():
    slots = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    n, nper, hop, nfft, T = 180000, 1920, 960, 3840, 186
    t = torch.arange(n, device=dev, dtype=torch.float64) / 12000.0
    x = torch.randn(slots, n, device=dev, dtype=torch.float64, generator=g) + 0.05 * torch.sin(2 * np.pi * 1000.0 * t)
    frames = x.unfold(1, nper, hop).reshape(slots * T, nper).contiguous()              # [M, 1920]
    win = 0.5 - 0.5 * torch.cos(2 * np.pi * torch.arange(nper, device=dev, dtype=torch.float64) / nper)
    k = torch.arange(1920, device=dev, dtype=torch.float64)
    ang = -2 * np.pi * torch.outer(torch.arange(nper, device=dev, dtype=torch.float64), k) / nfft
    W = torch.cat([torch.cos(ang), torch.sin(ang)], dim=1) * win[:, None]               # [1920, 3840]
    scale = float(win.sum()) ** 2

    def db(re, im):
        return 10 * torch.log10(1e-12 + (re.double() ** 2 + im.double() ** 2) / scale)

    ref = torch.fft.rfft(frames * win, n=nfft, dim=1)[:, :1920]
    ref_db = db(ref.real, ref.imag)
    out = {"tool": "dft_gemm_probe", "slots": slots, "gemm": "[%d x 1920] x [1920 x 3840]" % (slots * T),
           "gflop_per_slot": 2 * T * nper * 3840 / 1e9, "variants": {}}
    for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False),
                           ("fp64", torch.float64, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a, b = frames.to(dt), W.to(dt)
        y = a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        got = db(y[:, :1920], y[:, 1920:])
        out["variants"][name] = {"us_per_slot": round(best * 1e3 / slots, 3), "tflops": round(out["gflop_per_slot"] * slots / best, 1),
                                 "max_db_error": float((got - ref_db).abs().max()),
                                 "median_db_error": float((got - ref_db).abs().median())}
    torch.backends.cuda.matmul.allow_tf32 = False
    out["shipped_kernel"] = {"us_per_slot": 1.86, "max_db_error_vs_pocketfft": 5e-12, "source": "profiles/r2_bench_default.json, tests/test_gpu_parity.py"}
    out["parity_bar_db"] = 1e-9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
