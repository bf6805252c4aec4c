"""Randomised differential test of the stage-level drop-in functions on the GPU box (test infrastructure):
   python tools/fuzz_stages.py [trials] [seed]
Arbitrary host waterfalls (random shapes and oversampling factors, smooth / noisy / quantised / partly non-finite
values) go through the mirrors of the reference's functions - ft8_find_candidates, ft8_sync_score,
ft8_decode_candidate, ft8_extract_likelihood (+ ftx_normalize_logl), bp_decode - and every result is compared with the
CPU oracle on the same array: candidate lists and scores, LLRs, hard decisions, error counts, CRC fields and payloads
all EXACT (same spectrogram in, so no tolerance anywhere).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import ft8_oracle as orc  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator import ldpc_decoder as ld  # noqa: E402
from ft8_demodulator_b200.ft8_demodulator.ftx_types import FT8Candidate  # noqa: E402


def draw_mag(rng):
    tosr = int(rng.choice([1, 2, 2, 3, 4]))
    fosr = int(rng.choice([1, 2, 2, 3, 4, 5]))
    nb = int(rng.choice([60, 80, 93, 93, 100, 120]))             # symbols; below 80 no time offset is searched
    T = nb * tosr + int(rng.integers(0, tosr))
    F = 7 * fosr + int(rng.choice([1, 2, 5, 33, 64, 65, 127, 128, 129, 300]))
    kind = str(rng.choice(["noise", "noise", "smooth", "quantised", "nonfinite", "flat"]))
    if kind == "noise":
        mag = rng.standard_normal((F, T)) * 6.0 - 40.0
    elif kind == "smooth":
        mag = np.cumsum(rng.standard_normal((F, T)), axis=1) * 0.7 - 30.0
    elif kind == "quantised":
        mag = np.round(rng.standard_normal((F, T)) * 2.0)          # many exact ties
    elif kind == "flat":
        mag = np.full((F, T), -120.0)
        mag[rng.integers(0, F, 12), rng.integers(0, T, 12)] = rng.standard_normal(12) * 10
    else:
        mag = rng.standard_normal((F, T)) * 6.0 - 40.0
        bad = rng.choice([np.nan, np.inf, -np.inf], size=8)
        mag[rng.integers(0, F, 8), rng.integers(0, T, 8)] = bad
    return np.ascontiguousarray(mag, dtype=np.float64), tosr, fosr, kind


def run(trials, seed):
    rng = np.random.default_rng(seed)
    fails, n_lists, n_ties_raised, n_cands, n_scores = [], 0, 0, 0, 0
    t0 = time.time()
    for t in range(trials):
        mag, tosr, fosr, kind = draw_mag(rng)
        K = int(rng.choice([1, 3, 20, 64, 65, 300, 5000]))
        thr = float(rng.choice([-5.0, 0.0, 0.5, 2.0, 6.0]))
        desc = dict(trial=t, shape=list(mag.shape), tosr=tosr, fosr=fosr, kind=kind, K=K, thr=thr)
        wf = fd.create_waterfall_from_spectrogram(mag, tosr, fosr)
        with np.errstate(all="ignore"):
            try:
                ct, cf, cs, _ = orc.ft8_find_candidates(mag, tosr, fosr, K, thr)
                want = list(zip(ct.tolist(), cf.tolist(), cs.tolist()))
            except TypeError:
                want = None
                n_ties_raised += 1
        got = [(c.abs_time, c.abs_freq, float(c.score)) for c in fd.ft8_find_candidates(wf, K, thr)]
        if want is not None:
            n_lists += 1
            if got != want:
                k = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), min(len(got), len(want)))
                fails.append(dict(desc, why="find_candidates differs at %d of %d/%d: %r vs %r" % (
                    k, len(got), len(want), got[k] if k < len(got) else None, want[k] if k < len(want) else None)))
                continue
        # single scores and candidate decodes at random positions, in range or not
        t_lo, t_hi, _, f_hi = orc.search_ranges(mag.shape[0], mag.shape[1], tosr, fosr)
        for _ in range(6):
            at = int(rng.integers(t_lo - 30, max(t_lo, t_hi) + 30)) if t_hi > t_lo else int(rng.integers(-40, 200))
            af = int(rng.integers(0, max(1, f_hi)))
            cand = FT8Candidate(waterfall=wf, abs_time=at, abs_freq=af)
            n_scores += 1
            with np.errstate(all="ignore"):
                want_s = orc.ft8_sync_score(mag, tosr, fosr, at, af)
            got_s = fd.ft8_sync_score(wf, cand)
            if not (got_s == want_s or (np.isnan(got_s) and np.isnan(want_s))):
                fails.append(dict(desc, why="sync_score(%d,%d): %r vs %r" % (at, af, got_s, want_s)))
                continue
            if not (t_lo <= at < t_hi):
                continue
            n_cands += 1
            maxit = int(rng.choice([0, 1, 7, 20]))
            r = orc.ft8_decode_candidate(mag, tosr, fosr, at, af, maxit)
            l = np.zeros(174)
            fd.ft8_extract_likelihood(wf, cand, l)
            fd.ftx_normalize_logl(l)
            if not np.array_equal(l, r["llr"], equal_nan=True):
                fails.append(dict(desc, why="llr(%d,%d) differs" % (at, af)))
                continue
            ok, msg, st = fd.ft8_decode_candidate(wf, cand, maxit)
            got_r = (bool(ok), bytes(msg.payload), st.ldpc_errors, st.crc_extracted, st.crc_calculated)
            want_r = (r["ok"], r["payload"], r["ldpc_errors"], r["crc_extracted"], r["crc_calculated"])
            if got_r != want_r:
                fails.append(dict(desc, why="decode_candidate(%d,%d,%d): %r vs %r" % (at, af, maxit, got_r, want_r)))
            with np.errstate(all="ignore"):
                p_w, e_w = orc.bp_decode(r["llr"], maxit)
                p_g, e_g = ld.bp_decode(r["llr"].copy(), maxit)
            if e_w != e_g or not np.array_equal(p_w, p_g):
                fails.append(dict(desc, why="bp_decode(%d,%d,%d) differs" % (at, af, maxit)))
    return dict(tool="fuzz_stages", trials=trials, seed=seed, candidate_lists_compared=n_lists,
                lists_where_reference_raises=n_ties_raised, sync_scores_compared=n_scores,
                candidate_decodes_compared=n_cands, failures=fails[:20], n_failures=len(fails),
                seconds=round(time.time() - t0, 1))


def run_spectrogram(trials, seed):
    """calculate_spectrogram (mirror of spectrogram_analyse.py:19-66) on random geometries, lengths and input types against
    a NumPy restatement of the scipy call it replaces (Hann, zero-padded FFT, |X|^2 / (sum w)^2, 10*log10(1e-12 + .),
    fftshift): <= 1e-8 dB for float64 / complex128 input, 1e-4 for float32 / complex64 (the output is float32 there)."""
    from ft8_demodulator_b200.ft8_demodulator import spectrogram_analyse as sa
    rng = np.random.default_rng(seed)
    fails, worst = [], 0.0
    t0 = time.time()
    for t in range(trials):
        fs = rng.choice([2000, 3500, 4500, 6000, 7350, 8000, 9999.5, 10e3, 11025, 12000, 16000, 20000, 22050, 32768, 44100, 48000, 96000, 192000, 250000])
        fs = float(fs) if fs != int(fs) else int(fs)
        b, s = int(rng.integers(1, 6)), int(rng.integers(1, 9))
        nperseg = int(0.16 * fs)
        hop = nperseg - (nperseg - nperseg // s)
        nfft = int(fs / 6.25 * b)
        if nfft > 400000:
            continue
        n = int(rng.choice([nperseg - 1, nperseg, nperseg + 1, nperseg + hop - 1, nperseg + hop, int(nperseg * rng.uniform(1, 6))]))
        kind = str(rng.choice(["f64", "f32", "c128", "c64", "i16"]))
        x = rng.standard_normal(n) * 0.2 + np.cos(2 * np.pi * 0.07 * np.arange(n))
        if kind in ("c128", "c64"):
            x = x + 1j * rng.standard_normal(n) * 0.2
        x = x.astype({"f64": np.float64, "f32": np.float32, "c128": np.complex128, "c64": np.complex64, "i16": np.float64}[kind])
        if kind == "i16":
            x = np.round(x * 1000).astype(np.int16)
        desc = dict(trial=t, fs=fs, b=b, s=s, n=n, kind=kind, nperseg=nperseg, hop=hop, nfft=nfft)
        try:
            S, f, tt = sa.calculate_spectrogram(x, fs, b, s)
        except NotImplementedError as e:
            desc["unsupported"] = str(e)[:60]
            continue
        if n < nperseg:
            if S.shape != (1, 0) or len(f) or len(tt):
                fails.append(dict(desc, why="short input must give (array([[]]), [], [])"))
            continue
        win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nperseg) / nperseg)
        T = (n - nperseg) // hop + 1
        xw = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
        frames = np.stack([xw[j * hop:j * hop + nperseg] * win for j in range(T)])
        ref = np.fft.fftshift(10 * np.log10(1e-12 + np.abs(np.fft.fft(frames, n=nfft, axis=1)) ** 2 / win.sum() ** 2).T, axes=0)
        tol = 1e-4 if kind in ("f32", "c64") else 1e-8
        want_dtype = np.float32 if kind in ("f32", "c64") else np.float64
        if S.shape != ref.shape or S.dtype != want_dtype:
            fails.append(dict(desc, why="shape/dtype %r %r vs %r" % (S.shape, S.dtype, ref.shape)))
            continue
        err = float(np.abs(S - ref).max())
        worst = max(worst, err if tol == 1e-8 else 0.0)
        if not err < tol:
            fails.append(dict(desc, why="max |dB error| %g" % err))
        if not (np.array_equal(f, np.fft.fftshift(np.fft.fftfreq(nfft, 1 / fs))) and len(tt) == T):
            fails.append(dict(desc, why="axes differ"))
    return dict(tool="fuzz_stages.spectrogram", trials=trials, seed=seed, worst_float64_db_error=worst, failures=fails[:20],
                n_failures=len(fails), seconds=round(time.time() - t0, 1))


def main():
    out = run(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(json.dumps(out))
    out2 = run_spectrogram(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print(json.dumps(out2))
    sys.exit(1 if out["n_failures"] or out2["n_failures"] else 0)


if __name__ == "__main__":
    main()
