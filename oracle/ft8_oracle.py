"""ctypes front end of the CPU parity oracle (oracle/ft8_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs - never by the product package.
Function names mirror the reference functions they restate (see ft8_oracle.c for file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libft8oracle.so")

ERR_TIE = -2
LDPC_N = 174


class Record(C.Structure):
    _fields_ = [
        ("abs_time", C.c_int32), ("abs_freq", C.c_int32), ("score", C.c_double),
        ("ok", C.c_int32), ("ldpc_errors", C.c_int32), ("crc_extracted", C.c_int32),
        ("crc_calculated", C.c_int32), ("payload", C.c_uint8 * 10), ("plain", C.c_uint8 * LDPC_N),
        ("_pad", C.c_uint8 * 8), ("llr", C.c_double * LDPC_N),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "ft8_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_plan.argtypes = [C.c_double, C.c_int, C.c_int, ip, ip, ip]
        L.orc_num_frames.restype = C.c_long
        L.orc_num_frames.argtypes = [C.c_long, C.c_int, C.c_int]
        L.orc_spectrogram_pos.argtypes = [dp, C.c_long, C.c_int, C.c_int, C.c_int, dp]
        L.orc_sync_score.restype = C.c_double
        L.orc_sync_score.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_search_ranges.argtypes = [C.c_int] * 4 + [ip] * 4
        L.orc_score_map.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_find_candidates.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                          ip, ip, dp, C.POINTER(C.c_long)]
        L.orc_extract_likelihood.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_normalize_logl.argtypes = [dp]
        L.orc_ldpc_check.argtypes = [C.POINTER(C.c_uint8)]
        L.orc_bp_decode.argtypes = [dp, C.c_int, C.POINTER(C.c_uint8)]
        L.orc_compute_crc.restype = C.c_uint
        L.orc_compute_crc.argtypes = [C.POINTER(C.c_uint8), C.c_int]
        L.orc_extract_crc.restype = C.c_uint
        L.orc_extract_crc.argtypes = [C.POINTER(C.c_uint8)]
        L.orc_pack_bits.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8)]
        L.orc_decode_candidate.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(Record)]
        L.orc_masks.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_long] + [C.c_double] * 4 + \
                               [ip, ip, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_decode_message.argtypes = [dp, C.c_long, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(Record), C.c_int,
                                         ip, ip]
        L.orc_decode_message_c.argtypes = L.orc_decode_message.argtypes
        L.orc_spectrogram_pos_c.argtypes = L.orc_spectrogram_pos.argtypes
        L.orc_decode_batch.argtypes = [dp, C.c_long, C.c_long, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                       C.c_int, C.POINTER(Record), ip, C.c_int]
        L.orc_bp_batch.argtypes = [dp, C.c_long, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.c_int]
        assert L.orc_record_size() == C.sizeof(Record), (L.orc_record_size(), C.sizeof(Record))
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _none_nan(v):
    return float("nan") if v is None else float(v)


def plan(sample_rate, bins_per_tone=2, steps_per_symbol=2):
    """(nperseg, hop, nfft) of spectrogram_analyse.py:32-43."""
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().orc_plan(float(sample_rate), bins_per_tone, steps_per_symbol, C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def spectrogram_pos(wave, sample_rate, bins_per_tone=2, steps_per_symbol=2):
    """dB spectrogram rows f >= 0, shape [ceil(nfft/2), T] (what ft8_decode.py:323-325 keeps); real or complex audio."""
    cplx = np.iscomplexobj(wave)
    x = np.ascontiguousarray(np.asarray(wave, dtype=np.complex128).reshape(-1)).view(np.float64) if cplx else _f64(wave)
    n = len(x) // 2 if cplx else len(x)
    nperseg, hop, nfft = plan(sample_rate, bins_per_tone, steps_per_symbol)
    T = lib().orc_num_frames(n, nperseg, hop)
    if T <= 0:
        return np.zeros((0, 0))
    mag = np.empty(((nfft + 1) // 2, T), dtype=np.float64)
    fn = lib().orc_spectrogram_pos_c if cplx else lib().orc_spectrogram_pos
    rc = fn(_dp(x), n, nperseg, hop, nfft, _dp(mag))
    if rc:
        raise ValueError("orc_spectrogram_pos failed: %d" % rc)
    return mag


def search_ranges(F, T, tosr, fosr):
    v = [C.c_int() for _ in range(4)]
    lib().orc_search_ranges(F, T, tosr, fosr, *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def ft8_sync_score(mag, tosr, fosr, abs_time, abs_freq):
    m = _f64(mag)
    return lib().orc_sync_score(_dp(m), m.shape[0], m.shape[1], tosr, fosr, abs_time, abs_freq)


def score_map(mag, tosr, fosr):
    m = _f64(mag)
    t0, t1, f0, f1 = search_ranges(m.shape[0], m.shape[1], tosr, fosr)
    out = np.empty((max(t1 - t0, 0), max(f1 - f0, 0)), dtype=np.float64)
    if out.size:
        lib().orc_score_map(_dp(m), m.shape[0], m.shape[1], tosr, fosr, _dp(out))
    return out


def score_map_laplacian(mag, tosr, fosr):
    """The reference's score map (ft8_sync_score over ft8_find_candidates' ranges, ft8_decode.py:47-124) restated as a
    Laplacian of Costas PATH SUMS - the formulation csrc/sync_lap.cuh computes (same terms, different association: it
    agrees with score_map(), the reference-pinned running sum, to ~1e-14, not bit for bit; tests/test_oracle_golden.py).

    Per oversampling phase, with A[g][bin] the phase's columns and columns outside [0, nb) contributing nothing:
        P[h](f)  = sum_k A[h+k][f + b*C[k]]                       plain sum of the 7 cells of block h
        B[h](f)  = 4 P[h](f) - P[h](f+b) - P[h](f-b) - P[h-1](f) - P[h+1](f) + Corr[h](f)
        Corr[h]  = A[h+3][f-b] - A[h+3][f]                        tone 0 (k = 3) has no lower neighbour
                 + A[h-1][c0]  - A[h][c0]                         k = 0 has no earlier neighbour
                 + A[h+7][c6]  - A[h+6][c6]                       k = 6 has no later neighbour
                 + A[0][c(j-1)] - A[0][c(j)]        (h = -j, j = 1..6)       block cut by the first column
                 + A[nb-1][c(k+1)] - A[nb-1][c(k)]  (h = nb-1-k, k = 0..5)   block cut by the last column
        score(u, f) = (B[u-10] + B[u+26] + B[u+62]) / num_comparisons(u)
    mag [F, T] as elsewhere in this module.  NumPy, vectorised over frequency."""
    m = _f64(mag)
    F, T = m.shape
    s, b = tosr, fosr
    nb = T // s
    t0, t1, f0, f1 = search_ranges(F, T, s, b)
    nt, nf = max(t1 - t0, 0), max(f1 - f0, 0)
    out = np.full((nt, nf), -np.inf)
    if nt == 0 or nf == 0:
        return out
    Cp = [3, 1, 4, 0, 6, 5, 2]
    nu = nt // s
    pad = 2 * b
    f = np.arange(nf)
    fr = np.arange(-b, nf + b)                              # candidates including one tone of halo on both sides
    for rho in range(s):
        A = m[:, rho::s][:, :nb].T
        Ap = np.zeros((nb, F + 2 * pad))
        Ap[:, pad:pad + F] = A
        Ap[:, :pad] = A[:, :1]                              # bins outside the row: any finite value (they cancel)
        Ap[:, pad + F:] = A[:, -1:]

        def a(g, bins):
            return Ap[g, bins + pad] if 0 <= g < nb else np.zeros(len(bins))

        P = {h: sum(a(h + k, fr + b * Cp[k]) for k in range(7)) for h in range(-8, nb + 2)}
        B = {}
        for h in range(-7, nb):
            corr = a(h + 3, f - b) - a(h + 3, f)
            corr = corr + a(h - 1, f + b * Cp[0]) - a(h, f + b * Cp[0])
            corr = corr + a(h + 7, f + b * Cp[6]) - a(h + 6, f + b * Cp[6])
            if -6 <= h <= -1:
                corr = corr + a(0, f + b * Cp[-h - 1]) - a(0, f + b * Cp[-h])
            if nb - 6 <= h <= nb - 1:
                k = nb - 1 - h
                corr = corr + a(nb - 1, f + b * Cp[k + 1]) - a(nb - 1, f + b * Cp[k])
            ctr, up, dn = slice(b, b + nf), slice(2 * b, 2 * b + nf), slice(0, nf)
            B[h] = 4 * P[h][ctr] - P[h][up] - P[h][dn] - P[h - 1][ctr] - P[h + 1][ctr] + corr
        for u in range(nu):
            tot, ncmp = np.zeros(nf), 0
            for mm in range(3):
                h = u - 10 + 36 * mm
                if h in B:
                    tot = tot + B[h]
                for k in range(7):
                    g = h + k
                    if 0 <= g < nb:
                        ncmp += (Cp[k] > 0) + 1 + (1 if (k > 0 and g > 0) else 0) + (1 if (k < 6 and g + 1 < nb) else 0)
            with np.errstate(all="ignore"):
                row = tot / ncmp if ncmp > 0 else np.full(nf, -np.inf)
            out[u * s + rho] = np.where(np.isfinite(row), row, -np.inf)
    return out


def score_map_drift(mag, tosr, fosr, drift_bins_per_symbol):
    """CPU statement of the drift-hypothesis Costas search (an EXTENSION of the new repo: BASELINE north_star names it
    for the satellite path, the reference has no counterpart - parity unpinned; with the single hypothesis 0.0 it is
    ft8_sync_score's map, ft8_decode.py:47-100, which is what pins this function).

    For hypothesis d the tone grid of symbol q (counted from the candidate's first symbol) is shifted by
    round-half-even(d * q) bins.  A hypothesis is scored with the reference's running sum over the shifted cells; cells
    whose symbol or bins leave the waterfall are skipped.  Every (abs_time, abs_freq) gets the best score over the
    hypotheses (first maximum wins) and that hypothesis' index.  NumPy, vectorised over frequency only: every map
    entry sees the reference's sequence of additions.  mag [F, T] as elsewhere in this module.
    Returns (best [nt, nf] float64, hyp [nt, nf] int32)."""
    m = _f64(mag)
    F, T = m.shape
    t0, t1, f0, f1 = search_ranges(F, T, tosr, fosr)
    nt, nf = max(t1 - t0, 0), max(f1 - f0, 0)
    best = np.full((nt, nf), -np.inf)
    hyp = np.zeros((nt, nf), dtype=np.int32)
    if nt == 0 or nf == 0:
        return best, hyp
    costas = [3, 1, 4, 0, 6, 5, 2]
    nb = T // tosr
    f = np.arange(f0, f1)
    for ti in range(nt):
        at = t0 + ti
        base = at // tosr                                   # Python floor division, as the reference's block index
        for h, d in enumerate(drift_bins_per_symbol):
            score = np.zeros(nf)
            num = np.zeros(nf, dtype=np.int64)
            for blk_i in range(3):
                for k in range(7):
                    q = blk_i * 36 + k
                    blk = base + q
                    if blk < 0 or blk >= nb:
                        continue
                    tone = costas[k]
                    bins = f + int(np.rint(d * q)) + tone * fosr
                    ok = (bins - (fosr if tone > 0 else 0) >= 0) & (bins + fosr < F)
                    bi = np.where(ok, bins, fosr)                               # any valid index where masked out
                    col = at + q * tosr
                    p = m[bi, col]

                    def add(term, cond):
                        nonlocal score, num
                        score = np.where(cond, score + term, score)
                        num = num + cond

                    if tone > 0:
                        add(p - m[bi - fosr, col], ok)
                    add(p - m[bi + fosr, col], ok)
                    if k > 0 and blk > 0:
                        add(p - m[bi, col - tosr], ok)
                    if k < 6 and blk + 1 < nb:
                        add(p - m[bi, col + tosr], ok)
            with np.errstate(invalid="ignore", divide="ignore"):
                v = np.where((num == 0) | ~np.isfinite(score), -np.inf, score / np.maximum(num, 1))
            better = v > best[ti]
            best[ti] = np.where(better, v, best[ti])
            hyp[ti] = np.where(better, h, hyp[ti])
    return best, hyp


def ft8_find_candidates(mag, tosr, fosr, num_candidates, min_score):
    """Returns (abs_time[int32], abs_freq[int32], score[f64], n_passers).  Raises TypeError on a heap tie."""
    m = _f64(mag)
    K = int(num_candidates)
    ct = np.zeros(max(K, 1), dtype=np.int32)
    cf = np.zeros(max(K, 1), dtype=np.int32)
    cs = np.zeros(max(K, 1), dtype=np.float64)
    npass = C.c_long(0)
    n = lib().orc_find_candidates(_dp(m), m.shape[0], m.shape[1], tosr, fosr, K, float(min_score),
                                  ct.ctypes.data_as(C.POINTER(C.c_int)), cf.ctypes.data_as(C.POINTER(C.c_int)),
                                  _dp(cs), C.byref(npass))
    if n == ERR_TIE:
        raise TypeError("exact score tie inside the candidate heap (reference raises TypeError)")
    if n < 0:
        raise ValueError("orc_find_candidates failed: %d" % n)
    return ct[:n].copy(), cf[:n].copy(), cs[:n].copy(), npass.value


def ft8_extract_likelihood(mag, tosr, fosr, abs_time, abs_freq, normalize=False):
    m = _f64(mag)
    out = np.zeros(LDPC_N, dtype=np.float64)
    lib().orc_extract_likelihood(_dp(m), m.shape[0], m.shape[1], tosr, fosr, abs_time, abs_freq, _dp(out))
    if normalize:
        with np.errstate(all="ignore"):
            lib().orc_normalize_logl(_dp(out))
    return out


def ftx_normalize_logl(log174):
    a = _f64(log174).copy()
    lib().orc_normalize_logl(_dp(a))
    return a


def ldpc_check(bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    return lib().orc_ldpc_check(b.ctypes.data_as(C.POINTER(C.c_uint8)))


def bp_decode(llr, max_iterations):
    a = _f64(llr)
    plain = np.zeros(LDPC_N, dtype=np.uint8)
    err = lib().orc_bp_decode(_dp(a), int(max_iterations), plain.ctypes.data_as(C.POINTER(C.c_uint8)))
    return plain, err


def bp_batch(llr, max_iterations, n_threads=1):
    a = _f64(llr).reshape(-1, LDPC_N)
    plain = np.zeros((a.shape[0], LDPC_N), dtype=np.uint8)
    err = np.zeros(a.shape[0], dtype=np.int32)
    lib().orc_bp_batch(_dp(a), a.shape[0], int(max_iterations), plain.ctypes.data_as(C.POINTER(C.c_uint8)),
                       err.ctypes.data_as(C.POINTER(C.c_int32)), int(n_threads))
    return plain, err


def compute_crc(message: bytes, num_bits: int) -> int:
    b = (C.c_uint8 * len(message)).from_buffer_copy(bytes(message))
    return lib().orc_compute_crc(b, num_bits)


def extract_crc(a91: bytes) -> int:
    b = (C.c_uint8 * len(a91)).from_buffer_copy(bytes(a91))
    return lib().orc_extract_crc(b)


def pack_bits(bits, num_bits):
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros((num_bits + 7) // 8, dtype=np.uint8)
    lib().orc_pack_bits(b.ctypes.data_as(C.POINTER(C.c_uint8)), num_bits, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return bytes(out)


def record_to_dict(r: Record) -> dict:
    return dict(abs_time=r.abs_time, abs_freq=r.abs_freq, score=r.score, ok=bool(r.ok), ldpc_errors=r.ldpc_errors,
                crc_extracted=r.crc_extracted, crc_calculated=r.crc_calculated, payload=bytes(r.payload),
                plain=np.frombuffer(bytes(r.plain), dtype=np.uint8).copy(),
                llr=np.array(r.llr[:], dtype=np.float64))


def ft8_decode_candidate(mag, tosr, fosr, abs_time, abs_freq, max_iterations):
    m = _f64(mag)
    r = Record()
    with np.errstate(all="ignore"):
        lib().orc_decode_candidate(_dp(m), m.shape[0], m.shape[1], tosr, fosr, abs_time, abs_freq, max_iterations,
                                   C.byref(r))
    return record_to_dict(r)


def masks(sample_rate, nperseg, hop, nfft, T, freq_min=None, freq_max=None, time_min=None, time_max=None):
    klo, khi, jlo, jhi = C.c_int(), C.c_int(), C.c_long(), C.c_long()
    lib().orc_masks(float(sample_rate), nperseg, hop, nfft, T, _none_nan(freq_min), _none_nan(freq_max),
                    _none_nan(time_min), _none_nan(time_max), C.byref(klo), C.byref(khi), C.byref(jlo), C.byref(jhi))
    return klo.value, khi.value, jlo.value, jhi.value


def decode_all_candidates(wave, sample_rate, bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=10,
                          max_iterations=20, freq_min=None, freq_max=None, time_min=None, time_max=None):
    """Records (dicts) for ALL candidates in the reference's candidate order.  Complex audio (the reference keeps the
    f >= 0 rows of its two-sided spectrogram, ft8_decode.py:323-325) goes in as interleaved (re, im) float64."""
    cplx = np.iscomplexobj(wave)
    x = np.ascontiguousarray(np.asarray(wave, dtype=np.complex128).reshape(-1)).view(np.float64) if cplx else _f64(wave)
    n_samples = len(x) // 2 if cplx else len(x)
    K = int(max_candidates)
    recs = (Record * max(K, 1))()
    F, T = C.c_int(), C.c_int()
    fn = lib().orc_decode_message_c if cplx else lib().orc_decode_message
    n = fn(_dp(x), n_samples, float(sample_rate), bins_per_tone, steps_per_symbol, K,
                                 float(min_score), max_iterations, _none_nan(freq_min), _none_nan(freq_max),
                                 _none_nan(time_min), _none_nan(time_max), recs, K, C.byref(F), C.byref(T))
    if n == ERR_TIE:
        raise TypeError("exact score tie inside the candidate heap (reference raises TypeError)")
    if n < 0:
        raise IndexError("input shorter than one analysis window (reference raises IndexError)")
    return [record_to_dict(recs[i]) for i in range(n)]


def decode_ft8_message(wave, sample_rate, bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=10,
                       max_iterations=20, freq_min=None, freq_max=None, time_min=None, time_max=None):
    """The reference's return value as plain tuples:
    [(payload bytes, crc, ldpc_errors, crc_extracted, time_sec, freq_hz, score, abs_time, abs_freq)]."""
    out = []
    for r in decode_all_candidates(wave, sample_rate, bins_per_tone, steps_per_symbol, max_candidates, min_score,
                                   max_iterations, freq_min, freq_max, time_min, time_max):
        if r["ok"]:
            out.append((r["payload"], r["crc_calculated"], r["ldpc_errors"], r["crc_extracted"],
                        r["abs_time"] / sample_rate, (r["abs_freq"] / bins_per_tone) * 6.25, r["score"],
                        r["abs_time"], r["abs_freq"]))
    return out


def decode_batch(waves, sample_rate, bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=10,
                 max_iterations=20, n_threads=1):
    """Threaded batch decode of equal-length slots; returns per-slot lists of record dicts."""
    x = _f64(waves)
    B, n = x.shape
    K = int(max_candidates)
    recs = (Record * (B * K))()
    ncand = np.zeros(B, dtype=np.int32)
    lib().orc_decode_batch(_dp(x), B, n, float(sample_rate), bins_per_tone, steps_per_symbol, K, float(min_score),
                           max_iterations, recs, ncand.ctypes.data_as(C.POINTER(C.c_int)), int(n_threads))
    out = []
    for i in range(B):
        if ncand[i] < 0:
            out.append(None)
        else:
            out.append([record_to_dict(recs[i * K + c]) for c in range(ncand[i])])
    return out
