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
    """dB spectrogram rows f >= 0, shape [ceil(nfft/2), T] (what ft8_decode.py:323-325 keeps)."""
    x = _f64(wave)
    nperseg, hop, nfft = plan(sample_rate, bins_per_tone, steps_per_symbol)
    T = lib().orc_num_frames(len(x), nperseg, hop)
    if T <= 0:
        return np.zeros((0, 0))
    mag = np.empty(((nfft + 1) // 2, T), dtype=np.float64)
    rc = lib().orc_spectrogram_pos(_dp(x), len(x), nperseg, hop, nfft, _dp(mag))
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
    """Records (dicts) for ALL candidates in the reference's candidate order."""
    x = _f64(wave)
    K = int(max_candidates)
    recs = (Record * max(K, 1))()
    F, T = C.c_int(), C.c_int()
    n = lib().orc_decode_message(_dp(x), len(x), float(sample_rate), bins_per_tone, steps_per_symbol, K,
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
