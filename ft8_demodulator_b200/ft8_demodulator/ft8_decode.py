"""FT8 decode pipeline - the reference's `ft8_decode.py` API on the GPU.

Every public function keeps the name, arguments and return value of
`ft8_tools/ft8_demodulator/ft8_decode.py` (file:line cited per function); the arithmetic runs in the
CUDA library (csrc/stft.cu, sync.cu, decode.cu) through the C ABI of include/ft8b200.h.
`decode_slots` is the batched entry the reference lacks: same per-slot semantics,
many equal-length slots per call.

Deliberate differences: no printing, no PNG side effect; an input shorter
than one analysis window returns [] (the reference raises IndexError although its own test expects
[]); exact score ties, on which the reference raises TypeError, are ordered by scan position.
"""
from typing import List, Tuple

import numpy as np
import torch

from .. import _lib, engine
from .crc import compute_crc, extract_crc
from .ftx_types import FT8Candidate, FT8DecodeStatus, FT8Message, FT8Protocol, FT8Waterfall  # noqa: F401
from .ldpc_decoder import bp_decode  # noqa: F401  (the reference re-exports it from here)
from .spectrogram_analyse import (FT8_SYMBOL_DURATION_S, FT8_SYMBOL_FREQ_INTERVAL_HZ,  # noqa: F401
                                  _as_real_audio, calculate_spectrogram, time_axis)

FT8_ND = 58
FT8_NUM_SYNC = 3
FT8_LENGTH_SYNC = 7
FT8_SYNC_OFFSET = 36
FT8_LDPC_N = 174
FT8_LDPC_K = 91
FT8_LDPC_K_BYTES = 12
FT8_Gray_map = [0, 1, 3, 2, 5, 6, 4, 7]
FT8_Costas_pattern = [3, 1, 4, 0, 6, 5, 2]


# ------------------------------------------------------------------------------------------------
# waterfall staging
# ------------------------------------------------------------------------------------------------
def _device_waterfall(wf: FT8Waterfall) -> torch.Tensor:
    """[1, T, F] float64 device copy of wf.mag ([F, T] on the host), cached on the waterfall object.
    The waterfall is treated as immutable once it has been handed to a GPU function."""
    mag = wf.mag
    key = (id(mag), mag.shape, mag.ctypes.data if isinstance(mag, np.ndarray) else 0, str(mag.dtype))
    cached = wf.__dict__.get("_b200_dev")
    if cached is not None and cached[0] == key:
        return cached[1]
    dev = engine.require_cuda()
    host = np.ascontiguousarray(np.asarray(mag, dtype=np.float64).T)
    t = engine._dev(host.reshape(1, host.shape[0], host.shape[1]), device=dev)
    wf.__dict__["_b200_dev"] = (key, t)
    wf.__dict__.pop("_b200_scores", None)
    return t


def _score_map(wf: FT8Waterfall):
    cached = wf.__dict__.get("_b200_scores")
    mag = _device_waterfall(wf)
    if cached is None:
        scores, t0 = engine.sync_scores(mag, wf.time_osr, wf.freq_osr)
        cached = (scores, t0, None)
        wf.__dict__["_b200_scores"] = cached
    return cached


# ------------------------------------------------------------------------------------------------
# ft8_decode.py:47-149
# ------------------------------------------------------------------------------------------------
def ft8_sync_score(wf: FT8Waterfall, candidate: FT8Candidate) -> float:
    """Costas sync score of one candidate (ft8_decode.py:47-100), evaluated on the GPU exactly as the reference does
    (one running sum, one division: ft8b_sync_score_at) - bit-equal on the same waterfall, at any abs_time.  A
    frequency whose eight tone bins do not fit the waterfall raises IndexError, as the reference's indexing does."""
    _check_candidate(wf, candidate)
    mag = _device_waterfall(wf)
    ct, cf, _ = _cand_tensors(candidate, mag.device)
    return float(engine.sync_score_at(mag, wf.time_osr, wf.freq_osr, ct, cf).cpu()[0, 0])


def ft8_find_candidates(wf: FT8Waterfall, num_candidates: int, min_score: int, strict_topk: bool = False) -> List[FT8Candidate]:
    """Candidate search with the reference's exact selection semantics (ft8_decode.py:102-149).
    strict_topk=True (opt-in extension): the num_candidates best scores instead of the reference's
    "first K passers in scan order + one record replacement" list."""
    scores, t0, _ = _score_map(wf)
    ct, cf, cs, nc, _ = engine.select_candidates(scores, t0, int(num_candidates), float(min_score), strict_topk=strict_topk,
                                                 mag=_device_waterfall(wf), time_osr=wf.time_osr, freq_osr=wf.freq_osr)
    n = int(nc.cpu()[0])
    ct, cf, cs = ct[0, :n].cpu().numpy(), cf[0, :n].cpu().numpy(), cs[0, :n].cpu().numpy()
    return [FT8Candidate(waterfall=wf, abs_time=int(ct[i]), abs_freq=int(cf[i]), score=float(cs[i])) for i in range(n)]


# ------------------------------------------------------------------------------------------------
# ft8_decode.py:151-215
# ------------------------------------------------------------------------------------------------
def ft8_extract_symbol(wf: np.ndarray, logl: np.ndarray) -> None:
    """Max-difference LLRs of one 8-FSK symbol (ft8_decode.py:151-161).  Host utility on 8 values; the
    decode path does this inside csrc/decode.cu."""
    s2 = [wf[g] for g in FT8_Gray_map]
    logl[0] = max(s2[4], s2[5], s2[6], s2[7]) - max(s2[0], s2[1], s2[2], s2[3])
    logl[1] = max(s2[2], s2[3], s2[6], s2[7]) - max(s2[0], s2[1], s2[4], s2[5])
    logl[2] = max(s2[1], s2[3], s2[5], s2[7]) - max(s2[0], s2[2], s2[4], s2[6])


def _cand_tensors(cand: FT8Candidate, dev):
    ct = torch.tensor([[cand.abs_time]], dtype=torch.int32, device=dev)
    cf = torch.tensor([[cand.abs_freq]], dtype=torch.int32, device=dev)
    nc = torch.tensor([1], dtype=torch.int32, device=dev)
    return ct, cf, nc


def _check_candidate(wf: FT8Waterfall, cand: FT8Candidate) -> None:
    if not (0 <= cand.abs_freq and cand.abs_freq + 7 * wf.freq_osr < wf.mag.shape[0]):
        raise IndexError("candidate frequency outside the waterfall")


def ft8_extract_likelihood(wf: FT8Waterfall, cand: FT8Candidate, log174: np.ndarray) -> None:
    """174 un-normalised LLRs written into log174 (ft8_decode.py:164-188)."""
    _check_candidate(wf, cand)
    mag = _device_waterfall(wf)
    ct, cf, nc = _cand_tensors(cand, mag.device)
    log174[:FT8_LDPC_N] = engine.extract_llr(mag, wf.time_osr, wf.freq_osr, ct, cf, nc, normalize=False)[0, 0].cpu().numpy()


def ftx_normalize_logl(log174: np.ndarray) -> None:
    """Scale log174 in place to variance 24 (ft8_decode.py:190-198)."""
    dev = engine.require_cuda()
    t = engine._dev(np.asarray(log174, dtype=np.float64).reshape(1, FT8_LDPC_N).copy(), device=dev)
    log174[:FT8_LDPC_N] = engine.normalize_llr_(t)[0].cpu().numpy()


def pack_bits(bit_array: np.ndarray, num_bits: int) -> bytearray:
    """MSB-first bit packing (ft8_decode.py:200-215); host byte utility."""
    bits = (np.asarray(bit_array).reshape(-1)[:num_bits] != 0).astype(np.uint8)
    return bytearray(np.packbits(bits).tobytes())


def ftx_compute_crc(data: bytearray, num_bits: int) -> int:
    return compute_crc(data, num_bits)


def ftx_extract_crc(data: bytearray) -> int:
    return extract_crc(data)


# ------------------------------------------------------------------------------------------------
# ft8_decode.py:225-286
# ------------------------------------------------------------------------------------------------
def ft8_decode_candidate(wf: FT8Waterfall, cand: FT8Candidate, max_iterations: int
                         ) -> Tuple[bool, FT8Message, FT8DecodeStatus]:
    """LLR extraction + LDPC + CRC for one candidate (ft8_decode.py:225-273)."""
    _check_candidate(wf, cand)
    mag = _device_waterfall(wf)
    ct, cf, nc = _cand_tensors(cand, mag.device)
    d = engine.decode_candidates(mag, wf.time_osr, wf.freq_osr, ct, cf, nc, int(max_iterations))
    message, status = FT8Message(), FT8DecodeStatus()
    status.ldpc_errors = int(d.ldpc_errors.cpu()[0, 0])
    status.crc_extracted = int(d.crc_extracted.cpu()[0, 0])
    status.crc_calculated = int(d.crc_calculated.cpu()[0, 0])
    ok = bool(d.ok.cpu()[0, 0])
    if ok:
        message.hash = status.crc_calculated
        message.payload = bytearray(d.payload[0, 0].cpu().numpy().tobytes())
    return ok, message, status


def create_waterfall_from_spectrogram(spectrogram: np.ndarray, time_osr: int, freq_osr: int) -> FT8Waterfall:
    """Wrap a [frequency, time] spectrogram (ft8_decode.py:275-286)."""
    if len(spectrogram.shape) != 2:
        raise ValueError("spectrogram must be a 2D array with shape (frequency, time)")
    return FT8Waterfall(mag=spectrogram, time_osr=time_osr, freq_osr=freq_osr)


# ------------------------------------------------------------------------------------------------
# ft8_decode.py:288-394 and the batched form
# ------------------------------------------------------------------------------------------------
def decode_geometry(n_samples: int, sample_rate, bins_per_tone: int, steps_per_symbol: int, freq_min=None,
                    freq_max=None, time_min=None, time_max=None):
    """Host-side index arithmetic of decode_ft8_message: window geometry and the inclusive frequency /
    time masks of ft8_decode.py:323-341 turned into contiguous [lo, hi) ranges.
    Returns dict(nperseg, hop, nfft, T_all, bin_lo, bin_hi, frame_lo, frame_hi)."""
    nperseg, hop, nfft = _lib.plan(sample_rate, bins_per_tone, steps_per_symbol)
    T_all = _lib.num_frames(n_samples, nperseg, hop)
    n_pos = (nfft - 1) // 2 + 1
    bin_lo, bin_hi = 0, n_pos
    if freq_min is not None or freq_max is not None:
        f = np.fft.fftfreq(nfft, 1 / sample_rate)[:n_pos]          # the f >= 0 rows, NumPy's rounding
        fmin = freq_min if freq_min is not None else f[0]
        fmax = freq_max if freq_max is not None else f[-1]
        idx = np.nonzero((f >= fmin) & (f <= fmax))[0]
        bin_lo, bin_hi = (int(idx[0]), int(idx[-1]) + 1) if len(idx) else (0, 0)
    frame_lo, frame_hi = 0, T_all
    if T_all > 0 and (time_min is not None or time_max is not None):
        t = time_axis(n_samples, nperseg, hop, sample_rate)
        tmin = time_min if time_min is not None else t[0]
        tmax = time_max if time_max is not None else t[-1]
        idx = np.nonzero((t >= tmin) & (t <= tmax))[0]
        frame_lo, frame_hi = (int(idx[0]), int(idx[-1]) + 1) if len(idx) else (0, 0)
    return dict(nperseg=nperseg, hop=hop, nfft=nfft, T_all=T_all, bin_lo=bin_lo, bin_hi=bin_hi, frame_lo=frame_lo,
                frame_hi=frame_hi)


def _results_from_arrays(out, slot: int, sample_rate, freq_osr: int) -> list:
    """The reference's return list for one slot: successes in candidate order, 5-tuples (ft8_decode.py:383-391).
    `out`: SlotDecodes of host tensors, or the dict of NumPy views BatchDecodes keeps of them."""
    a = out if isinstance(out, dict) else {f: getattr(out, f).numpy() for f in out.__dataclass_fields__}
    n = int(a["n_cand"][slot])
    res = []
    for c in np.flatnonzero(a["ok"][slot, :n]).tolist():       # scalar reads only for the successes
        crc = int(a["crc_calculated"][slot, c])
        msg = FT8Message(payload=bytearray(a["payload"][slot, c].tobytes()), hash=crc)
        st = FT8DecodeStatus(ldpc_errors=int(a["ldpc_errors"][slot, c]), crc_extracted=int(a["crc_extracted"][slot, c]),
                             crc_calculated=crc)
        abs_time, abs_freq = int(a["cand_time"][slot, c]), int(a["cand_freq"][slot, c])
        res.append((msg, st, abs_time / sample_rate, (abs_freq / freq_osr) * FT8_SYMBOL_FREQ_INTERVAL_HZ,
                    float(a["cand_score"][slot, c])))
    return res


class BatchDecodes:
    """Host-side result of decode_slots: packed arrays plus lazy per-slot materialisation."""

    def __init__(self, arrays: engine.SlotDecodes, sample_rate, freq_osr: int):
        self.arrays, self.sample_rate, self.freq_osr = arrays, sample_rate, freq_osr
        self._np = None             # NumPy views of the host tensors, made on first use

    def __len__(self):
        return self.arrays.n_cand.shape[0]

    def __getitem__(self, slot: int) -> list:
        if self._np is None:
            self._np = {f: getattr(self.arrays, f).numpy() for f in self.arrays.__dataclass_fields__}
        if not -len(self) <= slot < len(self):
            raise IndexError(slot)
        return _results_from_arrays(self._np, slot, self.sample_rate, self.freq_osr)

    def to_lists(self) -> list:
        return [self[i] for i in range(len(self))]

    def payload_sets(self) -> list:
        """Per slot: set of decoded payloads (bytes)."""
        a = self.arrays
        ok = a.ok.numpy().astype(bool)
        pl = a.payload.numpy()
        return [set(bytes(pl[i, c]) for c in np.nonzero(ok[i])[0]) for i in range(len(self))]


def decode_slots(waves, sample_rate, bins_per_tone: int = 2, steps_per_symbol: int = 2, max_candidates: int = 20,
                 min_score=10, max_iterations: int = 20, freq_min=None, freq_max=None, time_min=None, time_max=None,
                 chunk_slots: int = 0, return_device: bool = False, piece_slots: int = 256, pcm16: bool = False,
                 strict_topk: bool = False):
    """Decode a batch of equal-length slots: waves [B, N] (NumPy or torch, host or device, float32/float64, or
    complex64/complex128 baseband - the reference keeps the f >= 0 rows of the two-sided spectrogram then).

    With pcm16=True an int16 array is taken as PCM and normalised on the GPU as float32(x)/32767, the way the
    reference's WAV reader does on the host (src/tests/demodulator/from_wave.py:62-66) - half the host-to-device
    bytes of float32.  Without it integers are converted to float64 unscaled, as the reference would treat them.
    Per-slot semantics are exactly decode_ft8_message's.  Host input is copied to the GPU inside this
    call (pinned torch tensors copy asynchronously); results come back as a BatchDecodes (host) or, with
    return_device=True, as the engine.SlotDecodes device tensors.
    strict_topk=True (opt-in extension, not the reference's behaviour): every slot's candidate list is its
    max_candidates best sync scores instead of the reference's scan-order list (ft8_decode.py:127-140)."""
    dev = engine.require_cuda()
    if isinstance(waves, np.ndarray):
        if np.iscomplexobj(waves):
            # complex baseband, as the reference's SDR receive script passes it (src/tests/pluto-sdr/receive.py:68-78):
            # two-sided spectrogram, rows f >= 0 kept (ft8_decode.py:316-325) - a full-length complex transform in S1
            if waves.dtype != np.complex64:
                waves = waves.astype(np.complex128)
        elif waves.dtype not in (np.float32, np.float64) and not (pcm16 and waves.dtype == np.int16):
            waves = waves.astype(np.float64)
        waves = np.ascontiguousarray(waves)
        waves = torch.from_numpy(waves if waves.flags.writeable else waves.copy())
    if waves.dim() == 1:
        waves = waves.reshape(1, -1)
    if waves.dtype not in (torch.float32, torch.float64, torch.complex64, torch.complex128) and \
            not (pcm16 and waves.dtype == torch.int16):
        waves = waves.to(torch.float64)
    B, N = waves.shape
    g = decode_geometry(N, sample_rate, bins_per_tone, steps_per_symbol, freq_min, freq_max, time_min, time_max)
    params = engine.make_params(sample_rate, bins_per_tone, steps_per_symbol, max_candidates, min_score, max_iterations,
                                g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"], chunk_slots,
                                flags=_lib.FLAG_STRICT_TOPK if strict_topk else 0)
    if max_candidates > _lib.MAX_CANDIDATES:
        raise NotImplementedError("max_candidates=%d exceeds the kernel limit %d" % (max_candidates, _lib.MAX_CANDIDATES))
    if B == 0 or g["T_all"] <= 0:
        out = engine.alloc_slot_decodes(B, int(max_candidates), dev)
    elif waves.is_cuda:
        out = engine.SlotDecoder(B, N, params, device=dev).run(waves if waves.stride(1) == 1 else waves.contiguous())
    elif return_device:
        out = engine.decode_host_batch(waves.contiguous(), params, piece_slots=piece_slots, device=dev)
    else:       # host in, host out: the library's own host path (ft8b_decode_slots_host), no torch CUDA tensors involved
        out = engine.decode_host_native(waves if waves.stride(1) == 1 else waves.contiguous(), params,
                                        piece_slots=piece_slots, device=dev)
        return BatchDecodes(out, sample_rate, bins_per_tone)
    if return_device:
        return out
    return BatchDecodes(out.cpu(), sample_rate, bins_per_tone)


def decode_slots_drift(waves, sample_rate, drift_hz_per_s, bins_per_tone: int = 2, steps_per_symbol: int = 2,
                       max_candidates: int = 20, min_score=10, max_iterations: int = 20, freq_min=None, freq_max=None,
                       time_min=None, time_max=None, slots_per_pass: int = 256, strict_topk: bool = False):
    """EXTENSION (not in the reference; BASELINE north_star's drift-hypothesis sync for the satellite path):
    decode_slots with a Costas search over linear Doppler-drift hypotheses.

    `drift_hz_per_s` lists the hypotheses; 0.0 is always tried first, so a drift-free signal decodes exactly as in
    decode_slots (with drift_hz_per_s=[0.0] the two functions return identical results).  For hypothesis r the tone
    grid of symbol q of a candidate is shifted by round(r * 0.16 s * q / bin_width) bins, both in the sync score and
    in the LLR extraction.  strict_topk as in decode_slots (the K best scores instead of the reference's scan-order
    list).  Returns (BatchDecodes, drift_hz_per_s of every candidate as float64 [B, K])."""
    dev = engine.require_cuda()
    waves = waves if isinstance(waves, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(waves)))
    if waves.dim() == 1:
        waves = waves.reshape(1, -1)
    if waves.dtype not in (torch.float32, torch.float64):
        waves = waves.to(torch.float64)
    B, N = waves.shape
    K = int(max_candidates)
    rates = [0.0] + [float(r) for r in drift_hz_per_s if float(r) != 0.0]
    bin_hz = FT8_SYMBOL_FREQ_INTERVAL_HZ / bins_per_tone
    d_bins = [r * FT8_SYMBOL_DURATION_S / bin_hz for r in rates]
    g = decode_geometry(N, sample_rate, bins_per_tone, steps_per_symbol, freq_min, freq_max, time_min, time_max)
    out = engine.alloc_slot_decodes(B, K, dev)
    drift = torch.zeros((B, K), dtype=torch.float64, device=dev)
    if B and g["T_all"] > 0 and g["bin_hi"] > g["bin_lo"] and g["frame_hi"] > g["frame_lo"] and K > 0:
        rate_t = torch.tensor(rates, dtype=torch.float64, device=dev)
        dbin_t = torch.tensor(d_bins, dtype=torch.float64, device=dev)
        for lo in range(0, B, slots_per_pass):
            hi = min(lo + slots_per_pass, B)
            w = waves[lo:hi].to(dev).contiguous()
            mag = engine.stft_db(w, g["nperseg"], g["hop"], g["nfft"], g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"])
            score, hyp, t0 = engine.sync_scores_drift(mag, steps_per_symbol, bins_per_tone, d_bins)
            if score.numel() == 0:
                continue
            ct, cf, cs, nc, _ = engine.select_candidates(score, t0, K, float(min_score), strict_topk=strict_topk)
            live = torch.arange(K, device=dev)[None, :] < nc[:, None]
            ti = torch.where(live, ct - t0, torch.zeros_like(ct)).long()
            h = hyp[torch.arange(hi - lo, device=dev)[:, None], ti, cf.long()].long()      # hypothesis of each candidate
            dec = engine.decode_candidates_drift(mag, steps_per_symbol, bins_per_tone, ct, cf, nc, dbin_t[h], max_iterations)
            sl = engine.slice_slot_decodes(out, lo, hi)
            sl.cand_time.copy_(ct); sl.cand_freq.copy_(cf); sl.cand_score.copy_(cs); sl.n_cand.copy_(nc)
            sl.ok.copy_(dec.ok); sl.ldpc_errors.copy_(dec.ldpc_errors); sl.crc_extracted.copy_(dec.crc_extracted)
            sl.crc_calculated.copy_(dec.crc_calculated); sl.payload.copy_(dec.payload)
            drift[lo:hi] = torch.where(live, rate_t[h], torch.zeros_like(rate_t[h]))
    return BatchDecodes(out.cpu(), sample_rate, bins_per_tone), drift.cpu().numpy()


def decode_ft8_message(wave_data: np.ndarray, sample_rate: int, bins_per_tone: int = 2, steps_per_symbol: int = 2,
                       max_candidates: int = 20, min_score: int = 10, max_iterations: int = 20,
                       freq_min: float = None, freq_max: float = None, time_min: float = None, time_max: float = None,
                       strict_topk: bool = False) -> List[Tuple[FT8Message, FT8DecodeStatus, float, float, float]]:
    """Decode one waveform (ft8_decode.py:288-394).  Returns the reference's list of
    (FT8Message, FT8DecodeStatus, time_sec, freq_hz, score) in candidate (score-descending) order, with
    time_sec = abs_time / sample_rate and freq_hz = abs_freq / bins_per_tone * 6.25 as the reference defines them.
    strict_topk is the one extra keyword (default off = the reference's candidate list), see decode_slots."""
    if np.iscomplexobj(wave_data):
        x = np.asarray(wave_data).reshape(-1)
        x = np.ascontiguousarray(x if x.dtype == np.complex64 else x.astype(np.complex128))
    else:
        x = _as_real_audio(wave_data)
    nperseg, _, _ = _lib.plan(sample_rate, bins_per_tone, steps_per_symbol)
    if len(x) < nperseg:
        return []
    return decode_slots(x.reshape(1, -1), sample_rate, bins_per_tone, steps_per_symbol, max_candidates, min_score,
                        max_iterations, freq_min, freq_max, time_min, time_max, strict_topk=strict_topk)[0]


__all__ = [
    "ft8_sync_score", "ft8_find_candidates", "ft8_extract_symbol", "ft8_extract_likelihood", "ftx_normalize_logl",
    "pack_bits", "ftx_compute_crc", "ftx_extract_crc", "ft8_decode_candidate", "create_waterfall_from_spectrogram",
    "decode_ft8_message", "decode_slots", "decode_slots_drift", "decode_geometry", "BatchDecodes", "FT8Waterfall", "FT8Candidate",
    "FT8Message", "FT8DecodeStatus", "FT8Protocol", "bp_decode", "calculate_spectrogram",
]
