"""Device-side plumbing: torch tensors in, C-ABI calls on the current CUDA stream, torch tensors out.

PyTorch is used only for device memory, streams and copies; every arithmetic step is one of the
hand-written sm_100a kernels behind include/ft8b200.h.  Nothing here falls back to the CPU.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

_F64, _F32 = torch.float64, torch.float32


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("ft8_demodulator_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int:
    return 0 if t is None else t.data_ptr()


def _dev(t, dtype=None, device=None):
    """Contiguous tensor on the CUDA device (copies host data with a pinned, async H2D when needed)."""
    device = device or require_cuda()
    if isinstance(t, np.ndarray):
        t = np.ascontiguousarray(t)
        t = torch.from_numpy(t if t.flags.writeable else t.copy())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def wave_dtype_code(t: torch.Tensor) -> int:
    if t.dtype == _F32:
        return _lib.DTYPE_F32
    if t.dtype == _F64:
        return _lib.DTYPE_F64
    if t.dtype == torch.int16:
        return _lib.DTYPE_I16          # PCM, normalised on the GPU as float32(x)/32767 (from_wave.py:62-66)
    if t.dtype == torch.complex64:
        return _lib.DTYPE_C64
    if t.dtype == torch.complex128:
        return _lib.DTYPE_C128
    raise TypeError("audio must be float32, float64 or int16 PCM, got %s" % t.dtype)


def stft_db(wave: torch.Tensor, nperseg: int, hop: int, nfft: int, bin_lo: int, bin_hi: int, frame_lo: int,
            frame_hi: int) -> torch.Tensor:
    """S1.  wave [B, N] (f32/f64, CUDA) -> mag [B, frame_hi-frame_lo, bin_hi-bin_lo] float64."""
    assert wave.is_cuda and wave.dim() == 2 and wave.stride(1) == 1
    B, N = wave.shape
    mag = torch.empty((B, frame_hi - frame_lo, bin_hi - bin_lo), dtype=_F64, device=wave.device)
    _lib.check(_lib.lib().ft8b_stft_db(_p(wave), wave_dtype_code(wave), B, N, wave.stride(0), nperseg, hop, nfft, bin_lo,
                                       bin_hi, frame_lo, frame_hi, _p(mag), _stream()), "ft8b_stft_db")
    return mag


def sync_scores(mag: torch.Tensor, time_osr: int, freq_osr: int):
    """S2a.  mag [B, T, F] -> (scores [B, nt, nf] float64, t0).  Empty tensor when the search range is empty."""
    assert mag.is_cuda and mag.dtype == _F64 and mag.is_contiguous() and mag.dim() == 3
    B, T, F = mag.shape
    t0, nt, nf = _lib.search_ranges(T, F, time_osr, freq_osr)
    nt, nf = max(nt, 0), max(nf, 0)
    scores = torch.empty((B, nt, nf), dtype=_F64, device=mag.device)
    if B and nt and nf:
        _lib.check(_lib.lib().ft8b_sync_scores(_p(mag), B, T, F, time_osr, freq_osr, _p(scores), _stream()),
                   "ft8b_sync_scores")
    return scores, t0


def sync_score_at(mag: torch.Tensor, time_osr: int, freq_osr: int, cand_time: torch.Tensor, cand_freq: torch.Tensor):
    """ft8_sync_score at given positions: mag [B, T, F], cand_time / cand_freq i32 [B, n] -> scores f64 [B, n], each the
    reference's own running sum (bit-equal on the same spectrogram; ft8b_sync_score_at)."""
    assert mag.is_cuda and mag.dtype == _F64 and mag.is_contiguous() and mag.dim() == 3
    B, T, F = mag.shape
    assert cand_time.shape == cand_freq.shape and cand_time.shape[0] == B and cand_time.dtype == cand_freq.dtype == torch.int32
    assert cand_time.is_contiguous() and cand_freq.is_contiguous()
    n = cand_time.shape[1]
    out = torch.empty((B, n), dtype=_F64, device=mag.device)
    if B and n:
        _lib.check(_lib.lib().ft8b_sync_score_at(_p(mag), B, T, F, time_osr, freq_osr, _p(cand_time), _p(cand_freq), n,
                                                 _p(out), _stream()), "ft8b_sync_score_at")
    return out


def select_candidates(scores: torch.Tensor, t0: int, K: int, min_score: float, strict_topk: bool = False,
                      mag: torch.Tensor = None, time_osr: int = 0, freq_osr: int = 0):
    """S2b.  scores [B, nt, nf] -> (cand_time i32 [B,K], cand_freq i32 [B,K], cand_score f64 [B,K], n_cand i32 [B],
    n_pass i32 [B]).  strict_topk (opt-in extension): the K best passers instead of the reference's list; n_pass is
    not counted then (zeros).  With `mag` [B, T, F] (the spectrogram the map was made from, plus its oversampling) the
    scores that decide or are returned are re-evaluated in the reference's own summation order (ft8b_find_candidates)."""
    assert scores.is_cuda and scores.dtype == _F64 and scores.is_contiguous() and scores.dim() == 3
    B, nt, nf = scores.shape
    dev = scores.device
    ct = torch.zeros((B, K), dtype=torch.int32, device=dev)
    cf = torch.zeros((B, K), dtype=torch.int32, device=dev)
    cs = torch.zeros((B, K), dtype=_F64, device=dev)
    nc = torch.zeros((B,), dtype=torch.int32, device=dev)
    npass = torch.zeros((B,), dtype=torch.int32, device=dev)
    if strict_topk:
        _lib.check(_lib.lib().ft8b_select_topk(_p(scores) if scores.numel() else 0, B, nt, nf, t0, K, float(min_score),
                                               _p(ct), _p(cf), _p(cs), _p(nc), _stream()), "ft8b_select_topk")
    elif mag is not None and scores.numel():
        assert mag.is_cuda and mag.dtype == _F64 and mag.is_contiguous() and mag.dim() == 3 and mag.shape[0] == B
        _lib.check(_lib.lib().ft8b_find_candidates(_p(mag), _p(scores), B, mag.shape[1], mag.shape[2], time_osr, freq_osr, K,
                                                   float(min_score), _p(ct), _p(cf), _p(cs), _p(nc), _p(npass), _stream()),
                   "ft8b_find_candidates")
    else:
        _lib.check(_lib.lib().ft8b_select_candidates(_p(scores) if scores.numel() else 0, B, nt, nf, t0, K, float(min_score),
                                                     _p(ct), _p(cf), _p(cs), _p(nc), _p(npass), _stream()),
                   "ft8b_select_candidates")
    return ct, cf, cs, nc, npass


@dataclass
class CandidateDecodes:
    ok: torch.Tensor             # u8  [B,K]
    ldpc_errors: torch.Tensor    # i32 [B,K]
    crc_extracted: torch.Tensor  # i32 [B,K]
    crc_calculated: torch.Tensor  # i32 [B,K]
    payload: torch.Tensor        # u8  [B,K,10]
    llr: torch.Tensor = None     # f64 [B,K,174] (optional)
    plain: torch.Tensor = None   # u8  [B,K,174] (optional)


def decode_candidates(mag, time_osr, freq_osr, cand_time, cand_freq, n_cand, max_iterations, want_llr=False,
                      want_plain=False) -> CandidateDecodes:
    """S3 on spectrogram + candidate lists."""
    assert mag.is_cuda and mag.dtype == _F64 and mag.is_contiguous()
    B, T, F = mag.shape
    K = cand_time.shape[1]
    dev = mag.device
    out = CandidateDecodes(
        ok=torch.zeros((B, K), dtype=torch.uint8, device=dev),
        ldpc_errors=torch.zeros((B, K), dtype=torch.int32, device=dev),
        crc_extracted=torch.zeros((B, K), dtype=torch.int32, device=dev),
        crc_calculated=torch.zeros((B, K), dtype=torch.int32, device=dev),
        payload=torch.zeros((B, K, 10), dtype=torch.uint8, device=dev),
        llr=torch.zeros((B, K, 174), dtype=_F64, device=dev) if want_llr else None,
        plain=torch.zeros((B, K, 174), dtype=torch.uint8, device=dev) if want_plain else None)
    if B and K:
        _lib.check(_lib.lib().ft8b_decode_candidates(
            _p(mag), B, T, F, time_osr, freq_osr, _p(cand_time), _p(cand_freq), _p(n_cand), K, int(max_iterations),
            _p(out.ok), _p(out.ldpc_errors), _p(out.crc_extracted), _p(out.crc_calculated), _p(out.payload), _p(out.llr),
            _p(out.plain), _stream()), "ft8b_decode_candidates")
    return out


def sync_scores_drift(mag: torch.Tensor, time_osr: int, freq_osr: int, drift_bins_per_symbol, variant: int = 0):
    """Extension: best Costas score over linear-drift hypotheses.  Returns (best_score [B,nt,nf] f64,
    best_hyp [B,nt,nf] i32, t0).  `variant` forces one of the three kernels (_lib.DRIFT_WALK / DRIFT_BLOCKSUM /
    DRIFT_DIRECT; 0 = the library's choice) - for the cross-check tests and the A/B bench."""
    assert mag.is_cuda and mag.dtype == _F64 and mag.is_contiguous() and mag.dim() == 3
    B, T, F = mag.shape
    t0, nt, nf = _lib.search_ranges(T, F, time_osr, freq_osr)
    nt, nf = max(nt, 0), max(nf, 0)
    score = torch.empty((B, nt, nf), dtype=_F64, device=mag.device)
    hyp = torch.empty((B, nt, nf), dtype=torch.int32, device=mag.device)
    d = (C.c_double * len(drift_bins_per_symbol))(*[float(x) for x in drift_bins_per_symbol])
    if B and nt and nf:
        _lib.check(_lib.lib().ft8b_sync_scores_drift_variant(_p(mag), B, T, F, time_osr, freq_osr, len(drift_bins_per_symbol),
                                                             d, _p(score), _p(hyp), int(variant), _stream()),
                   "ft8b_sync_scores_drift")
        torch.cuda.current_stream().synchronize()      # the hypothesis table lives in constant memory: one call at a time
    return score, hyp, t0


def decode_candidates_drift(mag, time_osr, freq_osr, cand_time, cand_freq, n_cand, cand_drift, max_iterations) -> CandidateDecodes:
    """Extension: S3 with the tone grid of each candidate following its drift (bins per symbol)."""
    assert mag.is_cuda and mag.dtype == _F64 and mag.is_contiguous() and cand_drift.dtype == _F64
    B, T, F = mag.shape
    K = cand_time.shape[1]
    dev = mag.device
    out = CandidateDecodes(
        ok=torch.zeros((B, K), dtype=torch.uint8, device=dev), ldpc_errors=torch.zeros((B, K), dtype=torch.int32, device=dev),
        crc_extracted=torch.zeros((B, K), dtype=torch.int32, device=dev),
        crc_calculated=torch.zeros((B, K), dtype=torch.int32, device=dev),
        payload=torch.zeros((B, K, 10), dtype=torch.uint8, device=dev))
    if B and K:
        _lib.check(_lib.lib().ft8b_decode_candidates_drift(
            _p(mag), B, T, F, time_osr, freq_osr, _p(cand_time), _p(cand_freq), _p(n_cand), _p(cand_drift.contiguous()), K,
            int(max_iterations), _p(out.ok), _p(out.ldpc_errors), _p(out.crc_extracted), _p(out.crc_calculated),
            _p(out.payload), _stream()), "ft8b_decode_candidates_drift")
    return out


def extract_llr(mag, time_osr, freq_osr, cand_time, cand_freq, n_cand, normalize: bool) -> torch.Tensor:
    B, T, F = mag.shape
    K = cand_time.shape[1]
    llr = torch.zeros((B, K, 174), dtype=_F64, device=mag.device)
    s8 = torch.zeros((B, K), dtype=torch.uint8, device=mag.device)
    s32 = torch.zeros((B, K), dtype=torch.int32, device=mag.device)
    if B and K:
        _lib.check(_lib.lib().ft8b_extract_llr(_p(mag), B, T, F, time_osr, freq_osr, _p(cand_time), _p(cand_freq),
                                               _p(n_cand), K, 1 if normalize else 0, _p(llr), _p(s8), _p(s32), _stream()),
                   "ft8b_extract_llr")
    return llr


def ldpc_crc(llr: torch.Tensor, max_iterations: int) -> CandidateDecodes:
    """S3' on ready-made LLR vectors [n,174] float64."""
    assert llr.is_cuda and llr.dtype == _F64 and llr.is_contiguous() and llr.shape[-1] == 174
    n = llr.numel() // 174
    dev = llr.device
    out = CandidateDecodes(
        ok=torch.zeros((n,), dtype=torch.uint8, device=dev), ldpc_errors=torch.zeros((n,), dtype=torch.int32, device=dev),
        crc_extracted=torch.zeros((n,), dtype=torch.int32, device=dev),
        crc_calculated=torch.zeros((n,), dtype=torch.int32, device=dev),
        payload=torch.zeros((n, 10), dtype=torch.uint8, device=dev), llr=llr,
        plain=torch.zeros((n, 174), dtype=torch.uint8, device=dev))
    if n:
        _lib.check(_lib.lib().ft8b_ldpc_crc(_p(llr), n, int(max_iterations), _p(out.plain), _p(out.ldpc_errors),
                                            _p(out.crc_extracted), _p(out.crc_calculated), _p(out.payload), _p(out.ok),
                                            _stream()), "ft8b_ldpc_crc")
    return out


def normalize_llr_(llr: torch.Tensor) -> torch.Tensor:
    assert llr.is_cuda and llr.dtype == _F64 and llr.is_contiguous() and llr.shape[-1] == 174
    _lib.check(_lib.lib().ft8b_normalize_llr(_p(llr), llr.numel() // 174, _stream()), "ft8b_normalize_llr")
    return llr


def ldpc_check(bits: torch.Tensor) -> torch.Tensor:
    assert bits.is_cuda and bits.dtype == torch.uint8 and bits.is_contiguous() and bits.shape[-1] == 174
    n = bits.numel() // 174
    err = torch.zeros((n,), dtype=torch.int32, device=bits.device)
    _lib.check(_lib.lib().ft8b_ldpc_check(_p(bits), n, _p(err), _stream()), "ft8b_ldpc_check")
    return err


def selfcheck_divide(a: torch.Tensor, b: torch.Tensor):
    """(mismatches, examined) of the BP kernel's branch-free division against IEEE a / b (ft8b_selfcheck_divide)."""
    assert a.is_cuda and b.is_cuda and a.dtype == _F64 and b.dtype == _F64 and a.is_contiguous() and b.is_contiguous()
    assert a.numel() == b.numel()
    out = torch.zeros((2,), dtype=torch.int64, device=a.device)
    _lib.check(_lib.lib().ft8b_selfcheck_divide(_p(a), _p(b), a.numel(), _p(out), _stream()), "ft8b_selfcheck_divide")
    bad, seen = out.cpu().tolist()
    return int(bad), int(seen)


# ------------------------------------------------------------------------------------------------
# Whole path
# ------------------------------------------------------------------------------------------------
@dataclass
class SlotDecodes:
    """Struct-of-arrays result of decode_slots (device tensors unless .cpu() was called)."""
    cand_time: torch.Tensor      # i32 [B,K]
    cand_freq: torch.Tensor      # i32 [B,K]
    cand_score: torch.Tensor     # f64 [B,K]
    n_cand: torch.Tensor         # i32 [B]
    ok: torch.Tensor             # u8  [B,K]
    ldpc_errors: torch.Tensor    # i32 [B,K]
    crc_extracted: torch.Tensor  # i32 [B,K]
    crc_calculated: torch.Tensor  # i32 [B,K]
    payload: torch.Tensor        # u8  [B,K,10]

    def cpu(self) -> "SlotDecodes":
        return SlotDecodes(*[getattr(self, f).cpu() for f in self.__dataclass_fields__])

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.__dataclass_fields__)


def make_params(sample_rate, bins_per_tone, steps_per_symbol, max_candidates, min_score, max_iterations, bin_lo, bin_hi,
                frame_lo, frame_hi, chunk_slots=0, flags=0) -> _lib.DecodeParams:
    return _lib.DecodeParams(float(sample_rate), int(bins_per_tone), int(steps_per_symbol), int(max_candidates),
                             int(max_iterations), float(min_score), int(bin_lo), int(bin_hi), int(frame_lo),
                             int(frame_hi), int(chunk_slots), int(flags))


def alloc_slot_decodes(B: int, K: int, dev, zero: bool = True, pin: bool = False) -> SlotDecodes:
    """Result arrays of a batch; zero=False leaves them uninitialised (for callers that overwrite every element);
    pin=True (host arrays only): page-locked memory from torch's caching host allocator."""
    base = torch.zeros if zero else torch.empty
    mk = (lambda *a, **k: base(*a, pin_memory=True, **k)) if pin else base
    i32, u8 = torch.int32, torch.uint8
    return SlotDecodes(
        cand_time=mk((B, K), dtype=i32, device=dev), cand_freq=mk((B, K), dtype=i32, device=dev),
        cand_score=mk((B, K), dtype=_F64, device=dev), n_cand=mk((B,), dtype=i32, device=dev),
        ok=mk((B, K), dtype=u8, device=dev), ldpc_errors=mk((B, K), dtype=i32, device=dev),
        crc_extracted=mk((B, K), dtype=i32, device=dev), crc_calculated=mk((B, K), dtype=i32, device=dev),
        payload=mk((B, K, 10), dtype=u8, device=dev))


def slice_slot_decodes(o: SlotDecodes, lo: int, hi: int) -> SlotDecodes:
    """Row range [lo, hi) of every array (views: rows are contiguous)."""
    return SlotDecodes(*[getattr(o, f)[lo:hi] for f in o.__dataclass_fields__])


class SlotDecoder:
    """Reusable batched decoder: owns the chunk workspace (and, unless told otherwise, the result
    tensors) for one geometry."""

    def __init__(self, n_slots: int, n_samples: int, params: _lib.DecodeParams, device=None, alloc_out: bool = True):
        self.device = device or require_cuda()
        self.n_slots, self.n_samples, self.params = int(n_slots), int(n_samples), params
        K = params.max_candidates
        if K > _lib.MAX_CANDIDATES:
            raise NotImplementedError("max_candidates=%d exceeds the kernel limit %d" % (K, _lib.MAX_CANDIDATES))
        nbytes = _lib.lib().ft8b_decode_workspace_bytes(C.byref(params), self.n_slots, self.n_samples)
        if nbytes < 0:
            _lib.check(int(nbytes), "ft8b_decode_workspace_bytes")
        self.workspace = torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=self.device)
        self.out = alloc_slot_decodes(self.n_slots, K, self.device) if alloc_out else None

    def run(self, wave: torch.Tensor, out: SlotDecodes = None) -> SlotDecodes:
        """wave [B, N] f32/f64 on the device.  Asynchronous on the current stream."""
        assert wave.is_cuda and wave.dim() == 2 and wave.stride(1) == 1
        assert wave.shape[0] == self.n_slots and wave.shape[1] == self.n_samples
        o = out if out is not None else self.out
        _lib.check(_lib.lib().ft8b_decode_slots(
            _p(wave), wave_dtype_code(wave), self.n_slots, self.n_samples, wave.stride(0), C.byref(self.params),
            _p(self.workspace), self.workspace.numel(), _p(o.cand_time), _p(o.cand_freq), _p(o.cand_score), _p(o.n_cand),
            _p(o.ok), _p(o.ldpc_errors), _p(o.crc_extracted), _p(o.crc_calculated), _p(o.payload), _stream()),
            "ft8b_decode_slots")
        return o


def decode_host_native(waves: torch.Tensor, params: _lib.DecodeParams, piece_slots: int = 256, device=None) -> SlotDecodes:
    """Host [B, N] tensor -> HOST results through ft8b_decode_slots_host: the library's own C++ host path (two streams,
    double-buffered audio, pinned staging for pageable input).  No torch CUDA tensor is involved; torch only provides the
    host arrays here.  This is what decode_slots uses for host input."""
    dev = device or require_cuda()
    assert not waves.is_cuda and waves.dim() == 2 and waves.stride(1) == 1
    B, N = waves.shape
    K = params.max_candidates
    # the library writes every element of all nine arrays; large result sets go into page-locked arrays (cached by torch's
    # host allocator), which the library fills by direct device-to-host copies
    out = alloc_slot_decodes(B, K, torch.device("cpu"), zero=False, pin=B * K * 35 >= (1 << 20))
    if B:
        _lib.check(_lib.lib().ft8b_decode_slots_host(
            _p(waves), wave_dtype_code(waves), B, N, waves.stride(0), C.byref(params), dev.index if dev.index is not None else -1,
            int(piece_slots), _p(out.cand_time), _p(out.cand_freq), _p(out.cand_score), _p(out.n_cand), _p(out.ok),
            _p(out.ldpc_errors), _p(out.crc_extracted), _p(out.crc_calculated), _p(out.payload)), "ft8b_decode_slots_host")
    return out


def decode_host_batch(waves: torch.Tensor, params: _lib.DecodeParams, piece_slots: int = 256, device=None) -> SlotDecodes:
    """Host [B, N] tensor -> device results, with the H2D copy of piece i+1 overlapping the kernels of
    piece i (two device staging buffers, one copy stream).  Pinned host memory copies asynchronously."""
    dev = device or require_cuda()
    B, N = waves.shape
    K = params.max_candidates
    out = alloc_slot_decodes(B, K, dev)
    if B == 0:
        return out
    piece = min(piece_slots, B)
    bufs = [torch.empty((piece, N), dtype=waves.dtype, device=dev) for _ in range(2 if B > piece else 1)]
    decoders = {}
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    done = [None, None]
    copy_stream.wait_stream(main)
    for i, lo in enumerate(range(0, B, piece)):
        hi = min(lo + piece, B)
        n = hi - lo
        buf = bufs[i % len(bufs)]
        with torch.cuda.stream(copy_stream):
            if done[i % len(bufs)] is not None:
                copy_stream.wait_event(done[i % len(bufs)])
            buf[:n].copy_(waves[lo:hi], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        main.wait_event(ready)
        if n not in decoders:
            decoders[n] = SlotDecoder(n, N, params, device=dev, alloc_out=False)
        decoders[n].run(buf[:n], out=slice_slot_decodes(out, lo, hi))
        ev = torch.cuda.Event()
        ev.record(main)
        done[i % len(bufs)] = ev
    return out
