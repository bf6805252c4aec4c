"""The C-ABI library loads and exports every symbol include/ft8b200.h declares; host-only entry
points agree with the oracle; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from ft8_demodulator_b200 import _lib


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ft8b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ft8b_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libft8b200.so does not export " + n
    # and the ctypes binding covers exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_slot():
    L = _lib.lib()
    assert L.ft8b_version() == 100
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    assert L.ft8b_plan(-1.0, 2, 2, C.byref(a), C.byref(b), C.byref(c)) == _lib.ERR_ARG
    assert "ft8b_plan" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.plan(12000, 0, 2)


@pytest.mark.parametrize("fs", [12000, 6000, 10e3, 3500, 20000, 2000, 32768, 4500, 9999.5, 1_000_000])
@pytest.mark.parametrize("b,s", [(2, 2), (4, 4), (10, 10), (1, 1), (2, 8), (3, 5000)])
def test_geometry_matches_oracle(orc, fs, b, s):
    assert _lib.plan(fs, b, s) == orc.plan(fs, b, s)
    nperseg, hop, nfft = _lib.plan(fs, b, s)
    for n in (0, 10, nperseg - 1, nperseg, nperseg + hop, 180000):
        assert _lib.num_frames(n, nperseg, hop) == orc.lib().orc_num_frames(n, nperseg, hop)


@pytest.mark.parametrize("F,T,s,b", [(1920, 186, 2, 2), (100, 50, 2, 2), (64, 400, 4, 1), (14, 300, 1, 2), (961, 186, 2, 2)])
def test_search_ranges_match_oracle(orc, F, T, s, b):
    t0, nt, nf = _lib.search_ranges(T, F, s, b)
    o = orc.search_ranges(F, T, s, b)
    assert (t0, t0 + nt, 0, nf) == o


def test_workspace_query_is_host_only():
    from ft8_demodulator_b200 import engine
    p = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 1920, 0, 186, 16)
    n = _lib.lib().ft8b_decode_workspace_bytes(C.byref(p), 64, 180000)
    assert n >= 16 * 8 * (186 * 1920 + 88 * 1906) + 64 * 20 * 174 * 8
    bad = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 5000, 0, 186, 16)
    assert _lib.lib().ft8b_decode_workspace_bytes(C.byref(bad), 64, 180000) < 0


@pytest.mark.parametrize("chunk", [0, 1, 7, 16, 64, 100, 2045])
def test_workspace_bytes_bound_every_shorter_batch(chunk):
    """ft8b_decode_workspace_bytes(n) is an upper bound for every batch of at most n slots (the host path sizes its
    workspace once, for a full piece, and then decodes a shorter tail piece with it: with chunk_slots = 100 a 128-slot
    piece runs as 2 x 64 but a 72..100-slot tail as ONE chunk).  Host-only query, so it is checked here."""
    from ft8_demodulator_b200 import engine
    p = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 1920, 0, 186, chunk)
    q = lambda n: int(_lib.lib().ft8b_decode_workspace_bytes(C.byref(p), n, 180000))
    sizes = [q(n) for n in range(0, 300)] + [q(3000), q(4096)]
    assert all(b >= a for a, b in zip(sizes, sizes[1:])), "not monotone in n_slots"
    per_slot = 8 * (186 * 1920 + 88 * 1906)
    lim = chunk if chunk else 2045
    for n in (1, 64, 72, 100, 128, 299, 3000, 4096):
        # room for the largest chunk a batch of <= n slots can be cut into, plus the LLRs of all its candidates
        assert q(n) >= per_slot * min(lim, n) + 8 * 174 * 20 * n


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd, ldpc_decoder as ld, spectrogram_analyse as sa
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fd.decode_ft8_message(np.zeros(180000), 12000)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ld.bp_decode(np.ones(174), 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sa.calculate_spectrogram(np.zeros(180000), 12000)
    # and the raw ABI reports a CUDA/no-device error instead of computing anything
    rc = _lib.lib().ft8b_ldpc_check(1, 1, 1, None)
    assert rc != 0
    # the host-memory entry point: valid host arrays, no device -> FT8B_ERR_NO_DEVICE and untouched outputs
    from ft8_demodulator_b200 import engine
    x = np.zeros((2, 180000), dtype=np.float32)
    prm = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 1920, 0, 186)
    outs = [np.full((2, 20), 7, np.int32), np.full((2, 20), 7, np.int32), np.full((2, 20), 7.0), np.full(2, 7, np.int32),
            np.full((2, 20), 7, np.uint8), np.full((2, 20), 7, np.int32), np.full((2, 20), 7, np.int32),
            np.full((2, 20), 7, np.int32), np.full((2, 20, 10), 7, np.uint8)]
    rc = _lib.lib().ft8b_decode_slots_host(x.ctypes.data, _lib.DTYPE_F32, 2, 180000, 180000, C.byref(prm), -1, 0,
                                           *[a.ctypes.data for a in outs])
    assert rc == _lib.ERR_NO_DEVICE and "no CPU fallback" in _lib.last_error()
    assert all((a == 7).all() for a in outs)


def test_host_entry_argument_checks():
    """ft8b_decode_slots_host rejects malformed calls before touching any device (runs without a GPU)."""
    from ft8_demodulator_b200 import engine
    prm = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 1920, 0, 186)
    x = np.zeros((1, 180000), dtype=np.float32)
    o = [np.zeros(64 * 10, np.float64) for _ in range(9)]
    ptrs = [a.ctypes.data for a in o]
    L = _lib.lib()
    assert L.ft8b_decode_slots_host(x.ctypes.data, 9, 1, 180000, 180000, C.byref(prm), -1, 0, *ptrs) == _lib.ERR_ARG       # dtype
    assert L.ft8b_decode_slots_host(x.ctypes.data, _lib.DTYPE_F32, 1, 180000, 100, C.byref(prm), -1, 0, *ptrs) == _lib.ERR_ARG  # stride < n
    assert L.ft8b_decode_slots_host(None, _lib.DTYPE_F32, 1, 180000, 180000, C.byref(prm), -1, 0, *ptrs) == _lib.ERR_ARG
    assert L.ft8b_decode_slots_host(x.ctypes.data, _lib.DTYPE_F32, 1, 180000, 180000, C.byref(prm), -1, 0, *(ptrs[:8] + [None])) == _lib.ERR_ARG
    big = engine.make_params(12000, 2, 2, 2_000_000, 1, 20, 0, 1920, 0, 186)      # above FT8B_MAX_CANDIDATES (2^20)
    assert L.ft8b_decode_slots_host(x.ctypes.data, _lib.DTYPE_F32, 1, 180000, 180000, C.byref(big), -1, 0, *ptrs) == _lib.ERR_UNSUPPORTED
    assert L.ft8b_decode_slots_host(x.ctypes.data, _lib.DTYPE_F32, 0, 180000, 180000, C.byref(prm), -1, 0, *ptrs) == 0       # empty batch


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ft8_demodulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "ft8_oracle" not in src and "reference_loader" not in src, f


@pytest.mark.parametrize("threads,nt", [(1, 1), (3, 1), (4, 0), (0, 1)])
def test_host_staging_gather(threads, nt):
    """The pinned-staging gather of ft8b_decode_slots_host (persistent host threads, streaming stores) is host-only
    code: strided and misaligned sources, rows larger and smaller than a 256 KB block, empty input - exact copies,
    nothing written outside the destination."""
    import numpy as np
    L = _lib.lib()
    assert L.ft8b_set_host_staging(threads, nt) >= 1
    rng = np.random.default_rng(threads)
    for rows, rb, pitch, off in ((40, 720000, 720000, 0), (33, 300001, 300013, 4), (200, 50003, 50011, 3),
                                 (1, 9000001, 9000001, 1), (5, 17, 40, 0), (0, 10, 10, 0)):
        src = rng.integers(0, 255, rows * pitch + 64, dtype=np.uint8)
        dst = np.zeros(rows * rb + 64, dtype=np.uint8)
        assert L.ft8b_selftest_host_staging(dst.ctypes.data + off, src.ctypes.data + off, rows, rb, pitch) == 0
        want = src[off:off + rows * pitch].reshape(rows, pitch)[:, :rb].reshape(-1) if rows else src[:0]
        assert np.array_equal(dst[off:off + rows * rb], want), (rows, rb, pitch)
        assert not dst[:off].any() and not dst[off + rows * rb:].any()
    assert L.ft8b_selftest_host_staging(None, None, 1, 1, 1) == _lib.ERR_ARG
    L.ft8b_set_host_staging(0, 1)
