"""ctypes binding of libft8b200.so (the C ABI declared in include/ft8b200.h).

There is NO fallback: if the shared library is missing, or a compute entry point is called
without a CUDA device, the call raises.  Build with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C ft8_demodulator_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FT8B_LIB_PATH") or os.path.join(_HERE, "libft8b200.so")   # override: kernel A/B experiments

ERR_ARG, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3
DTYPE_F32, DTYPE_F64, DTYPE_I16, DTYPE_C64, DTYPE_C128 = 0, 1, 2, 3, 4
DRIFT_AUTO, DRIFT_WALK, DRIFT_BLOCKSUM, DRIFT_DIRECT = 0, 1, 2, 3
MAX_CANDIDATES = 1048576          # FT8B_MAX_CANDIDATES (strict top-K: 4096)
FLAG_STRICT_TOPK = 1
LDPC_N = 174


class DecodeParams(C.Structure):
    """ft8b_decode_params (include/ft8b200.h)."""
    _fields_ = [
        ("sample_rate", C.c_double), ("bins_per_tone", C.c_int32), ("steps_per_symbol", C.c_int32),
        ("max_candidates", C.c_int32), ("max_iterations", C.c_int32), ("min_score", C.c_double),
        ("bin_lo", C.c_int32), ("bin_hi", C.c_int32), ("frame_lo", C.c_int64), ("frame_hi", C.c_int64),
        ("chunk_slots", C.c_int32), ("flags", C.c_int32),
    ]


# every symbol include/ft8b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_ip = C.POINTER(C.c_int)
SIGNATURES = {
    "ft8b_version": (_i, []),
    "ft8b_last_error": (C.c_char_p, []),
    "ft8b_release": (_i, []),
    "ft8b_launch_count": (_i64, []),
    "ft8b_stage_timing": (_i, [_i]),
    "ft8b_stage_times": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ft8b_plan": (_i, [_d, _i, _i, _ip, _ip, _ip]),
    "ft8b_num_frames": (_i64, [_i64, _i, _i]),
    "ft8b_search_ranges": (_i, [_i, _i, _i, _i, _ip, _ip, _ip]),
    "ft8b_stft_db": (_i, [_vp, _i, _i64, _i64, _i64, _i, _i, _i, _i, _i, _i64, _i64, _vp, _vp]),
    "ft8b_sync_scores": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "ft8b_select_candidates": (_i, [_vp, _i64, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ft8b_find_candidates": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ft8b_sync_score_at": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "ft8b_select_topk": (_i, [_vp, _i64, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp]),
    "ft8b_decode_candidates": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp]),
    "ft8b_ldpc_crc": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ft8b_extract_llr": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ft8b_normalize_llr": (_i, [_vp, _i64, _vp]),
    "ft8b_ldpc_check": (_i, [_vp, _i64, _vp, _vp]),
    "ft8b_selfcheck_divide": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "ft8b_column_argmax": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp]),
    "ft8b_dechirp": (_i, [_vp, _i, _i64, _d, _d, _vp, _vp]),
    "ft8b_sync_scores_drift": (_i, [_vp, _i64, _i, _i, _i, _i, _i, C.POINTER(C.c_double), _vp, _vp, _vp]),
    "ft8b_sync_scores_drift_variant": (_i, [_vp, _i64, _i, _i, _i, _i, _i, C.POINTER(C.c_double), _vp, _vp, _i, _vp]),
    "ft8b_decode_candidates_drift": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                          _vp]),
    "ft8b_decode_workspace_bytes": (_i64, [C.POINTER(DecodeParams), _i64, _i64]),
    "ft8b_decode_slots_host": (_i, [_vp, _i, _i64, _i64, _i64, C.POINTER(DecodeParams), _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _vp, _vp]),
    "ft8b_last_host_call_ms": (_d, []),
    "ft8b_set_host_staging": (_i, [_i, _i]),
    "ft8b_selftest_host_staging": (_i, [_vp, _vp, _i64, _i64, _i64]),
    "ft8b_decode_slots": (_i, [_vp, _i, _i64, _i64, _i64, C.POINTER(DecodeParams), _vp, _i64, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Load the library (once).  Raises ImportError with build instructions if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libft8b200.so not found at %s - build it with `make -C %s` (nvcc, sm_100a). "
                "ft8_demodulator_b200 has no CPU fallback." % (LIB_PATH, os.path.join(_HERE, "csrc")))
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().ft8b_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """Translate a C-ABI status into the Python exception the mirror API documents."""
    if rc == 0:
        return
    msg = "%s%s" % ((what + ": ") if what else "", last_error())
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NO_DEVICE:
        raise RuntimeError("no CUDA device: " + msg)
    raise RuntimeError("CUDA error %d: %s" % (rc, msg))


def plan(sample_rate, bins_per_tone, steps_per_symbol):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().ft8b_plan(float(sample_rate), int(bins_per_tone), int(steps_per_symbol), C.byref(a), C.byref(b),
                          C.byref(c)), "ft8b_plan")
    return a.value, b.value, c.value


def num_frames(n_samples, nperseg, hop):
    return int(lib().ft8b_num_frames(int(n_samples), int(nperseg), int(hop)))


def search_ranges(T, F, time_osr, freq_osr):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().ft8b_search_ranges(int(T), int(F), int(time_osr), int(freq_osr), C.byref(a), C.byref(b), C.byref(c)),
          "ft8b_search_ranges")
    return a.value, b.value, c.value
