"""(174,91) LDPC decoding - the reference's `ldpc_decoder.py` API on the GPU.

`bp_decode` / `ldpc_check` keep the signatures of `ft8_tools/ft8_demodulator/ldpc_decoder.py:33-113`;
the sum-product iterations run in csrc/decode.cu (one warp per codeword, float64, the reference's
rational tanh/atanh and operation order, so results are bit-identical to the NumPy implementation).
`bp_decode_batch` is the batched form used for throughput.
"""
from typing import Tuple

import numpy as np

from .. import engine
from .constants import FTX_LDPC_M, FTX_LDPC_N  # noqa: F401  (re-exported like the reference)


def ldpc_check(codeword: np.ndarray) -> int:
    """Number of unsatisfied parity checks of a 174-bit word (ldpc_decoder.py:33-52)."""
    dev = engine.require_cuda()
    bits = (np.asarray(codeword).reshape(-1)[:FTX_LDPC_N] != 0).astype(np.uint8)
    if bits.shape[0] != FTX_LDPC_N:
        raise IndexError("codeword must hold %d bits" % FTX_LDPC_N)
    return int(engine.ldpc_check(engine._dev(bits.reshape(1, -1), device=dev)).cpu()[0])


def bp_decode_batch(llr, max_iterations: int):
    """Decode [n,174] LLR vectors.  Returns (plain uint8 [n,174], min_errors int32 [n]) as NumPy arrays."""
    dev = engine.require_cuda()
    a = np.ascontiguousarray(np.asarray(llr, dtype=np.float64).reshape(-1, FTX_LDPC_N))
    out = engine.ldpc_crc(engine._dev(a, device=dev), int(max_iterations))
    return out.plain.cpu().numpy(), out.ldpc_errors.cpu().numpy()


def bp_decode(codeword: np.ndarray, max_iterations: int) -> Tuple[np.ndarray, int]:
    """Belief-propagation decode of 174 LLRs -> (hard decisions uint8[174], min parity errors) (ldpc_decoder.py:54-113)."""
    plain, err = bp_decode_batch(np.asarray(codeword, dtype=np.float64).reshape(1, -1), max_iterations)
    return plain[0], int(err[0])


def fast_tanh(x):
    """Rational tanh approximation of ldpc_decoder.py:11-21 (host utility; the kernel has its own copy)."""
    x = np.clip(x, -4.97, 4.97)
    x2 = x * x
    return (x * (945.0 + x2 * (105.0 + x2))) / (945.0 + x2 * (420.0 + x2 * 15.0))


def fast_atanh(x):
    """Rational atanh approximation of ldpc_decoder.py:23-31 (host utility)."""
    x2 = x * x
    return (x * (945.0 + x2 * (-735.0 + x2 * 64.0))) / (945.0 + x2 * (-1050.0 + x2 * 225.0))


__all__ = ["ldpc_check", "bp_decode", "bp_decode_batch", "fast_tanh", "fast_atanh"]
