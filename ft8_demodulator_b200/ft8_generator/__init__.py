"""Reference-compatible names for the FT8 transmit chain (`ft8_tools.ft8_generator`), backed by the vectorised
generator in `ft8_demodulator_b200.synth`.  Same arguments and return types as
ft8_generator/{crc,ldpc,encoder,modulator}.py; waveforms are bit-identical to the reference's
(tests/test_oracle_vs_reference.py) and ~50x faster to produce.  Test-vector source, not part of the decode path.
"""
import numpy as np

from .. import synth

FT8_SYMBOL_NUM = synth.FT8_SYMBOL_NUM


def crc_generator(payload_10bytes) -> np.ndarray:
    """payload (10 bytes) -> a91 (12 bytes) with CRC-14 (crc.py:25-47)."""
    return synth.add_crc(payload_10bytes)


def get_crc_from_a91(a91_12bytes) -> np.uint16:
    """The 14 CRC bits of a91 (crc.py:49-51)."""
    a = np.asarray(a91_12bytes, dtype=np.uint8)
    return np.uint16(((int(a[9]) & 0x07) << 11) | (int(a[10]) << 3) | (int(a[11]) >> 5))


def ldpc_generator(a91_12bytes) -> np.ndarray:
    """a91 (12 bytes) -> 174-bit codeword in 22 bytes (ldpc.py:104-131)."""
    return synth.ldpc_encode(a91_12bytes)


def symbolIdSequence_generator(codeword) -> np.ndarray:
    """174 codeword bits -> 58 Gray-mapped 3-bit symbols (encoder.py:15-39)."""
    bits = np.unpackbits(np.asarray(codeword, dtype=np.uint8))[:174].reshape(58, 3)
    return synth.FT8_GRAY_MAP[(bits[:, 0] << 2) | (bits[:, 1] << 1) | bits[:, 2]].astype(np.uint8)


def itones_generator(symbolIdSequence) -> np.ndarray:
    """58 data symbols -> 79 tones S7 D29 S7 D29 S7 (encoder.py:41-62)."""
    d = np.asarray(symbolIdSequence, dtype=np.uint8)
    c = synth.FT8_COSTAS_PATTERN
    return np.concatenate([c, d[:29], c, d[29:], c]).astype(np.uint8)


def ft8_encode(payload) -> np.ndarray:
    """payload -> 79 tones (encoder.py:64-72)."""
    return synth.ft8_encode(payload)


def gfsk_modulation_waveform_generator(itones, fs: float) -> np.ndarray:
    """Instantaneous frequency track, (79+2)*sps samples (modulator.py:27-48)."""
    return synth.gfsk_frequency_track(itones, fs)


def ft8_modulation_waveform_generator(gfsk_waveform, fs: float, f0: float) -> np.ndarray:
    """Phase-accumulated complex burst from a frequency track (modulator.py:56-75)."""
    return synth.baseband_from_track(gfsk_waveform, fs, f0)


def ft8_baseband_generator(payload, fs: float, f0: float) -> np.ndarray:
    """Complex baseband burst (modulator.py:77-83)."""
    return synth.ft8_baseband(payload, fs, f0)


def ft8_generator(payload, fs: float, f0: float, fc: float) -> np.ndarray:
    """Real passband burst (modulator.py:85-90)."""
    return synth.ft8_generator(payload, fs, f0, fc)


__all__ = ['crc_generator', 'get_crc_from_a91', 'symbolIdSequence_generator', 'itones_generator', 'ft8_encode',
           'gfsk_modulation_waveform_generator', 'ft8_modulation_waveform_generator', 'ft8_baseband_generator',
           'ft8_generator', 'ldpc_generator']
