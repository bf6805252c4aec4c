"""Containers the FT8 demodulator API hands around.

They carry the names, fields, defaults and validation callers of the reference rely on
(`ft8_tools/ft8_demodulator/ftx_types.py:9-60`): a waterfall is a [frequency, time] dB array plus its two
oversampling factors, a candidate is a (time, frequency) position on a waterfall, a decode yields a 10-byte
payload with its CRC and a status record.  No arithmetic lives here - that is the CUDA library's job.
"""
import enum
from dataclasses import dataclass, field

import numpy as np

PAYLOAD_BYTES = 10


class FT8Protocol(enum.Enum):
    FT8 = 1


def _require_2d(mag) -> None:
    if np.ndim(mag) != 2:
        raise ValueError("mag must be a 2D array with shape (frequency, time)")      # same message as ftx_types.py:24


@dataclass
class FT8Waterfall:
    """ftx_types.py:14-34.  `mag[f, t]` in dB; `time_osr` columns per symbol, `freq_osr` rows per tone."""
    mag: np.ndarray
    time_osr: int
    freq_osr: int

    def __post_init__(self):
        _require_2d(self.mag)

    # rows of the waterfall / whole symbols it spans
    num_bins = property(lambda self: self.mag.shape[0])
    num_blocks = property(lambda self: self.mag.shape[1] // self.time_osr)


@dataclass
class FT8Candidate:
    """ftx_types.py:37-47: a sync hypothesis; offsets passed to get_log_power are in symbols / tones."""
    waterfall: FT8Waterfall
    abs_time: int = 0
    abs_freq: int = 0
    score: float = 0.0

    def get_log_power(self, time_offset: int, freq_offset: int):
        grid = self.waterfall
        row = self.abs_freq + grid.freq_osr * freq_offset
        col = self.abs_time + grid.time_osr * time_offset
        return grid.mag[row, col]


@dataclass
class FT8Message:
    """ftx_types.py:49-53: payload (77 message bits, byte 9 masked with 0xF8) and its CRC-14 as `hash`."""
    payload: bytearray = field(default_factory=lambda: bytearray(PAYLOAD_BYTES))
    hash: int = 0


@dataclass
class FT8DecodeStatus:
    """ftx_types.py:55-60: what the LDPC / CRC stage reports for one candidate."""
    ldpc_errors: int = 0
    crc_extracted: int = 0
    crc_calculated: int = 0
