"""Result and container types of the FT8 demodulator.

Same names, fields and validation as the reference's `ft8_tools/ft8_demodulator/ftx_types.py:9-60`
so code written against it keeps working; the arithmetic lives in the CUDA library.
"""
from dataclasses import dataclass, field
from enum import Enum

import numpy as np


class FT8Protocol(Enum):
    FT8 = 1


@dataclass
class FT8Waterfall:
    """dB spectrogram with shape (frequency, time) plus its oversampling factors (ftx_types.py:14-34)."""
    mag: np.ndarray
    time_osr: int
    freq_osr: int

    def __post_init__(self):
        if len(self.mag.shape) != 2:
            raise ValueError("mag must be a 2D array with shape (frequency, time)")

    @property
    def num_bins(self) -> int:
        return self.mag.shape[0]

    @property
    def num_blocks(self) -> int:
        return self.mag.shape[1] // self.time_osr


@dataclass
class FT8Candidate:
    """A (time, frequency) sync hypothesis on a waterfall (ftx_types.py:37-47)."""
    waterfall: "FT8Waterfall"
    abs_time: int = 0
    abs_freq: int = 0
    score: float = 0.0

    def get_log_power(self, time_offset: int, freq_offset: int):
        wf = self.waterfall
        return wf.mag[self.abs_freq + freq_offset * wf.freq_osr, self.abs_time + time_offset * wf.time_osr]


@dataclass
class FT8Message:
    payload: bytearray = field(default_factory=lambda: bytearray(10))
    hash: int = 0


@dataclass
class FT8DecodeStatus:
    ldpc_errors: int = 0
    crc_extracted: int = 0
    crc_calculated: int = 0
