"""WAV ingest with the semantics of the reference's `src/tests/demodulator/from_wave.py`.

`read_wave_file` returns what the reference returns (float32 in [-1, 1], first channel of stereo);
`decode_ft8_from_wave` decodes a file, shipping 16-bit PCM to the GPU as int16 (normalised there with the
same float32 arithmetic) instead of as float32 - SURVEY.md 8(f) row 3.
"""
import wave

import numpy as np

from .ft8_demodulator.ft8_decode import decode_ft8_message, decode_slots

_DTYPES = {1: np.uint8, 2: np.int16, 4: np.int32}


def read_pcm(wave_path: str):
    """(raw integer samples of the first channel, dtype, sample_rate)."""
    with wave.open(wave_path, "rb") as wf:
        n_channels, width, rate = wf.getnchannels(), wf.getsampwidth(), wf.getframerate()
        raw = wf.readframes(wf.getnframes())
    if width not in _DTYPES:
        raise ValueError(f"Unsupported sample width: {width}")          # from_wave.py:55-56
    data = np.frombuffer(raw, dtype=_DTYPES[width])
    if n_channels == 2:                                                  # from_wave.py:60-62
        data = data[::2]
    return data, _DTYPES[width], rate


def read_wave_file(wave_path: str):
    """(float32 audio normalised by iinfo(dtype).max, sample_rate) - from_wave.py:24-69."""
    data, dtype, rate = read_pcm(wave_path)
    audio = data.astype(np.float32)
    audio /= np.iinfo(dtype).max
    return audio, rate


def decode_ft8_from_wave(wave_path: str, freq_min=None, freq_max=None, time_min=None, time_max=None, bins_per_tone=2,
                         steps_per_symbol=2, max_candidates=20, min_score=10, max_iterations=20):
    """Decode one WAV file; returns decode_ft8_message's list."""
    data, dtype, rate = read_pcm(wave_path)
    kw = dict(bins_per_tone=bins_per_tone, steps_per_symbol=steps_per_symbol, max_candidates=max_candidates,
              min_score=min_score, max_iterations=max_iterations, freq_min=freq_min, freq_max=freq_max,
              time_min=time_min, time_max=time_max)
    if dtype == np.int16:
        return decode_slots(np.ascontiguousarray(data).reshape(1, -1), rate, pcm16=True, **kw)[0]
    audio = data.astype(np.float32)
    audio /= np.iinfo(dtype).max
    return decode_ft8_message(audio, rate, **kw)
