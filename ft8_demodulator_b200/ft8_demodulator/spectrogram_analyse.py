"""STFT spectrogram - the reference's `spectrogram_analyse.py` API on the GPU.

`calculate_spectrogram` has the signature and return layout of
`ft8_tools/ft8_demodulator/spectrogram_analyse.py:19-66` (two-sided, fftshifted, dB, shape
[nfft, frames]); the transform runs in csrc/stft.cu in float64.
"""
import numpy as np

from .. import _lib, engine

FT8_SYMBOL_DURATION_S = 0.16 / 1
FT8_SYMBOL_FREQ_INTERVAL_HZ = 6.25 * 1
FT8_BAUD_RATE = 1 / FT8_SYMBOL_DURATION_S
SPECTROGRAM_BINS_PER_TONE = 10
SPECTROGRAM_STEPS_PER_SYMBOL = 10
FT8_NUM_SYNC_SEQUENCE = 3
FT8_NUM_SYNC_SYMBOLS_PER_SEQUENCE = 7
FT8_SYNC_PATTERN = [3, 1, 4, 0, 6, 5, 2]
FT8_SYNC_SEQUENCE_OFFSET = 36


def frequency_axis(nfft: int, sample_rate) -> np.ndarray:
    """fftshift(fftfreq(nfft, 1/fs)) with NumPy's own rounding (k * (1/(n*d)))."""
    return np.fft.fftshift(np.fft.fftfreq(nfft, 1 / sample_rate))


def time_axis(n_samples: int, nperseg: int, hop: int, sample_rate) -> np.ndarray:
    """Segment-centre times scipy.signal.spectrogram reports."""
    return np.arange(nperseg / 2, n_samples - nperseg / 2 + 1, hop) / float(sample_rate)


def _as_real_audio(wave_data) -> np.ndarray:
    x = np.asarray(wave_data)
    if np.iscomplexobj(x):
        raise NotImplementedError("the decode path takes real audio (the reference's callers pass np.real(...)); "
                                  "complex input is supported by calculate_spectrogram only")
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    return np.ascontiguousarray(x.reshape(-1))


def calculate_spectrogram(wave_data: np.ndarray, sample_rate: int, bins_per_tone: int = 2,
                          steps_per_symbol: int = 2) -> tuple:
    """dB spectrogram of `wave_data`: (spectrogram[nfft, frames], f[nfft], t[frames])."""
    nperseg, hop, nfft = _lib.plan(sample_rate, bins_per_tone, steps_per_symbol)
    if np.iscomplexobj(wave_data):
        return _calculate_spectrogram_complex(np.asarray(wave_data).reshape(-1), sample_rate, nperseg, hop, nfft)
    x = _as_real_audio(wave_data)
    if len(x) < nperseg:                               # spectrogram_analyse.py:37-39
        return np.array([[]]), np.array([]), np.array([])
    T = _lib.num_frames(len(x), nperseg, hop)
    nbin = nfft // 2 + 1 if nfft % 2 == 0 else (nfft + 1) // 2
    dev = engine.require_cuda()
    mag = engine.stft_db(engine._dev(x.reshape(1, -1), device=dev), nperseg, hop, nfft, 0, nbin, 0, T)
    pos = mag[0].transpose(0, 1).cpu().numpy()         # [nbin, T]
    # two-sided, fftshifted rows: -fs/2 .. -df, 0 .. ; real input => S[-k] = S[k]
    if nfft % 2 == 0:
        neg = pos[nfft // 2:0:-1]                      # bins nfft/2 (Nyquist, sits at -fs/2), nfft/2-1, ..., 1
        spec = np.concatenate([neg, pos[:nfft // 2]], axis=0)
    else:
        neg = pos[(nfft - 1) // 2:0:-1]
        spec = np.concatenate([neg, pos], axis=0)
    if x.dtype == np.float32:
        spec = spec.astype(np.float32)
    return np.ascontiguousarray(spec), frequency_axis(nfft, sample_rate), time_axis(len(x), nperseg, hop, sample_rate)


def _calculate_spectrogram_complex(x: np.ndarray, sample_rate, nperseg: int, hop: int, nfft: int) -> tuple:
    """Complex audio (the satellite pre-processor's input, frequency_correction.py:190-192): full two-sided
    transform of length nfft on the GPU, fftshifted on the host like spectrogram_analyse.py:63-64."""
    if len(x) < nperseg:
        return np.array([[]]), np.array([]), np.array([])
    if x.dtype != np.complex64:
        x = x.astype(np.complex128)
    T = _lib.num_frames(len(x), nperseg, hop)
    dev = engine.require_cuda()
    mag = engine.stft_db(engine._dev(np.ascontiguousarray(x).reshape(1, -1), device=dev), nperseg, hop, nfft, 0, nfft, 0, T)
    spec = np.fft.fftshift(mag[0].transpose(0, 1).cpu().numpy(), axes=0)
    if x.dtype == np.complex64:
        spec = spec.astype(np.float32)
    return np.ascontiguousarray(spec), frequency_axis(nfft, sample_rate), time_axis(len(x), nperseg, hop, sample_rate)


def select_frequency_band(spectrogram: np.ndarray, f: np.ndarray, f_min: float, f_max: float) -> tuple:
    """Rows with f_min <= f <= f_max (spectrogram_analyse.py:68-82)."""
    mask = (f >= f_min) & (f <= f_max)
    return spectrogram[mask], f[mask]
