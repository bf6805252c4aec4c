"""Satellite pre-processor: estimate and remove the Doppler drift of a single dominant FT8 signal.

Mirror of the reference's `ft8_tools/ft8_beacon_receiver/frequency_correction.py`: same
function names, arguments, defaults and return values.  The data-parallel steps run on the GPU -
the complex STFT (csrc/stft.cu), the per-column ridge arg-max and the two de-chirp multiplies
(csrc/drift.cu) - so the spectrogram never leaves HBM; the regression and template-correlation logic works on a
few hundred ridge points and stays on the host (closed-form least squares instead of scikit-learn, which the
reference imports but does not list in requirements.txt).  No plots, no logging handlers.

Parity for this row is at payload level: the estimated drift agrees with the
reference to ~1e-9 relative and the corrected waveform decodes to the same payloads.
"""
import numpy as np
import torch
from scipy.special import erf

from .. import _lib, engine
from ..ft8_demodulator.spectrogram_analyse import time_axis  # noqa: F401  (API parity helper)

DEFAULT_PARAMS = {                       # frequency_correction.py:146-160
    'nsync_sym': 7,
    'ndata_sym': 58,
    'zscore_threshold': 5,
    'max_iteration_num': 400,
    'debug_plots': True,                 # accepted and ignored: nothing is plotted here
    'window_size_factor': 4,
    'max_variance_factor': 0.0001,
    'fit_middle_percent': 100,
    'bins_per_tone': 2,
    'steps_per_symbol': 2,
    'poly_degree': 2,
    'precise_sync': True,
}


def gfsk_pulse(bt, t):
    """GFSK frequency pulse (frequency_correction.py:27-40)."""
    k = np.pi * np.sqrt(2.0 / np.log(2.0))
    return 0.5 * (erf(k * bt * (t + 0.5)) - erf(k * bt * (t - 0.5)))


def detect_signal_continuity(max_freq_indices, window_size=8, max_variance=10.0):
    """Segments where the ridge is locally linear (frequency_correction.py:42-115): for every window of
    `window_size` ridge points fit a line and take the population variance of the residuals; windows with
    variance < max_variance are 'signal'.  Returns (segments [(start, end)], continuity_metric = -variance)."""
    y = np.asarray(max_freq_indices, dtype=np.float64)
    n = len(y)
    if n < window_size:
        return [], np.zeros(n)
    w = int(window_size)
    win = np.lib.stride_tricks.sliding_window_view(y, w)              # [n-w+1, w]
    x = np.arange(w, dtype=np.float64)
    xc = x - x.mean()
    yc = win - win.mean(axis=1, keepdims=True)
    slope = (yc @ xc) / (xc @ xc)
    resid = yc - slope[:, None] * xc[None, :]
    continuity_metric = -np.mean(resid * resid, axis=1)               # -np.var(residuals)
    is_signal = continuity_metric > -max_variance
    segments, in_segment, start = [], False, 0
    for i in range(len(is_signal)):                                   # :96-107
        if is_signal[i] and not in_segment:
            in_segment, start = True, i
        elif not is_signal[i] and in_segment:
            in_segment = False
            if i - start >= 1:
                segments.append((start, i))
    if in_segment:
        segments.append((start, n - 1))                               # :110-111
    return segments, continuity_metric


def _polyfit_coefs(x, y, degree):
    """Least-squares polynomial, returned like sklearn's PolynomialFeatures + LinearRegression:
    (intercept, coefs) with coefs[k] the coefficient of x**k (coefs[0] is the dummy bias column = 0)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    xm, ym = x.mean(), y.mean()                                       # LinearRegression centres X and y
    V = np.stack([x ** k for k in range(1, degree + 1)], axis=1)
    Vm = V.mean(axis=0)
    sol, *_ = np.linalg.lstsq(V - Vm, y - ym, rcond=None)
    coefs = np.concatenate([[0.0], sol])
    intercept = ym - Vm @ sol
    del xm
    return float(intercept), coefs


def _ridge(wave_dev: torch.Tensor, fs, bins_per_tone, steps_per_symbol):
    """Ridge of the non-negative-frequency half of the spectrogram: arg-max row per column
    (frequency_correction.py:190-224).  Everything stays on the device; only T int32 indices come back."""
    nperseg, hop, nfft = _lib.plan(fs, bins_per_tone, steps_per_symbol)
    n = wave_dev.shape[-1]
    T = _lib.num_frames(n, nperseg, hop)
    if T <= 0:
        raise IndexError("input shorter than one analysis window")
    mag = engine.stft_db(wave_dev.reshape(1, -1), nperseg, hop, nfft, 0, nfft, 0, T)       # [1, T, nfft] natural order
    n_pos = (nfft - 1) // 2 + 1                                        # rows with f >= 0 (f = fftshift(fftfreq))
    idx = torch.zeros((T,), dtype=torch.int32, device=wave_dev.device)
    _lib.check(_lib.lib().ft8b_column_argmax(mag.data_ptr(), T, nfft, 0, n_pos, idx.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "ft8b_column_argmax")
    return idx.cpu().numpy().astype(np.int64), n_pos


def _dechirp(wave_dev: torch.Tensor, c2: float, c3: float) -> torch.Tensor:
    out = torch.empty(wave_dev.shape, dtype=torch.complex128, device=wave_dev.device)
    _lib.check(_lib.lib().ft8b_dechirp(wave_dev.data_ptr(), engine.wave_dtype_code(wave_dev), wave_dev.numel(), float(c2),
                                       float(c3), out.data_ptr(), torch.cuda.current_stream().cuda_stream), "ft8b_dechirp")
    return out


def correct_frequency_drift(wave_complex: np.ndarray, fs: float, sym_bin: float, sym_t: float, params=None, *,
                            trace: dict = None):
    """Remove linear (and optionally quadratic) frequency drift from a complex FT8 capture.

    Same contract as frequency_correction.py:118-659: returns (corrected complex signal at the input rate,
    estimated drift per sample = Hz/s / fs).  `params` is completed in place with the defaults, as the reference does.
    `trace` (extension, keyword only): a dict that receives the intermediate ridge / segment / fit values."""
    if params is None:
        params = dict(DEFAULT_PARAMS)
    else:
        for key, value in DEFAULT_PARAMS.items():
            if key not in params:
                params[key] = value
    bins_per_tone, steps_per_symbol = params['bins_per_tone'], params['steps_per_symbol']
    nsync_sym, ndata_sym = params['nsync_sym'], params['ndata_sym']
    window_size = params['window_size_factor'] * steps_per_symbol
    time_osr, freq_osr = steps_per_symbol, bins_per_tone

    dev = engine.require_cuda()
    wave_host = np.ascontiguousarray(np.asarray(wave_complex).reshape(-1))
    if not np.iscomplexobj(wave_host):
        wave_host = wave_host.astype(np.complex128)
    elif wave_host.dtype != np.complex64:
        wave_host = wave_host.astype(np.complex128)
    nsamples = len(wave_host)
    wave_dev = engine._dev(wave_host, device=dev)

    # ---- ridge of the spectrogram and its locally-linear segments                                    :190-243
    max_freq_indices, freq_bins = _ridge(wave_dev, fs, bins_per_tone, steps_per_symbol)
    max_variance = params['max_variance_factor'] * (freq_bins ** 2)
    signal_segments, _ = detect_signal_continuity(max_freq_indices, window_size=window_size, max_variance=max_variance)
    if trace is not None:
        trace.update(ridge1=max_freq_indices, segments=list(signal_segments))
    if not signal_segments:
        return wave_complex, 0.0
    start_idx, end_idx = max(signal_segments, key=lambda s: s[1] - s[0])
    freq_step = sym_bin / freq_osr
    time_step = sym_t / time_osr
    max_freqs = max_freq_indices * freq_step
    t_axis = np.arange(len(max_freqs)) * time_step

    # ---- linear fit over the longest segment, first de-chirp                                         :300-355
    seg_time, seg_freq = t_axis[start_idx:end_idx], max_freqs[start_idx:end_idx]
    fit_middle_percent = params['fit_middle_percent']
    if fit_middle_percent < 100:
        trim = int(len(seg_time) * ((100 - fit_middle_percent) / 2 / 100))
        if trim > 0 and 2 * trim < len(seg_time):
            seg_time, seg_freq = seg_time[trim:len(seg_time) - trim], seg_freq[trim:len(seg_freq) - trim]
    _, coefs = _polyfit_coefs(seg_time, seg_freq, 1)
    f_shift_rate = coefs[1]                                            # Hz/s
    if trace is not None:
        trace.update(f_shift_rate_linear=float(f_shift_rate))
    lin_dev = _dechirp(wave_dev, f_shift_rate / (2.0 * fs * fs), 0.0)  # exp(-2j*pi*(rate*n^2/2/fs)/fs)
    if not params['precise_sync']:
        return lin_dev.cpu().numpy(), f_shift_rate / fs

    # ---- second ridge, Costas-template time sync on it                                               :364-463
    max_freq_indices_2nd, _ = _ridge(lin_dev, fs, bins_per_tone, steps_per_symbol)
    max_freqs_2nd = max_freq_indices_2nd * freq_step
    if trace is not None:
        trace.update(ridge2=max_freq_indices_2nd)
    sync_seq = np.array([3, 1, 4, 0, 6, 5, 2]) + 1
    sync_seq = sync_seq - np.mean(sync_seq)
    samples_per_sym = time_osr * 2
    gfsk_shape = gfsk_pulse(bt=2.0, t=np.linspace(-1, 1, samples_per_sym + 1))
    sync_correlation_seq = np.zeros((nsync_sym - 1) * time_osr + samples_per_sym + 1)
    for sym_idx in range(nsync_sym):
        sync_correlation_seq[sym_idx * time_osr:sym_idx * time_osr + samples_per_sym + 1] += gfsk_shape * sync_seq[sym_idx]
    three_sync = np.zeros((3 * nsync_sym + ndata_sym - 1) * time_osr + 1 + samples_per_sym)
    for i in range(3):
        s0 = i * (nsync_sym + ndata_sym // 2) * time_osr
        three_sync[s0:s0 + len(sync_correlation_seq)] = sync_correlation_seq
    start_idx, end_idx = max(signal_segments, key=lambda s: s[1] - s[0])
    end_idx = end_idx + window_size - 2                               # :421
    masked = np.zeros_like(max_freqs_2nd)
    masked[start_idx:end_idx] = max_freqs_2nd[start_idx:end_idx]
    masked[start_idx:end_idx] = masked[start_idx:end_idx] - np.mean(masked[start_idx:end_idx])
    sync_correlation = np.correlate(masked, three_sync, mode='full')
    peak = int(np.argmax(sync_correlation))
    peak_block = peak - (len(three_sync) - 1) + samples_per_sym // 2
    if trace is not None:
        trace.update(sync_block=int(peak_block))

    # ---- regression over the three Costas blocks, final de-chirp                                     :505-659
    regression_x, regression_y = np.array([]), np.array([])
    for i in range(3):
        s0 = i * (nsync_sym + ndata_sym // 2) * time_osr + peak_block
        e0 = s0 + (nsync_sym - 1) * time_osr
        if s0 < len(masked):
            hi = min(e0, len(masked))
            regression_x = np.append(regression_x, np.arange(s0, hi) * (sym_t / time_osr))
            regression_y = np.append(regression_y, masked[s0:hi])   # negative s0 slices like the reference's
    if len(regression_x) < 10:
        return lin_dev.cpu().numpy(), f_shift_rate / fs
    degree = params['poly_degree']
    if len(regression_x) > degree + 1:
        intercept, c = _polyfit_coefs(regression_x, regression_y, degree)
        rate_final = c[1] if len(c) > 1 else 0.0                        # Hz/s
        acc_final = c[2] if len(c) > 2 else 0.0                         # Hz/s^2
        if degree == 1:
            final_dev = _dechirp(lin_dev, rate_final / (2.0 * fs * fs), 0.0)
        elif degree == 2:
            final_dev = _dechirp(lin_dev, rate_final / (2.0 * fs * fs), acc_final / (3.0 * fs ** 3))
        else:
            return lin_dev.cpu().numpy(), f_shift_rate / fs           # :636-638
        poly = lambda x: intercept + sum(c[k] * x ** k for k in range(1, degree + 1))  # noqa: E731
        x0, x1 = regression_x[0], regression_x[-1]
        f_shift_rate_real = (poly(x0) - poly(x1)) / (x0 - x1) + f_shift_rate
        return final_dev.cpu().numpy(), f_shift_rate_real / fs
    return lin_dev.cpu().numpy(), f_shift_rate / fs
