"""Synthetic FT8 transmit chain (test-vector / benchmark-input source).

The reference builds its test signals with `ft8_tools.ft8_generator`
(encoder.py:15-72, crc.py:25-47, ldpc.py:104-131, modulator.py:27-90), whose per-sample
Python loops cost ~3 s per signal.  This module is a vectorised restatement that produces the
same waveform (bit-identical frequency track and phase recursion; checked against the
reference in tests/test_oracle_vs_reference.py), plus a batched on-device generator used by
bench.py to synthesise thousands of slots in HBM.  It is SURVEY.md section 8(f) row 2 - an
input source, not part of the accelerated decode path.
"""
import math

import numpy as np

from .ft8_demodulator import constants as _C

FT8_SYMBOL_TIME_S = 0.16
FT8_SYMBOL_FREQ_INTERVAL_HZ = 6.25
FT8_SYMBOL_NUM = 79
FT8_GRAY_MAP = np.array([0, 1, 3, 2, 5, 6, 4, 7], dtype=np.uint8)
FT8_COSTAS_PATTERN = np.array([3, 1, 4, 0, 6, 5, 2], dtype=np.uint8)

_G = None


def _generator_rows():
    global _G
    if _G is None:
        _G = _C.generator_matrix()
    return _G


def crc14(message: bytes, num_bits: int) -> int:
    """CRC-14, polynomial 0x2757, MSB first (reference ft8_generator/crc.py:10-23)."""
    rem = 0
    idx = 0
    for bit in range(num_bits):
        if bit % 8 == 0:
            rem ^= message[idx] << 6
            idx += 1
        rem = ((rem << 1) ^ 0x2757) if (rem & 0x2000) else (rem << 1)
        rem &= 0x3FFF
    return rem


def crc14_bits(payload_bits: np.ndarray) -> np.ndarray:
    """Vectorised CRC-14 of [n,77] payload bit arrays (zero-extended to 82 bits) -> [n,14] CRC bits, MSB first."""
    n = payload_bits.shape[0]
    bits = np.zeros((n, 88), dtype=np.int64)
    bits[:, :77] = payload_bits
    by = np.packbits(bits.astype(np.uint8), axis=1).astype(np.int64)      # [n,11] message bytes
    rem = np.zeros(n, dtype=np.int64)
    for bit in range(82):
        if bit % 8 == 0:
            rem ^= by[:, bit // 8] << 6
        top = (rem & 0x2000) != 0
        rem = np.where(top, (rem << 1) ^ 0x2757, rem << 1) & 0x3FFF
    return ((rem[:, None] >> np.arange(13, -1, -1)[None, :]) & 1).astype(np.uint8)


def add_crc(payload10) -> np.ndarray:
    """77-bit payload (10 bytes, low 3 bits of byte 9 ignored) -> a91 (12 bytes). crc.py:25-47."""
    a91 = np.zeros(12, dtype=np.uint8)
    if isinstance(payload10, (bytes, bytearray)):
        payload10 = np.frombuffer(bytes(payload10), dtype=np.uint8)
    a91[:10] = np.asarray(payload10, dtype=np.uint8)[:10]
    a91[9] &= 0xF8
    c = crc14(bytes(a91), 96 - 14)
    a91[9] |= (c >> 11) & 0xFF
    a91[10] = (c >> 3) & 0xFF
    a91[11] = (c << 5) & 0xFF
    return a91


def ldpc_encode(a91) -> np.ndarray:
    """a91 (12 bytes) -> 174-bit codeword as 22 bytes (ldpc.py:104-131); parity = G.msg over GF(2)."""
    msg = np.unpackbits(np.asarray(a91, dtype=np.uint8))[:_C.FTX_LDPC_K]
    par = (_generator_rows().astype(np.int64) @ msg.astype(np.int64)) & 1
    bits = np.concatenate([msg, par.astype(np.uint8), np.zeros(2, dtype=np.uint8)])
    return np.packbits(bits)


def codeword_bits(payload10) -> np.ndarray:
    """174 codeword bits (uint8 0/1) for a payload."""
    return np.unpackbits(ldpc_encode(add_crc(payload10)))[:_C.FTX_LDPC_N]


def ft8_encode(payload10) -> np.ndarray:
    """Payload -> 79 tone indices: S7 D29 S7 D29 S7 with Gray-mapped data (encoder.py:15-72)."""
    bits = codeword_bits(payload10)
    sym = bits.reshape(58, 3)
    data = FT8_GRAY_MAP[(sym[:, 0] << 2) | (sym[:, 1] << 1) | sym[:, 2]]
    return np.concatenate([FT8_COSTAS_PATTERN, data[:29], FT8_COSTAS_PATTERN, data[29:], FT8_COSTAS_PATTERN]
                          ).astype(np.uint8)


def _gauss_pulse(sps: int) -> np.ndarray:
    """BT=2.0 Gaussian-filtered unit pulse over 3 symbols (modulator.py:20-35)."""
    from scipy.special import erf
    t = (np.arange(3 * sps) - 1.5 * sps) / sps
    k = np.pi * np.sqrt(2 / np.log(2))
    return 0.5 * (erf(k * 2.0 * (t + 0.5)) - erf(k * 2.0 * (t - 0.5)))


def gfsk_frequency_track(itones, fs: float) -> np.ndarray:
    """Instantaneous frequency offset (Hz), (79+2)*sps samples (modulator.py:27-48).

    The per-symbol accumulation order is kept so the result equals the reference bit for bit."""
    sps = int(FT8_SYMBOL_TIME_S * fs)
    pulse = _gauss_pulse(sps)
    freq = np.zeros((FT8_SYMBOL_NUM + 2) * sps, dtype=np.float64)
    for i in range(FT8_SYMBOL_NUM):
        freq[i * sps:i * sps + 3 * sps] += (FT8_SYMBOL_FREQ_INTERVAL_HZ * float(itones[i])) * pulse
    freq[:2 * sps] += (FT8_SYMBOL_FREQ_INTERVAL_HZ * float(itones[0])) * pulse[sps:]
    freq[FT8_SYMBOL_NUM * sps:] += (FT8_SYMBOL_FREQ_INTERVAL_HZ * float(itones[-1])) * pulse[:2 * sps]
    return freq


def baseband_from_track(freq_track, fs: float, f0: float) -> np.ndarray:
    """Complex burst from an instantaneous-frequency track (modulator.py:56-75): sin(phi) - j cos(phi) with the
    sequential fmod phase recursion of :66-68 and the 1/8-symbol raised-cosine ramps of :70-73."""
    sps = int(FT8_SYMBOL_TIME_S * fs)
    n = FT8_SYMBOL_NUM * sps
    dphi = 2 * np.pi * np.asarray(freq_track, dtype=np.float64) / fs + 2 * np.pi * f0 / fs
    phi = np.empty(n, dtype=np.float64)
    acc = 0.0
    two_pi = 2 * np.pi
    fmod = math.fmod
    dl = dphi.tolist()
    for i in range(n):
        phi[i] = acc
        acc = fmod(acc + dl[i], two_pi)
    wave = np.sin(phi) - 1j * np.cos(phi)
    nramp = sps // 8
    i = np.arange(nramp)
    wave[:nramp] *= 0.5 * (1 - np.cos(8 * np.pi * i / sps))
    wave[n - 1 - i] *= 0.5 * (1 + np.cos(8 * np.pi * i / sps))
    return wave


def ft8_baseband(payload10, fs: float, f0: float) -> np.ndarray:
    """Complex baseband burst, 79*sps samples (modulator.py:77-83)."""
    return baseband_from_track(gfsk_frequency_track(ft8_encode(payload10), fs), fs, f0)


def ft8_generator(payload10, fs: float, f0: float, fc: float = 0.0) -> np.ndarray:
    """Real passband burst (modulator.py:85-90)."""
    bb = ft8_baseband(payload10, fs, f0)
    return np.real(bb * np.exp(1j * 2 * np.pi * fc * np.arange(len(bb)) / fs))


# ----------------------------------------------------------------------------------------
# Seeded slot recipes of SURVEY.md section 8(d) (shared by tests and bench.py)
# ----------------------------------------------------------------------------------------
C1_PAYLOAD = bytes.fromhex("1c3f8a6ae207a1e39450")


def embed(burst, n_total: int, start: int) -> np.ndarray:
    out = np.zeros(n_total, dtype=np.float64)
    out[start:start + len(burst)] = burst
    return out


def make_slot(payload10, seed: int, fs: int = 12000, f0: float = 500.0, start_s: float = 0.5, snr_db: float = -10.0,
              duration_s: float = 15.0) -> np.ndarray:
    """One slot: a burst at `start_s`, white noise over the full Nyquist band with
    noise_power = mean(burst^2)/10^(snr/10) (the reference tests' convention,
    test_ft8_standard.py:51-54).  Returned as float32-representable float64 so the float32 copy
    handed to the GPU and the float64 copy handed to the oracle hold identical values."""
    w = ft8_generator(np.frombuffer(bytes(payload10), dtype=np.uint8), fs, f0, 0.0)
    n = int(round(duration_s * fs))
    slot = embed(w, n, int(round(start_s * fs)))
    sigma = math.sqrt(float(np.mean(w ** 2)) / 10 ** (snr_db / 10))
    slot += sigma * np.random.default_rng(seed).standard_normal(n)
    return slot.astype(np.float32).astype(np.float64)


def c1_slot() -> np.ndarray:
    """Config C1: payload 1c3f..9450, f0=500 Hz, start 0.5 s, -10 dB, default_rng(0)."""
    return make_slot(C1_PAYLOAD, seed=0)


def c3_slot(index: int):
    """Config C3 slot `index`: random payload, f0~U[200,2900], start~U[0,2] s, SNR~U[-16,-4] dB; seed=index.
    Returns (slot, payload bytes with byte 9 masked to 77 bits)."""
    rng = np.random.default_rng(1_000_003 + index)
    payload = rng.integers(0, 256, 10, dtype=np.uint8)
    payload[9] &= 0xF8
    f0 = float(rng.uniform(200, 2900))
    start = float(rng.uniform(0, 2.0))
    snr = float(rng.uniform(-16, -4))
    return make_slot(bytes(payload), seed=index, f0=f0, start_s=start, snr_db=snr), bytes(payload)


def c2_slot(n_signals: int = 50, seed: int = 2, fs: int = 12000):
    """Config C2: `n_signals` overlapping bursts over 200-2950 Hz, SNR -24..0 dB against a common
    noise floor sigma^2 = mean(unit burst^2).  Returns (slot, [payload bytes], f0[], start_s[], snr_db[])."""
    rng = np.random.default_rng(seed)
    n = 15 * fs
    slot = np.zeros(n, dtype=np.float64)
    payloads, f0s, starts, snrs = [], [], [], []
    p0 = None
    for _ in range(n_signals):
        payload = rng.integers(0, 256, 10, dtype=np.uint8)
        payload[9] &= 0xF8
        f0 = float(rng.uniform(200, 2950))
        start = float(rng.uniform(0, 2.0))
        snr = float(rng.uniform(-24, 0))
        w = ft8_generator(payload, fs, f0, 0.0)
        if p0 is None:
            p0 = float(np.mean(w ** 2))
        s0 = int(round(start * fs))
        slot[s0:s0 + len(w)] += 10 ** (snr / 20) * w
        payloads.append(bytes(payload)); f0s.append(f0); starts.append(start); snrs.append(snr)
    slot += math.sqrt(p0) * rng.standard_normal(n)
    return slot.astype(np.float32).astype(np.float64), payloads, f0s, starts, snrs


def c5_llr(n_words: int, seed: int = 0, snr2500_range=(-26.0, -20.0)):
    """Config C5: standalone LLR vectors.  Random payload -> codeword -> 58 data tones; per symbol 8
    complex bins z~CN(0,1) with the tone bin raised by 10^(s/20), s = SNR_2500 + 26 dB (per-bin
    Es/N0); dB magnitudes -> max-difference LLRs -> sqrt(24/var) normalisation (the arithmetic of
    ft8_decode.py:151-198, vectorised).  Returns (llr[n,174] float64, codeword bits[n,174] uint8)."""
    rng = np.random.default_rng(seed)
    G = _generator_rows().astype(np.uint8)
    msg = np.zeros((n_words, _C.FTX_LDPC_K), dtype=np.uint8)
    msg[:, :77] = rng.integers(0, 2, (n_words, 77), dtype=np.uint8)
    msg[:, 77:] = crc14_bits(msg[:, :77])
    par = (msg.astype(np.int32) @ G.T.astype(np.int32)) & 1
    bits = np.concatenate([msg, par.astype(np.uint8)], axis=1)
    sym = bits.reshape(n_words, 58, 3)
    tone = FT8_GRAY_MAP[(sym[..., 0] << 2) | (sym[..., 1] << 1) | sym[..., 2]].astype(np.int64)
    snr = rng.uniform(snr2500_range[0], snr2500_range[1], n_words)
    amp = 10 ** ((snr + 26.0) / 20)
    z = (rng.standard_normal((n_words, 58, 8)) + 1j * rng.standard_normal((n_words, 58, 8))) / math.sqrt(2)
    z[np.arange(n_words)[:, None], np.arange(58)[None, :], tone] += amp[:, None]
    db = 10 * np.log10(1e-12 + np.abs(z) ** 2)
    s2 = db[..., FT8_GRAY_MAP]          # s2[j] = wf[gray[j]]
    llr = np.empty((n_words, 58, 3))
    llr[..., 0] = s2[..., 4:8].max(-1) - s2[..., 0:4].max(-1)
    llr[..., 1] = s2[..., [2, 3, 6, 7]].max(-1) - s2[..., [0, 1, 4, 5]].max(-1)
    llr[..., 2] = s2[..., [1, 3, 5, 7]].max(-1) - s2[..., [0, 2, 4, 6]].max(-1)
    llr = llr.reshape(n_words, 174)
    var = llr.var(axis=1)
    llr *= np.sqrt(24.0 / var)[:, None]
    return llr, bits


def c4_capture(seed: int = 4, fs: int = 12000, f0: float = 300.0, fc: float = 500.0, drift_hz_s: float = 100.0,
               esn0_db: float = 35.0, payload10: bytes = C1_PAYLOAD) -> np.ndarray:
    """Config C4 (satellite pass): the recipe of the reference's test_correction.py:185-257 with a seeded RNG.
    Complex baseband burst at f0, up-converted by fc, zero-padded by its own length on both sides, multiplied by a
    linear chirp exp(2j*pi*k*n^2/(2 fs^2)) (k in Hz/s), plus complex AWGN with N0 = mean|x|^2 / 10^(EsN0/10),
    noise power N0*fs split evenly over the real and imaginary parts.  Returns complex128."""
    bb = ft8_baseband(np.frombuffer(bytes(payload10), dtype=np.uint8), fs, f0)
    x = bb * np.exp(1j * 2 * np.pi * fc * np.arange(len(bb)) / fs)
    pad = np.zeros(len(x), dtype=complex)
    x = np.concatenate((pad, x, pad))
    n = np.arange(len(x))
    x = x * np.exp(2j * np.pi * (drift_hz_s / fs) * n ** 2 / (2 * fs))
    energy = np.sum(np.abs(x) ** 2) / len(x)
    noise_std = math.sqrt(energy / (10 ** (esn0_db / 10)) * fs / 2)
    rng = np.random.default_rng(seed)
    return x + (rng.normal(0, noise_std, len(x)) + 1j * rng.normal(0, noise_std, len(x)))


def drifting_slot(payload10, seed: int, drift_hz_s: float, fs: int = 12000, f0: float = 1000.0, start_s: float = 0.5,
                  snr_db: float = -8.0, duration_s: float = 15.0) -> np.ndarray:
    """A 15-s real slot whose burst drifts linearly in frequency (satellite pass, BASELINE config 4: up to
    +-50 Hz per slot = +-3.33 Hz/s).  Same noise convention as make_slot."""
    bb = ft8_baseband(np.frombuffer(bytes(payload10), dtype=np.uint8), fs, f0)
    t = np.arange(len(bb)) / fs
    w = np.real(bb * np.exp(2j * np.pi * 0.5 * drift_hz_s * t * t))
    n = int(round(duration_s * fs))
    slot = embed(w, n, int(round(start_s * fs)))
    sigma = math.sqrt(float(np.mean(w ** 2)) / 10 ** (snr_db / 10))
    slot += sigma * np.random.default_rng(seed).standard_normal(n)
    return slot.astype(np.float32).astype(np.float64)


def synth_batch_torch(n_slots: int, device, seed: int = 0, fs: int = 12000, duration_s: float = 15.0,
                      snr_range=(-16.0, -4.0), f0_range=(200.0, 2900.0), start_range=(0.0, 2.0), dtype=None):
    """C3-style batch generated on `device` with torch (plumbing; bench input only).

    Each slot holds one burst with random payload/f0/start/SNR.  Phase is a float64 cumulative sum
    of the same GFSK frequency track the numpy generator uses.  Returns (waves[n_slots, N] float32,
    payloads uint8[n_slots,10] on CPU)."""
    import torch
    dtype = dtype or torch.float32
    rng = np.random.default_rng(seed)
    sps = int(FT8_SYMBOL_TIME_S * fs)
    n = int(round(duration_s * fs))
    nb = FT8_SYMBOL_NUM * sps
    payloads = rng.integers(0, 256, (n_slots, 10), dtype=np.uint8)
    payloads[:, 9] &= 0xF8
    tones = np.stack([ft8_encode(p) for p in payloads]).astype(np.float64)          # [B,79]
    f0 = rng.uniform(f0_range[0], f0_range[1], n_slots)
    start = np.round(rng.uniform(start_range[0], start_range[1], n_slots) * fs).astype(np.int64)
    snr = rng.uniform(snr_range[0], snr_range[1], n_slots)
    pulse = torch.as_tensor(_gauss_pulse(sps), device=device)                        # [3*sps]
    out = torch.empty((n_slots, n), dtype=dtype, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    ramp_i = torch.arange(sps // 8, device=device, dtype=torch.float64)
    ramp_up = 0.5 * (1 - torch.cos(8 * math.pi * ramp_i / sps))
    ramp_down = torch.flip(0.5 * (1 + torch.cos(8 * math.pi * ramp_i / sps)), dims=[0])
    chunk = 64
    for c0 in range(0, n_slots, chunk):
        c1 = min(c0 + chunk, n_slots)
        b = c1 - c0
        tn = torch.as_tensor(tones[c0:c1], device=device)                            # [b,79]
        ext = torch.cat([tn[:, :1], tn, tn[:, -1:]], dim=1)                          # [b,81] edge-extended
        freq = torch.zeros((b, (FT8_SYMBOL_NUM + 2) * sps), dtype=torch.float64, device=device)
        # overlap-add of 3-symbol pulses; symbol i of `ext` is centred on samples [(i-1)*sps, (i+2)*sps)
        fr = freq.view(b, FT8_SYMBOL_NUM + 2, sps)
        p3 = pulse.view(3, sps)
        for k in range(3):
            lo = max(0, 1 - k)
            hi = min(FT8_SYMBOL_NUM + 2, FT8_SYMBOL_NUM + 2 + 1 - k)
            # contribution of ext[:, i] part k lands in output symbol slot i-1+k
            src = ext[:, lo:hi]
            dst_lo = lo - 1 + k
            fr[:, dst_lo:dst_lo + src.shape[1], :] += FT8_SYMBOL_FREQ_INTERVAL_HZ * src[:, :, None] * p3[k][None, None, :]
        dphi = 2 * math.pi * (freq[:, :nb] + torch.as_tensor(f0[c0:c1], device=device)[:, None]) / fs
        phi = torch.cumsum(dphi, dim=1) - dphi
        burst = torch.sin(phi)
        burst[:, :sps // 8] *= ramp_up
        burst[:, nb - sps // 8:] *= ramp_down
        power = (burst ** 2).mean(dim=1)
        sigma = torch.sqrt(power / torch.as_tensor(10 ** (snr[c0:c1] / 10), device=device))
        slot = sigma[:, None] * torch.randn((b, n), dtype=torch.float64, device=device, generator=gen)
        idx = torch.as_tensor(start[c0:c1], device=device)[:, None] + torch.arange(nb, device=device)[None, :]
        slot.scatter_add_(1, idx, burst)
        out[c0:c1] = slot.to(dtype)
    return out, payloads
