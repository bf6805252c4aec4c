"""NumPy model of csrc/stft1920 (3-pass Stockham 15x16x8 on M=1920 with paired real-FFT untangling).
Index arithmetic is written exactly as in the CUDA kernel so it can be checked without a GPU."""
import numpy as np

M = 1920
W = np.exp(-2j * np.pi * np.arange(M) / M)


def bf3(u):
    return [u[0] + u[1] + u[2]] + [u[0] + u[1] * np.exp(-2j * np.pi * r / 3) + u[2] * np.exp(-4j * np.pi * r / 3) for r in (1, 2)]


def bf4(u):
    s02, d02, s13, d13 = u[0] + u[2], u[0] - u[2], u[1] + u[3], u[1] - u[3]
    return [s02 + s13, d02 - 1j * d13, s02 - s13, d02 + 1j * d13]


def bf5(u):
    return [sum(u[a] * np.exp(-2j * np.pi * a * q / 5) for a in range(5)) for q in range(5)]


def bf15(x):
    Y = [bf5([x[3 * a + b] for a in range(5)]) for b in range(3)]
    out = [None] * 15
    for q in range(5):
        v = bf3([Y[b][q] * np.exp(-2j * np.pi * b * q / 15) for b in range(3)])
        for r in range(3):
            out[q + 5 * r] = v[r]
    return out


def bf16(x):
    Y = [bf4([x[4 * a + b] for a in range(4)]) for b in range(4)]
    out = [None] * 16
    for q in range(4):
        v = bf4([Y[b][q] * np.exp(-2j * np.pi * b * q / 16) for b in range(4)])
        for r in range(4):
            out[q + 4 * r] = v[r]
    return out


def bf8(a):
    b = [a[t] + a[t + 4] for t in range(4)]
    c = [a[t] - a[t + 4] for t in range(4)]
    c = [c[t] * np.exp(-2j * np.pi * t / 8) for t in range(4)]
    e, o = bf4(b), bf4(c)
    return [e[0], o[0], e[1], o[1], e[2], o[2], e[3], o[3]]


def frame_spectrum(xw):
    """xw: 1920 windowed real samples -> X[0..1920] of the 3840-point zero-padded DFT."""
    z = np.zeros(M, complex)
    z[:960] = xw[0::2] + 1j * xw[1::2]
    s = np.zeros(M, complex)
    # P1: R=15, Ns=1, 128 butterflies; in z[j + 128 t] (t >= 8 is zero padding), out s[j*15 + f]
    for j in range(128):
        u = [z[j + 128 * t] for t in range(15)]
        y = bf15(u)
        for f in range(15):
            s[j * 15 + f] = y[f]
    # P2: R=16, Ns=15, 120 butterflies; in s[j + 120 t], tw W240^(t k) = W[8 t k], out s[(j-k)*16 + k + 15 f]
    s2 = np.zeros(M, complex)
    for j in range(120):
        k = j % 15
        u = [s[j + 120 * t] * W[8 * t * k] for t in range(16)]
        y = bf16(u)
        for f in range(16):
            s2[(j - k) * 16 + k + 15 * f] = y[f]
    # P3: R=8, Ns=240; butterfly j in [0,240): in s2[j + 240 t], tw W[t j]; Z[j + 240 f] stays in registers
    def p3(j):
        return bf8([s2[j + 240 * t] * W[t * j] for t in range(8)])
    X = np.zeros(M + 1, complex)
    WP = np.exp(-2j * np.pi * np.arange(M + 1) / (2 * M))

    def pair(z1, zr, k):      # writes X[k] and X[M-k]
        zc = np.conj(zr)
        E, O = 0.5 * (z1 + zc), z1 - zc
        P = WP[k] * O
        X[k] = complex(E.real + 0.5 * P.imag, E.imag - 0.5 * P.real)
        X[M - k] = complex(E.real - 0.5 * P.imag, -E.imag - 0.5 * P.real)
    for i in range(1, 120):   # thread i: butterflies i and 240-i
        u1, u2 = p3(i), p3(240 - i)
        for t in range(8):
            pair(u1[t], u2[7 - t], i + 240 * t)
    u0, u120 = p3(0), p3(120)  # thread 0
    pair(u0[0], u0[0], 0)
    for t in range(1, 5):
        pair(u0[t], u0[8 - t], 240 * t)
    for t in range(4):
        pair(u120[t], u120[7 - t], 120 + 240 * t)
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1920)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1920) / 1920)
    ref = np.fft.fft(np.concatenate([x * w, np.zeros(1920)]))[:1921]
    got = frame_spectrum(x * w)
    print("max err", np.abs(got - ref).max())
