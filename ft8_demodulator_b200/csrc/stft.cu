// stft.cu - S1: sliding-window STFT -> dB spectrogram (float64), one frame per CTA pass.
//
// Replaces calculate_spectrogram (reference spectrogram_analyse.py:19-66), whose arithmetic is
// scipy.signal.spectrogram(window='hann', nperseg, noverlap, nfft, detrend=False,
// return_onesided=False, scaling='spectrum') followed by 10*log10(1e-12 + |S|), and the row/column
// selection decode_ft8_message applies to it (ft8_decode.py:323-341).
//
// Design (B200): the whole frame lives in shared memory as M complex float64 values and is
// transformed in place by a mixed-radix Stockham FFT (radix 4/2/3/5 butterflies in registers, any
// other prime through an O(p) per-output pass).  Real input and even nfft use the packed
// real-to-complex trick (M = nfft/2); odd nfft falls back to a complex transform of length nfft.
// Hann window, power scaling, +1e-12, log10 and the band/time crop are fused into the load and the
// store, so HBM sees each audio sample ~once (the hop overlap is served by L2) and each dB value once.
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "stft1920.cuh"

namespace ft8b {

constexpr int kMaxRadices = 24;
constexpr int kMaxElemsPerThread = 16;   // M <= 16 * blockDim

struct StftPlan {
    int nfft = 0, nperseg = 0, M = 0, even = 0;
    int nrad = 0;
    int radix[kMaxRadices];
    double2 *tw = nullptr;      // exp(-2*pi*i*q/M), q in [0,M)
    double2 *twpost = nullptr;  // exp(-2*pi*i*k/nfft), k in [0,M]   (even mode only)
    double *win = nullptr;      // periodic Hann, nperseg
    double scale = 0.0;         // 1/(sum win)^2
    int nthreads = 0;
    size_t smem = 0;
    int big = 0;                // M does not fit shared memory: stft_db_big_kernel works in a global-memory scratch
    // canonical-geometry fast path (nperseg 1920, nfft 3840): extra twiddle tables for stft1920_kernel
    int fast = 0;
    double2 *tw2 = nullptr, *tw3 = nullptr, *logtab = nullptr;
};

struct StftArgs {
    const void *wave;
    int dtype;
    int64_t n_slots, slot_stride;
    int nperseg, hop, nfft, M, even;
    int nrad;
    int radix[kMaxRadices];
    const double2 *tw, *twpost;
    const double *win;
    double scale;
    int bin_lo, bin_hi;
    int64_t frame_lo;
    int64_t frames_out;   // frame_hi - frame_lo
    double *mag;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i : (x + iy)(-i) = y - ix
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }

template <int R>
__device__ __forceinline__ void butterfly(double2 (&u)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(double2 (&u)[2]) {
    double2 a = u[0], b = u[1];
    u[0] = cadd(a, b);
    u[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void butterfly<4>(double2 (&u)[4]) {
    double2 s02 = cadd(u[0], u[2]), d02 = csub(u[0], u[2]);
    double2 s13 = cadd(u[1], u[3]), d13 = csub(u[1], u[3]);
    u[0] = cadd(s02, s13);
    u[2] = csub(s02, s13);
    u[1] = cadd(d02, mul_mi(d13));
    u[3] = cadd(d02, mul_pi(d13));
}
template <>
__device__ __forceinline__ void butterfly<3>(double2 (&u)[3]) {
    const double kS = 0.86602540378443864676;  // sin(2*pi/3)
    double2 t1 = cadd(u[1], u[2]);
    double2 t2 = make_double2(u[0].x - 0.5 * t1.x, u[0].y - 0.5 * t1.y);
    double2 d = csub(u[1], u[2]);
    double2 t3 = make_double2(kS * d.x, kS * d.y);
    u[0] = cadd(u[0], t1);
    u[1] = cadd(t2, mul_mi(t3));
    u[2] = cadd(t2, mul_pi(t3));
}
template <>
__device__ __forceinline__ void butterfly<5>(double2 (&u)[5]) {
    const double c1 = 0.30901699437494742410;   // cos(2*pi/5)
    const double c2 = -0.80901699437494742410;  // cos(4*pi/5)
    const double s1 = 0.95105651629515357212;   // sin(2*pi/5)
    const double s2 = 0.58778525229247312917;   // sin(4*pi/5)
    double2 a = u[0];
    double2 t1 = cadd(u[1], u[4]), t2 = cadd(u[2], u[3]);
    double2 t3 = csub(u[1], u[4]), t4 = csub(u[2], u[3]);
    double2 m1 = make_double2(a.x + c1 * t1.x + c2 * t2.x, a.y + c1 * t1.y + c2 * t2.y);
    double2 m2 = make_double2(a.x + c2 * t1.x + c1 * t2.x, a.y + c2 * t1.y + c1 * t2.y);
    double2 n1 = make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    double2 n2 = make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    u[0] = make_double2(a.x + t1.x + t2.x, a.y + t1.y + t2.y);
    u[1] = cadd(m1, mul_mi(n1));
    u[4] = cadd(m1, mul_pi(n1));
    u[2] = cadd(m2, mul_mi(n2));
    u[3] = cadd(m2, mul_pi(n2));
}

// One in-place Stockham pass of radix R: every thread pulls its butterflies into registers, the
// block synchronises, then results go back to their autosorted positions.
//   in : s[j + t*T],            T = M/R, j in [0,T), k = j mod Ns
//   out: s[(j-k)*R + k + t*Ns], twiddle exp(-2*pi*i*t*k/(Ns*R)) = tw[t*k*M/(Ns*R)]
template <int R, int MAXB>
__device__ __forceinline__ void fft_pass(double2 *s, int M, int Ns, const double2 *__restrict__ tw, int tid,
                                         int nthr) {
    const int T = M / R;
    const int twstep = M / (Ns * R);
    double2 u[MAXB][R];
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = tid + b * nthr;
        if (j < T) {
            const int k = j % Ns;
#pragma unroll
            for (int t = 0; t < R; ++t) u[b][t] = s[j + t * T];
            if (Ns > 1) {
#pragma unroll
                for (int t = 1; t < R; ++t) u[b][t] = cmul(u[b][t], __ldg(&tw[t * k * twstep]));
            }
            butterfly<R>(u[b]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = tid + b * nthr;
        if (j < T) {
            const int k = j % Ns;
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; ++t) s[j0 + t * Ns] = u[b][t];
        }
    }
    __syncthreads();
}

// Any other prime radix p: each thread evaluates its outputs directly (p terms each).
__device__ __noinline__ void fft_pass_prime(double2 *s, int M, int Ns, int p, const double2 *__restrict__ tw,
                                            int tid, int nthr) {
    const int T = M / p;
    const int stepA = M / (Ns * p);
    const int stepB = M / p;
    double2 v[kMaxElemsPerThread];
#pragma unroll
    for (int b = 0; b < kMaxElemsPerThread; ++b) {
        const int o = tid + b * nthr;
        if (o < M) {
            const int k = o % Ns;
            const int t = (o / Ns) % p;
            const int j = (o / (Ns * p)) * Ns + k;
            int base = k * stepA + t * stepB;
            if (base >= M) base -= M;
            double2 acc = make_double2(0.0, 0.0);
            int idx = 0;
            for (int q = 0; q < p; ++q) {
                acc = cadd(acc, cmul(s[j + q * T], __ldg(&tw[idx])));
                idx += base;
                if (idx >= M) idx -= M;
            }
            v[b] = acc;
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < kMaxElemsPerThread; ++b) {
        const int o = tid + b * nthr;
        if (o < M) s[o] = v[b];
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ double load_sample(const T *x, int i) { return (double)x[i]; }
// int16 PCM: float32(x) / 32767.0f in float32, exactly what from_wave.py:62-66 computes on the host
template <>
__device__ __forceinline__ double load_sample<short>(const short *x, int i) { return (double)__fdiv_rn((float)x[i], 32767.0f); }

template <typename T> struct is_cplx { static constexpr bool value = false; };
template <> struct is_cplx<float2> { static constexpr bool value = true; };
template <> struct is_cplx<double2> { static constexpr bool value = true; };
__device__ __forceinline__ double2 load_cplx(const float2 *x, int i) { const float2 v = x[i]; return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ double2 load_cplx(const double2 *x, int i) { return x[i]; }
template <typename T> __device__ __forceinline__ double2 load_cplx(const T *, int) { return make_double2(0.0, 0.0); }
template <> __device__ __forceinline__ double load_sample<float2>(const float2 *, int) { return 0.0; }
template <> __device__ __forceinline__ double load_sample<double2>(const double2 *, int) { return 0.0; }

// BOUND = the largest block size the instantiation may be launched with (register budget).
template <typename WaveT, int BOUND>
__global__ void __launch_bounds__(BOUND, 1) stft_db_kernel(const StftArgs a) {
    extern __shared__ double2 s[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int M = a.M;
    const int64_t n_items = a.n_slots * a.frames_out;
    const int Fout = a.bin_hi - a.bin_lo;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t slot = item / a.frames_out;
        const int64_t fr = item - slot * a.frames_out;
        const WaveT *x = reinterpret_cast<const WaveT *>(a.wave) + slot * a.slot_stride + (fr + a.frame_lo) * a.hop;
        // ---- windowed load (zero padded to nfft) ----
        if (is_cplx<WaveT>::value) {          // complex audio: plain length-nfft transform, two-sided output
            for (int n = tid; n < M; n += nthr) {
                double2 v = make_double2(0.0, 0.0);
                if (n < a.nperseg) {
                    const double w = __ldg(&a.win[n]);
                    const double2 c = load_cplx(x, n);
                    v = make_double2(w * c.x, w * c.y);
                }
                s[n] = v;
            }
        } else if (a.even) {
            for (int n = tid; n < M; n += nthr) {
                const int i0 = 2 * n, i1 = 2 * n + 1;
                double re = 0.0, im = 0.0;
                if (i0 < a.nperseg) re = __ldg(&a.win[i0]) * load_sample(x, i0);
                if (i1 < a.nperseg) im = __ldg(&a.win[i1]) * load_sample(x, i1);
                s[n] = make_double2(re, im);
            }
        } else {
            for (int n = tid; n < M; n += nthr) {
                double re = 0.0;
                if (n < a.nperseg) re = __ldg(&a.win[n]) * load_sample(x, n);
                s[n] = make_double2(re, 0.0);
            }
        }
        __syncthreads();
        // ---- in-place mixed-radix FFT of length M ----
        int Ns = 1;
        for (int p = 0; p < a.nrad; ++p) {
            const int r = a.radix[p];
            switch (r) {
                case 4: fft_pass<4, 4>(s, M, Ns, a.tw, tid, nthr); break;
                case 2: fft_pass<2, 8>(s, M, Ns, a.tw, tid, nthr); break;
                case 3: fft_pass<3, 6>(s, M, Ns, a.tw, tid, nthr); break;
                case 5: fft_pass<5, 4>(s, M, Ns, a.tw, tid, nthr); break;
                default: fft_pass_prime(s, M, Ns, r, a.tw, tid, nthr); break;
            }
            Ns *= r;
        }
        // ---- untangle the packed real transform, |X|^2 * scale, dB, cropped store ----
        double *out = a.mag + (slot * a.frames_out + fr) * (int64_t)Fout;
        for (int k = a.bin_lo + tid; k < a.bin_hi; k += nthr) {
            double2 X;
            if (a.even) {
                const double2 z = s[k == M ? 0 : k];
                const double2 zr = s[k == 0 ? 0 : M - k];
                const double2 zc = make_double2(zr.x, -zr.y);
                const double2 e = make_double2(0.5 * (z.x + zc.x), 0.5 * (z.y + zc.y));
                const double2 o = csub(z, zc);
                const double2 pw = cmul(__ldg(&a.twpost[k]), o);
                X = make_double2(e.x + 0.5 * pw.y, e.y - 0.5 * pw.x);
            } else {
                X = s[k];
            }
            const double pwr = (X.x * X.x + X.y * X.y) * a.scale;
            out[k - a.bin_lo] = 10.0 * log10(1e-12 + pwr);
        }
        __syncthreads();
    }
}

// Frames too long for shared memory (M > 14 336 points: 96 kHz audio upwards, SDR baseband): the same Stockham
// factorisation, out of place between two global-memory buffers of the CTA (L2-resident for moderate M), every output
// of a pass summed directly from its r inputs (the index mapping of fft_pass_prime, for every radix: O(M r) per pass).
// Generality over speed - this path exists so that no sample rate is refused; the reference's own SDR script runs
// decode_ft8_message at 1 MHz (nfft = 320 000).
constexpr int kBigThreads = 512;
template <typename WaveT>
__global__ void __launch_bounds__(kBigThreads) stft_db_big_kernel(const StftArgs a, double2 *scratch) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int M = a.M;
    double2 *A = scratch + (size_t)blockIdx.x * 2 * (size_t)M, *B = A + M;
    const int64_t n_items = a.n_slots * a.frames_out;
    const int Fout = a.bin_hi - a.bin_lo;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t slot = item / a.frames_out;
        const int64_t fr = item - slot * a.frames_out;
        const WaveT *x = reinterpret_cast<const WaveT *>(a.wave) + slot * a.slot_stride + (fr + a.frame_lo) * a.hop;
        for (int n = tid; n < M; n += nthr) {
            double2 v = make_double2(0.0, 0.0);
            if (is_cplx<WaveT>::value) {
                if (n < a.nperseg) {
                    const double w = __ldg(&a.win[n]);
                    const double2 c = load_cplx(x, n);
                    v = make_double2(w * c.x, w * c.y);
                }
            } else if (a.even) {
                const int i0 = 2 * n, i1 = 2 * n + 1;
                if (i0 < a.nperseg) v.x = __ldg(&a.win[i0]) * load_sample(x, i0);
                if (i1 < a.nperseg) v.y = __ldg(&a.win[i1]) * load_sample(x, i1);
            } else if (n < a.nperseg) {
                v.x = __ldg(&a.win[n]) * load_sample(x, n);
            }
            A[n] = v;
        }
        __syncthreads();
        double2 *src = A, *dst = B;
        int Ns = 1;
        for (int p = 0; p < a.nrad; ++p) {
            const int r = a.radix[p];
            const int T = M / r, stepA = M / (Ns * r), stepB = M / r;
            for (int o = tid; o < M; o += nthr) {
                const int k = o % Ns;
                const int t = (o / Ns) % r;
                const int j = (o / (Ns * r)) * Ns + k;
                int base = (int)(((int64_t)k * stepA + (int64_t)t * stepB) % M);
                double2 acc = make_double2(0.0, 0.0);
                int idx = 0;
                for (int q = 0; q < r; ++q) {
                    acc = cadd(acc, cmul(src[j + (int64_t)q * T], __ldg(&a.tw[idx])));
                    idx += base;
                    if (idx >= M) idx -= M;
                }
                dst[o] = acc;
            }
            __syncthreads();
            double2 *tmp = src; src = dst; dst = tmp;
            Ns *= r;
        }
        double *out = a.mag + (slot * a.frames_out + fr) * (int64_t)Fout;
        for (int k = a.bin_lo + tid; k < a.bin_hi; k += nthr) {
            double2 X;
            if (a.even) {
                const double2 z = src[k == M ? 0 : k];
                const double2 zr = src[k == 0 ? 0 : M - k];
                const double2 zc = make_double2(zr.x, -zr.y);
                const double2 e = make_double2(0.5 * (z.x + zc.x), 0.5 * (z.y + zc.y));
                const double2 o = csub(z, zc);
                const double2 pw = cmul(__ldg(&a.twpost[k]), o);
                X = make_double2(e.x + 0.5 * pw.y, e.y - 0.5 * pw.x);
            } else {
                X = src[k];
            }
            const double pwr = (X.x * X.x + X.y * X.y) * a.scale;
            out[k - a.bin_lo] = 10.0 * log10(1e-12 + pwr);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Plan cache
// ------------------------------------------------------------------------------------------
static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int>, StftPlan> g_plans;

static void factorize(int M, int *radix, int *nrad) {
    int n = 0;
    while (M % 4 == 0) { radix[n++] = 4; M /= 4; }
    for (int p : {2, 3, 5})
        while (M % p == 0) { radix[n++] = p; M /= p; }
    for (int p = 7; M > 1; p += 2)
        while (M % p == 0) { radix[n++] = p; M /= p; }
    *nrad = n;
}

static int get_plan(int nperseg, int nfft, bool cplx, const StftPlan **out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return FT8B_ERR_NO_DEVICE;
    }
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto key = std::make_tuple(dev, nperseg, cplx ? -nfft : nfft);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = &it->second; return 0; }
    StftPlan p;
    p.nfft = nfft;
    p.nperseg = nperseg;
    p.even = (!cplx && nfft % 2 == 0) ? 1 : 0;     // packed real transform; complex input uses the full length
    p.M = p.even ? nfft / 2 : nfft;
    factorize(p.M, p.radix, &p.nrad);
    int nthr = (p.M + 7) / 8;
    nthr = ((nthr + 31) / 32) * 32;
    if (nthr < 64) nthr = 64;
    if (nthr > 1024) nthr = 1024;
    p.nthreads = nthr;
    p.smem = sizeof(double2) * (size_t)p.M;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if ((int64_t)p.M > (int64_t)kMaxElemsPerThread * nthr || p.smem > (size_t)max_smem) {
        // e.g. 96 kHz audio (M = 15 360) or SDR baseband: the frame does not fit a CTA's shared memory
        if ((int64_t)p.M > (int64_t)1 << 24) {
            set_error("nfft=%d needs a %d-point FFT per frame; the limit is %d points", nfft, p.M, 1 << 24);
            return FT8B_ERR_UNSUPPORTED;
        }
        p.big = 1;
        p.nthreads = 512;
        p.smem = 0;
    }
    const long double two_pi = 6.283185307179586476925286766559L;
    std::vector<double2> tw(p.M), twp(p.M + 1);
    for (int q = 0; q < p.M; ++q) {
        long double ang = -two_pi * (long double)q / (long double)p.M;
        tw[q] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    for (int k = 0; k <= p.M; ++k) {
        long double ang = -two_pi * (long double)k / (long double)nfft;
        twp[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    if (nperseg == 1920 && nfft == 3840 && !cplx) {
        // Quarter-turn symmetry made exact in the table: W^k = -i conj(W^(960-k)), so entries 480..960 are the mirrored
        // entries 480..0.  The specialised kernel may then keep only entries 0..480 (in shared memory) and form the rest
        // on the fly, and still agree bit for bit with every path that reads the full table.
        twp[480] = make_double2(0.70710678118654752440, -0.70710678118654752440);
        for (int k = 481; k <= 960; ++k) twp[k] = make_double2(-twp[960 - k].y, -twp[960 - k].x);
    }
    std::vector<double> win(nperseg);
    double wsum = 0.0;
    for (int i = 0; i < nperseg; ++i) {  // scipy get_window('hann', nperseg): periodic Hann
        win[i] = 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)i / (double)nperseg);
        wsum += win[i];
    }
    p.scale = 1.0 / (wsum * wsum);  // scaling='spectrum'
    FT8B_CUDA_TRY(cudaMalloc(&p.tw, sizeof(double2) * tw.size()));
    FT8B_CUDA_TRY(cudaMalloc(&p.twpost, sizeof(double2) * twp.size()));
    FT8B_CUDA_TRY(cudaMalloc(&p.win, sizeof(double) * win.size()));
    FT8B_CUDA_TRY(cudaMemcpy(p.tw, tw.data(), sizeof(double2) * tw.size(), cudaMemcpyHostToDevice));
    FT8B_CUDA_TRY(cudaMemcpy(p.twpost, twp.data(), sizeof(double2) * twp.size(), cudaMemcpyHostToDevice));
    FT8B_CUDA_TRY(cudaMemcpy(p.win, win.data(), sizeof(double) * win.size(), cudaMemcpyHostToDevice));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<float, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<double, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<float, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<double, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<short, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<short, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<float2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<float2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<double2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    FT8B_CUDA_TRY(cudaFuncSetAttribute(stft_db_kernel<double2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    if (nperseg == 1920 && nfft == 3840 && !cplx) {
        std::vector<double2> t2(16 * 15), t3(8 * kTw3Stride);
        for (int t = 0; t < 16; ++t)
            for (int k = 0; k < 15; ++k) {
                long double ang = -two_pi * (long double)(t * k) / 240.0L;
                t2[t * 15 + k] = make_double2((double)cosl(ang), (double)sinl(ang));
            }
        for (int t = 0; t < 8; ++t)
            for (int j = 0; j <= 240; ++j) {       // column 240 (= W8^t) serves the self-paired butterfly 0, see P3
                long double ang = -two_pi * (long double)(t * j) / 1920.0L;
                t3[t * kTw3Stride + j] = make_double2((double)cosl(ang), (double)sinl(ang));
            }
        std::vector<double2> lt(64);
        for (int i = 0; i < 64; ++i) {
            const double inv = 1.0 / (1.0 + ((double)i + 0.5) / 64.0);
            lt[i] = make_double2(inv, (double)(-10.0L * log10l((long double)inv)));
        }
        FT8B_CUDA_TRY(cudaMalloc(&p.logtab, sizeof(double2) * lt.size()));
        FT8B_CUDA_TRY(cudaMemcpy(p.logtab, lt.data(), sizeof(double2) * lt.size(), cudaMemcpyHostToDevice));
        FT8B_CUDA_TRY(cudaMalloc(&p.tw2, sizeof(double2) * t2.size()));
        FT8B_CUDA_TRY(cudaMalloc(&p.tw3, sizeof(double2) * t3.size()));
        FT8B_CUDA_TRY(cudaMemcpy(p.tw2, t2.data(), sizeof(double2) * t2.size(), cudaMemcpyHostToDevice));
        FT8B_CUDA_TRY(cudaMemcpy(p.tw3, t3.data(), sizeof(double2) * t3.size(), cudaMemcpyHostToDevice));
#define FT8B_STFT1920_ATTR(T, MB, V)                                                                                   \
    do {                                                                                                               \
        FT8B_CUDA_TRY(cudaFuncSetAttribute(stft1920_kernel<T, MB, V>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)stft1920_smem_bytes<T, V>()));                                         \
    } while (0)
        FT8B_STFT1920_ATTR(float, 3, true);  FT8B_STFT1920_ATTR(float, 3, false);
        FT8B_STFT1920_ATTR(double, 3, true); FT8B_STFT1920_ATTR(double, 3, false);
        FT8B_STFT1920_ATTR(short, 3, true);  FT8B_STFT1920_ATTR(short, 3, false);
#undef FT8B_STFT1920_ATTR
        p.fast = 1;
    }
    auto ins = g_plans.emplace(key, p);
    *out = &ins.first->second;
    return 0;
}

#ifdef FT8B_STFT_PROF
extern "C" int ft8b_debug_stft_prof(unsigned long long *out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_stft_prof, sizeof(g_stft_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_stft_prof, z, sizeof(z)); }
    return 0;
}
#endif

int release_plans() {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (auto &kv : g_plans) {
        cudaFree(kv.second.tw);
        cudaFree(kv.second.twpost);
        cudaFree(kv.second.win);
        cudaFree(kv.second.tw2);
        cudaFree(kv.second.tw3);
        cudaFree(kv.second.logtab);
    }
    g_plans.clear();
    return 0;
}

int stft_db(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride, int nperseg,
            int hop, int nfft, int bin_lo, int bin_hi, int64_t frame_lo, int64_t frame_hi, double *mag,
            cudaStream_t stream) {
    if (!wave || !mag) FT8B_ARG_FAIL("ft8b_stft_db: null pointer");
    if (wave_dtype < FT8B_DTYPE_F32 || wave_dtype > FT8B_DTYPE_C128) FT8B_ARG_FAIL("ft8b_stft_db: bad dtype %d", wave_dtype);
    const bool cplx = wave_dtype == FT8B_DTYPE_C64 || wave_dtype == FT8B_DTYPE_C128;
    if (nperseg < 1 || hop < 1 || nfft < nperseg) FT8B_ARG_FAIL("ft8b_stft_db: need 1 <= nperseg <= nfft, hop >= 1");
    if (n_slots < 0 || slot_stride < n_samples) FT8B_ARG_FAIL("ft8b_stft_db: bad batch geometry");
    const int64_t T = ft8b_num_frames(n_samples, nperseg, hop);
    // real audio: bins 0..nfft/2 (even) / 0..(nfft-1)/2 (odd); complex audio: all nfft bins in natural FFT order
    const int max_bin = cplx ? nfft : ((nfft % 2 == 0) ? nfft / 2 + 1 : (nfft + 1) / 2);
    if (bin_lo < 0 || bin_hi > max_bin || bin_lo > bin_hi) FT8B_ARG_FAIL("ft8b_stft_db: bin range [%d,%d) outside [0,%d]", bin_lo, bin_hi, max_bin);
    if (frame_lo < 0 || frame_hi > T || frame_lo > frame_hi) FT8B_ARG_FAIL("ft8b_stft_db: frame range outside [0,%lld]", (long long)T);
    if (n_slots == 0 || bin_lo == bin_hi || frame_lo == frame_hi) return 0;
    const StftPlan *plan = nullptr;
    int rc = get_plan(nperseg, nfft, cplx, &plan);
    if (rc) return rc;
    const char *force_generic = getenv("FT8B_STFT_GENERIC");
    if (plan->fast && !(force_generic && force_generic[0] == '1')) {
        Stft1920Args f;
        f.wave = wave; f.n_slots = n_slots; f.slot_stride = slot_stride; f.hop = hop;
        f.win = plan->win; f.tw2 = plan->tw2; f.tw3 = plan->tw3; f.twpost = plan->twpost; f.logtab = plan->logtab; f.scale = plan->scale;
        f.bin_lo = bin_lo; f.bin_hi = bin_hi; f.frame_lo = frame_lo; f.frames_out = frame_hi - frame_lo; f.mag = mag;
        const int64_t items = n_slots * f.frames_out;
        // grid-stride CTAs: 3 per SM x 148 SMs x 16 waves (measured per 4096 slots: 1 wave 8.47 ms, 4: 8.12, 8: 8.02,
        // 16: 7.98, 64: 8.05 - the block scheduler evens out SMs; beyond 16 the per-CTA table staging shows)
#ifndef FT8B_STFT_WAVES
#define FT8B_STFT_WAVES 16
#endif
        constexpr int64_t kGrid1920 = 148 * 3 * FT8B_STFT_WAVES;
        const int64_t g1920 = items < kGrid1920 ? items : kGrid1920;
        // vector loads of sample pairs need every frame start to be 2-element aligned
        // the staged path copies sample pairs with cp.async: every frame start must be pair aligned
        const size_t pair_bytes = wave_dtype == FT8B_DTYPE_F32 ? 8 : (wave_dtype == FT8B_DTYPE_I16 ? 4 : 16);
        const bool vec = (reinterpret_cast<uintptr_t>(wave) % pair_bytes == 0) && (slot_stride % 2 == 0) && (hop % 2 == 0);
        {
            const size_t eb = pair_bytes / 2;
            f.copy16 = vec && (reinterpret_cast<uintptr_t>(wave) % 16 == 0) && ((slot_stride * eb) % 16 == 0) &&
                       (((size_t)hop * eb) % 16 == 0);
        }
#define FT8B_STFT1920(T, MB)                                                                              \
    do {                                                                                                  \
        if (vec) stft1920_kernel<T, MB, true><<<(unsigned)g1920, kStft1920Threads, stft1920_smem_bytes<T, true>(), stream>>>(f);   \
        else stft1920_kernel<T, MB, false><<<(unsigned)g1920, kStft1920Threads, stft1920_smem_bytes<T, false>(), stream>>>(f);     \
    } while (0)
        if (wave_dtype == FT8B_DTYPE_F32) {
            FT8B_STFT1920(float, 3);
        } else if (wave_dtype == FT8B_DTYPE_I16) {
            FT8B_STFT1920(short, 3);
        } else {
            FT8B_STFT1920(double, 3);
        }
#undef FT8B_STFT1920
        FT8B_LAUNCH_CHECK("stft1920_kernel");
        return 0;
    }
    StftArgs a;
    a.wave = wave; a.dtype = wave_dtype; a.n_slots = n_slots; a.slot_stride = slot_stride;
    a.nperseg = nperseg; a.hop = hop; a.nfft = nfft; a.M = plan->M; a.even = plan->even;
    a.nrad = plan->nrad;
    for (int i = 0; i < kMaxRadices; ++i) a.radix[i] = i < plan->nrad ? plan->radix[i] : 1;
    a.tw = plan->tw; a.twpost = plan->twpost; a.win = plan->win; a.scale = plan->scale;
    a.bin_lo = bin_lo; a.bin_hi = bin_hi; a.frame_lo = frame_lo; a.frames_out = frame_hi - frame_lo;
    a.mag = mag;
    const int64_t n_items = n_slots * a.frames_out;
    if (plan->big) {
        // scratch: two M-point buffers per CTA, stream-ordered allocation (the one path of the library that allocates per
        // call: it is rare, and a cached buffer would have to be guarded against concurrent streams)
        int64_t gridb = n_items < 2 * 148 ? n_items : 2 * 148;
        while (gridb > 1 && (size_t)gridb * 2 * plan->M * sizeof(double2) > ((size_t)2 << 30)) gridb /= 2;   // <= 2 GiB
        double2 *scratch = nullptr;
        FT8B_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)gridb * 2 * plan->M * sizeof(double2), stream));
        switch (wave_dtype) {
            case FT8B_DTYPE_C64: stft_db_big_kernel<float2><<<(unsigned)gridb, kBigThreads, 0, stream>>>(a, scratch); break;
            case FT8B_DTYPE_C128: stft_db_big_kernel<double2><<<(unsigned)gridb, kBigThreads, 0, stream>>>(a, scratch); break;
            case FT8B_DTYPE_I16: stft_db_big_kernel<short><<<(unsigned)gridb, kBigThreads, 0, stream>>>(a, scratch); break;
            case FT8B_DTYPE_F32: stft_db_big_kernel<float><<<(unsigned)gridb, kBigThreads, 0, stream>>>(a, scratch); break;
            default: stft_db_big_kernel<double><<<(unsigned)gridb, kBigThreads, 0, stream>>>(a, scratch); break;
        }
        const cudaError_t le = cudaGetLastError();
        FT8B_CUDA_TRY(cudaFreeAsync(scratch, stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (le != cudaSuccess) return cuda_fail(le, "stft_db_big_kernel");
        return 0;
    }
    int64_t grid = n_items;
    const int64_t cap = 148 * 64;  // persistent-ish: many CTAs per SM, grid-stride over frames
    if (grid > cap) grid = cap;
    const bool small = plan->nthreads <= 256;
    if (wave_dtype == FT8B_DTYPE_C64) {
        if (small) stft_db_kernel<float2, 256><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
        else stft_db_kernel<float2, 1024><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
    } else if (wave_dtype == FT8B_DTYPE_C128) {
        if (small) stft_db_kernel<double2, 256><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
        else stft_db_kernel<double2, 1024><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
    } else if (wave_dtype == FT8B_DTYPE_I16) {
        if (small) stft_db_kernel<short, 256><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
        else stft_db_kernel<short, 1024><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
    } else if (wave_dtype == FT8B_DTYPE_F32) {
        if (small) stft_db_kernel<float, 256><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
        else stft_db_kernel<float, 1024><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
    } else {
        if (small) stft_db_kernel<double, 256><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
        else stft_db_kernel<double, 1024><<<(unsigned)grid, plan->nthreads, plan->smem, stream>>>(a);
    }
    FT8B_LAUNCH_CHECK("stft_db_kernel");
    return 0;
}

}  // namespace ft8b
