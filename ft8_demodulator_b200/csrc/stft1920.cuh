// stft1920.cuh - specialised S1 kernel for the canonical FT8 geometry (fs = 12 kHz, osr 2x2):
// nperseg = 1920, nfft = 3840, i.e. a packed real transform of M = 1920 complex float64 points.
//
// Plan (validated index-for-index by tools/model_stft1920.py): three in-place Stockham passes
// 15 x 16 x 8 with 128 threads per frame.
//   P1 radix 15, Ns = 1   : reads the windowed audio straight from global memory (the upper half of the
//                           zero-padded frame never exists in memory), writes s[j*15 + f]
//   P2 radix 16, Ns = 15  : s[j + 120 t] * W240^(t k)  ->  s[(j-k)*16 + k + 15 f]
//   P3 radix  8, Ns = 240 : thread i owns butterflies i+1 and 239-i, so Z[k] and Z[M-k] meet in registers;
//                           the real-FFT untangling, |X|^2 scaling, +1e-12, 10*log10 and the band crop
//                           are applied there and the dB values go straight to HBM.
// Shared-memory traffic per frame is 4 x 30 KB (one write + read per pass boundary), 4 block barriers.
#pragma once
#include <cuda_pipeline.h>

#include "common.cuh"

#ifdef FT8B_STFT_PROF
// experiment build only (tools/build_variant.sh ... -DFT8B_STFT_PROF): warp-cycles per phase of the frame loop
__device__ unsigned long long g_stft_prof[16];
#define FT8B_PROF_DECL long long _pt = clock64(); unsigned long long _pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define FT8B_PROF(i) do { const long long _n = clock64(); _pa[i] += (unsigned long long)(_n - _pt); _pt = _n; } while (0)
#define FT8B_PROF_END do { if ((threadIdx.x & 31) == 0) for (int _i = 0; _i < 8; ++_i) atomicAdd(&g_stft_prof[_i], _pa[_i]); } while (0)
#else
#define FT8B_PROF_DECL
#define FT8B_PROF(i)
#define FT8B_PROF_END
#endif

namespace ft8b {

constexpr int kM1920 = 1920;
constexpr int kStft1920Threads = 128;
constexpr int kTw3Stride = 248;     // row pitch of the P3 twiddle table: rows start on 128-byte lines, and not a power of two (L1 sets)

struct Stft1920Args {
    const void *wave;
    int64_t n_slots, slot_stride;
    int hop;
    const double *win;        // [1920]
    const double2 *tw2;       // [16][15]  exp(-2*pi*i*t*k/240)
    const double2 *tw3;       // [8][kTw3Stride]  exp(-2*pi*i*t*j/1920), j = 0..240
    const double2 *twpost;    // [1921]    exp(-2*pi*i*k/3840)
    const double2 *logtab;    // [64]      (1/c_i, 10*log10(c_i)), c_i = 1/fl(1/(1+(i+0.5)/64))
    double scale;
    int bin_lo, bin_hi;
    int64_t frame_lo, frames_out;
    double *mag;
    int copy16;               // every frame starts on a 16-byte boundary: stage it with 16-byte cp.async.cg
};

namespace f1920 {

// Butterfly / polynomial constants live in constant memory: DFMA/DMUL read them as c[bank][offset]
// operands, whereas 64-bit immediates would need two UMOV each, every time they are used.
__device__ __constant__ double kK[40] = {
    /* 0 */ 0.86602540378443864676,   // sin(2*pi/3)
    /* 1 */ 0.30901699437494742410,   // cos(2*pi/5)
    /* 2 */ -0.80901699437494742410,  // cos(4*pi/5)
    /* 3 */ 0.95105651629515357212,   // sin(2*pi/5)
    /* 4 */ 0.58778525229247312917,   // sin(4*pi/5)
    /* 5 */ 0.91354545764260089550, /* 6 */ 0.40673664307580020775,    // cos, sin(2*pi*1/15)
    /* 7 */ 0.66913060635885821383, /* 8 */ 0.74314482547739423501,    // 2/15
    /* 9 */ -0.10452846326765347140, /* 10 */ 0.99452189536827333692,  // 4/15
    /* 11 */ -0.97814760073380563793, /* 12 */ -0.20791169081775933710, // 8/15
    /* 13 */ 0.92387953251128675613, /* 14 */ 0.38268343236508977173,  // cos, sin(pi/8)
    /* 15 */ 0.70710678118654752440,  // sqrt(1/2)
    /* 16 */ 0.62042067417029334,     // (10/ln10)/7 ... alternating-sign Taylor coefficients of 10*log10(1+r)
    /* 17 */ -0.72382411986534223, /* 18 */ 0.86858894383841068, /* 19 */ -1.0857361797980134,
    /* 20 */ 1.4476482397306845, /* 21 */ -2.1714723595960267, /* 22 */ 4.3429448190325183,
    /* 23 */ 3.0102999566398120,      // 10*log10(2)
    /* 24 */ 1e-12,
    /* 25 */ 1e-300,                  // fast_db: carries inf / NaN into the result, vanishes for finite powers
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // * (-i)
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }   // * (+i)
// multiply by the constant (c - i*s)
__device__ __forceinline__ double2 cmulc(double2 a, double c, double s) {
    return make_double2(a.x * c + a.y * s, a.y * c - a.x * s);
}

__device__ __forceinline__ void bf4(double2 &a, double2 &b, double2 &c, double2 &d) {
    const double2 s02 = cadd(a, c), d02 = csub(a, c), s13 = cadd(b, d), d13 = csub(b, d);
    a = cadd(s02, s13);
    b = cadd(d02, mul_mi(d13));
    c = csub(s02, s13);
    d = cadd(d02, mul_pi(d13));
}
__device__ __forceinline__ void bf3(double2 &a, double2 &b, double2 &c) {
    const double kS = kK[0];
    const double2 t1 = cadd(b, c);
    const double2 t2 = make_double2(a.x - 0.5 * t1.x, a.y - 0.5 * t1.y);
    const double2 d = csub(b, c);
    const double2 t3 = make_double2(kS * d.x, kS * d.y);
    a = cadd(a, t1);
    b = cadd(t2, mul_mi(t3));
    c = cadd(t2, mul_pi(t3));
}
__device__ __forceinline__ void bf5(double2 &u0, double2 &u1, double2 &u2, double2 &u3, double2 &u4) {
    const double c1 = kK[1], c2 = kK[2], s1 = kK[3], s2 = kK[4];
    const double2 a = u0;
    const double2 t1 = cadd(u1, u4), t2 = cadd(u2, u3), t3 = csub(u1, u4), t4 = csub(u2, u3);
    const double2 m1 = make_double2(a.x + c1 * t1.x + c2 * t2.x, a.y + c1 * t1.y + c2 * t2.y);
    const double2 m2 = make_double2(a.x + c2 * t1.x + c1 * t2.x, a.y + c2 * t1.y + c1 * t2.y);
    const double2 n1 = make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    const double2 n2 = make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    u0 = make_double2(a.x + t1.x + t2.x, a.y + t1.y + t2.y);
    u1 = cadd(m1, mul_mi(n1));
    u4 = cadd(m1, mul_pi(n1));
    u2 = cadd(m2, mul_mi(n2));
    u3 = cadd(m2, mul_pi(n2));
}

// 5-point DFT whose inputs u3 = u4 = 0 / u2 = u3 = u4 = 0 (the zero-padded half of the frame): the sums and
// differences with a zero operand are dropped by hand - the compiler may not fold x + 0.0 (it would turn -0 into +0).
__device__ __forceinline__ void bf5_in3(double2 &u0, double2 &u1, double2 &u2, double2 &u3, double2 &u4) {
    const double c1 = kK[1], c2 = kK[2], s1 = kK[3], s2 = kK[4];
    const double2 a = u0, t1 = u1, t2 = u2;
    const double2 m1 = make_double2(a.x + c1 * t1.x + c2 * t2.x, a.y + c1 * t1.y + c2 * t2.y);
    const double2 m2 = make_double2(a.x + c2 * t1.x + c1 * t2.x, a.y + c2 * t1.y + c1 * t2.y);
    const double2 n1 = make_double2(s1 * t1.x + s2 * t2.x, s1 * t1.y + s2 * t2.y);
    const double2 n2 = make_double2(s2 * t1.x - s1 * t2.x, s2 * t1.y - s1 * t2.y);
    u0 = make_double2(a.x + t1.x + t2.x, a.y + t1.y + t2.y);
    u1 = cadd(m1, mul_mi(n1));
    u4 = cadd(m1, mul_pi(n1));
    u2 = cadd(m2, mul_mi(n2));
    u3 = cadd(m2, mul_pi(n2));
}
__device__ __forceinline__ void bf5_in2(double2 &u0, double2 &u1, double2 &u2, double2 &u3, double2 &u4) {
    const double c1 = kK[1], c2 = kK[2], s1 = kK[3], s2 = kK[4];
    const double2 a = u0, t1 = u1;
    const double2 m1 = make_double2(a.x + c1 * t1.x, a.y + c1 * t1.y);
    const double2 m2 = make_double2(a.x + c2 * t1.x, a.y + c2 * t1.y);
    const double2 n1 = make_double2(s1 * t1.x, s1 * t1.y);
    const double2 n2 = make_double2(s2 * t1.x, s2 * t1.y);
    u0 = cadd(a, t1);
    u1 = cadd(m1, mul_mi(n1));
    u4 = cadd(m1, mul_pi(n1));
    u2 = cadd(m2, mul_mi(n2));
    u3 = cadd(m2, mul_pi(n2));
}

// 15-point DFT, t = 3a + b in, f = q + 5r out (in place, natural order on both sides).
// HALF: inputs x[8..14] are zero (P1 of the zero-padded frame) and are never read.
template <bool HALF = false>
__device__ __forceinline__ void bf15(double2 (&x)[15]) {
    // cos/sin(2*pi*m/15)
    const double C1 = kK[5], S1 = kK[6], C2 = kK[7], S2 = kK[8], C3 = kK[1], S3 = kK[3];
    const double C4 = kK[9], S4 = kK[10], C6 = kK[2], S6 = kK[4], C8 = kK[11], S8 = kK[12];
    if (HALF) {
        bf5_in3(x[0], x[3], x[6], x[9], x[12]);
        bf5_in3(x[1], x[4], x[7], x[10], x[13]);
        bf5_in2(x[2], x[5], x[8], x[11], x[14]);
    } else {
        bf5(x[0], x[3], x[6], x[9], x[12]);     // b = 0 : Y0[q] in x[3q]
        bf5(x[1], x[4], x[7], x[10], x[13]);    // b = 1 : Y1[q] in x[3q+1]
        bf5(x[2], x[5], x[8], x[11], x[14]);    // b = 2 : Y2[q] in x[3q+2]
    }
    x[4] = cmulc(x[4], C1, S1);   x[5] = cmulc(x[5], C2, S2);     // q = 1 : W15^1, W15^2
    x[7] = cmulc(x[7], C2, S2);   x[8] = cmulc(x[8], C4, S4);     // q = 2 : W15^2, W15^4
    x[10] = cmulc(x[10], C3, S3); x[11] = cmulc(x[11], C6, S6);   // q = 3 : W15^3, W15^6
    x[13] = cmulc(x[13], C4, S4); x[14] = cmulc(x[14], C8, S8);   // q = 4 : W15^4, W15^8
    bf3(x[0], x[1], x[2]);        // q = 0 : f = 0, 5, 10
    bf3(x[3], x[4], x[5]);        // q = 1 : f = 1, 6, 11
    bf3(x[6], x[7], x[8]);
    bf3(x[9], x[10], x[11]);
    bf3(x[12], x[13], x[14]);
    // now x[3q + r] = X[q + 5r]; caller maps with out_index15()
}
__device__ __forceinline__ constexpr int out15(int pos) { return (pos / 3) + 5 * (pos % 3); }   // f held at x[pos]

// 16-point DFT, t = 4a + b in, f = q + 4r out; after the call x[4q + r] = X[q + 4r]
__device__ __forceinline__ void bf16(double2 (&x)[16]) {
    const double C = kK[13], S = kK[14], H = kK[15];
    bf4(x[0], x[4], x[8], x[12]);     // b = 0 : Y0[q] in x[4q]
    bf4(x[1], x[5], x[9], x[13]);     // b = 1 : Y1[q] in x[4q+1]
    bf4(x[2], x[6], x[10], x[14]);
    bf4(x[3], x[7], x[11], x[15]);
    // Yb[q] *= W16^(b q)
    x[5] = cmulc(x[5], C, S);    x[6] = cmulc(x[6], H, H);    x[7] = cmulc(x[7], S, C);        // q=1: W^1 W^2 W^3
    x[9] = cmulc(x[9], H, H);    x[10] = mul_mi(x[10]);       x[11] = cmulc(x[11], -H, H);     // q=2: W^2 W^4 W^6
    x[13] = cmulc(x[13], S, C);  x[14] = cmulc(x[14], -H, H); x[15] = cmulc(x[15], -C, -S);    // q=3: W^3 W^6 W^9
    bf4(x[0], x[1], x[2], x[3]);
    bf4(x[4], x[5], x[6], x[7]);
    bf4(x[8], x[9], x[10], x[11]);
    bf4(x[12], x[13], x[14], x[15]);
}
__device__ __forceinline__ constexpr int out16(int pos) { return (pos / 4) + 4 * (pos % 4); }

// 8-point DFT in place, natural order in and out
__device__ __forceinline__ void bf8(double2 (&a)[8]) {
    const double H = kK[15];
    double2 b0 = cadd(a[0], a[4]), b1 = cadd(a[1], a[5]), b2 = cadd(a[2], a[6]), b3 = cadd(a[3], a[7]);
    double2 c0 = csub(a[0], a[4]), c1 = csub(a[1], a[5]), c2 = csub(a[2], a[6]), c3 = csub(a[3], a[7]);
    c1 = cmulc(c1, H, H);
    c2 = mul_mi(c2);
    c3 = cmulc(c3, -H, H);
    bf4(b0, b1, b2, b3);
    bf4(c0, c1, c2, c3);
    a[0] = b0; a[1] = c0; a[2] = b1; a[3] = c1; a[4] = b2; a[5] = c2; a[6] = b3; a[7] = c3;
}

// 10*log10(v) for finite v >= 1e-12 without the libdevice log10 (a division, ~35 FP64 instructions and a
// slow-path branch per element).  v = 2^e * m, m in [1,2); i = top 6 mantissa bits; r = m * (1/c_i) - 1 with
// |r| <= 2^-7; 10*log10(v) = e*10*log10(2) + 10*log10(c_i) + (10/ln 10) * (r - r^2/2 + ... + r^5/5).
// The series is cut after r^5: the first dropped term is at most (10/ln 10) * 2^-42 / 6 = 1.6e-13 dB, an order of
// magnitude below what the FFT's own rounding leaves against pocketfft (~5e-12 dB) and far inside the 1e-9 dB bound.
// `tab` is this lane's view of the shared table: 8 interleaved copies, entry i of copy q at tab8[i*8 + q], so the
// 8 lanes of a quarter-warp (one 128-byte shared-memory wavefront for 16-byte loads) never collide.
__device__ __forceinline__ double fast_db(double v, const double2 *__restrict__ tab) {
    // Branch-free and select-free.  inf / NaN inputs (never produced by finite audio) reach the result through the
    // term v * 1e-300: for every finite power it is below half an ulp of the table value it is added to (v < 1e280 -
    // float32 audio cannot exceed 1e90) and changes nothing; +inf stays +inf and NaN stays NaN, as 10*log10 would give
    // (one DFMA instead of the compare, the v + v and the two selects that used to patch them in: -3 instructions per bin).
    const unsigned hi = (unsigned)__double2hiint(v);
    const int e = (int)(hi >> 20) - 1023;
    // entry i = top 6 mantissa bits of this lane's copy: byte offset i * 128 = (hi >> 7) & 0x1f80 (one shift, one mask, one add)
    const double2 t = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(tab) + ((hi >> 7) & 0x1f80u));
    unsigned mh;                                    // (hi & 0xfffff) | 0x3ff00000 as ONE three-input logic op
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(mh) : "r"(hi), "r"(0x000fffffu), "r"(0x3ff00000u));
    const double m = __hiloint2double((int)mh, __double2loint(v));
    const double r = fma(m, t.x, -1.0);
    double p = kK[18];                              //  (10/ln10)/5
    p = fma(p, r, kK[19]);                          // -(10/ln10)/4
    p = fma(p, r, kK[20]);                          //  (10/ln10)/3
    p = fma(p, r, kK[21]);                          // -(10/ln10)/2
    p = fma(p, r, kK[22]);                          //  (10/ln10)
    return fma(r, p, fma((double)e, kK[23], fma(v, kK[25], t.y)));
}

// Untangle the pair (Z[k], Z[M-k]) of the packed transform into dB for bins k and M-k
// (spectrogram_analyse.py:59-60: 10*log10(1e-12 + |S|), S = |X|^2 / (sum w)^2).
__device__ __forceinline__ void untangle_pair(double2 z, double2 zr, double2 w, double scale4, const double2 *__restrict__ tab,
                                              double &d1, double &d2) {
    // X[k] = E + (-i) w O with E = (z + conj zr)/2, O = (z - conj zr)/2.  The halves are left out here and folded into
    // scale4 = scale/4: scaling by a power of two commutes with rounding, so the dB values are the same bit for bit.
    // The contractions are spelled out so that the full-band and the cropped store paths (two inlined copies of this
    // code) round the same way: their outputs are compared bit for bit.
    const double2 e = make_double2(z.x + zr.x, z.y - zr.y);      // z + conj zr
    const double2 d = make_double2(z.x - zr.x, z.y + zr.y);      // z - conj zr
    const double2 p = make_double2(fma(w.x, d.x, -__dmul_rn(w.y, d.y)), fma(w.x, d.y, __dmul_rn(w.y, d.x)));
    const double re1 = e.x + p.y, im1 = e.y - p.x;
    const double re2 = e.x - p.y, im2 = e.y + p.x;
    d1 = fast_db(fma(fma(re1, re1, __dmul_rn(im1, im1)), scale4, kK[24]), tab);
    d2 = fast_db(fma(fma(re2, re2, __dmul_rn(im2, im2)), scale4, kK[24]), tab);
}

// general form with band crop: stores bins k and M-k when they fall inside [bin_lo, bin_hi)
__device__ __forceinline__ void emit_pair(double2 z, double2 zr, int k, bool mirror, const Stft1920Args &a, double *out,
                                          const double2 *__restrict__ tab) {
    double d1, d2;
    // post-twiddle of k >= 960 as -i W3840^(k - 960), exactly as the full-band path forms it (same bits on both paths)
    const double2 w0 = __ldg(&a.twpost[k >= 960 ? k - 960 : k]);
    untangle_pair(z, zr, k >= 960 ? mul_mi(w0) : w0, 0.25 * a.scale, tab, d1, d2);
    const int k2 = kM1920 - k;
    if (k >= a.bin_lo && k < a.bin_hi) out[k - a.bin_lo] = d1;
    if (mirror && k2 >= a.bin_lo && k2 < a.bin_hi) out[k2 - a.bin_lo] = d2;
}

}  // namespace f1920

// samples 2n, 2n+1 as one vector load (the caller guarantees 8/16-byte alignment when VEC is set)
__device__ __forceinline__ double2 load_pair(const float *x, int n) {
    const float2 v = reinterpret_cast<const float2 *>(x)[n];
    return make_double2((double)v.x, (double)v.y);
}
__device__ __forceinline__ double2 load_pair(const double *x, int n) { return reinterpret_cast<const double2 *>(x)[n]; }
__device__ __forceinline__ double from_pcm16(short v) { return (double)__fdiv_rn((float)v, 32767.0f); }   // from_wave.py:62-66
__device__ __forceinline__ double2 load_pair(const short *x, int n) {
    const short2 v = reinterpret_cast<const short2 *>(x)[n];
    return make_double2(from_pcm16(v.x), from_pcm16(v.y));
}
__device__ __forceinline__ double to_sample(float v) { return (double)v; }
__device__ __forceinline__ double to_sample(double v) { return v; }
__device__ __forceinline__ double to_sample(short v) { return from_pcm16(v); }

#ifndef FT8B_STFT_SMEMTAB
#define FT8B_STFT_SMEMTAB 1      // P3 twiddles + post-twiddles in shared memory (7.93 -> 7.81 ms per 4096 slots; 0 = __ldg from L1)
#endif
constexpr int kTw3Cols = 121;       // columns 0..120 of tw3 are the only ones P3 touches
constexpr int kTwpLen = 481;        // post-twiddles 0..480; 481..960 follow from W^k = -i conj(W^(960-k))
constexpr int kSmemTabLen = FT8B_STFT_SMEMTAB ? 7 * kTw3Cols + kTwpLen : 0;

template <typename WaveT, bool VEC>
constexpr size_t stft1920_smem_bytes() {
    return sizeof(double2) * (kM1920 + 512 + 16 * 15 + kSmemTabLen) + (VEC ? sizeof(WaveT) * kM1920 : 16);
}

// MB = CTAs per SM the register allocation must allow.  3 (168 registers, no spills, 12 warps/SM, 92 KB of L1 left
// for the twiddle tables) measured 13 % faster than 4 (128 registers, 264 B of spills, 28 KB of L1) and 8 % faster than 2.
template <typename WaveT, int MB, bool VEC>
__global__ void __launch_bounds__(kStft1920Threads, MB) stft1920_kernel(const Stft1920Args a) {
    using namespace f1920;
    extern __shared__ __align__(16) unsigned char smem1920[];     // stft1920_smem_bytes<WaveT, VEC>()
    double2 *s = reinterpret_cast<double2 *>(smem1920);           // [1920] the frame being transformed
    double2 *logtab8 = s + kM1920;                                  // [64][8] interleaved copies of the log table
    double2 *tw2s = logtab8 + 512;                                // [16*15]
    // VEC path: the next frame's 1920 samples are staged here with cp.async while this frame is transformed
#if FT8B_STFT_SMEMTAB
    // P3 twiddles and post-twiddles live in shared memory (21 KB per CTA, 215 KB for the three CTAs of an SM): their
    // 11 loads per thread and frame no longer depend on what the audio and spectrogram streams leave of L1
    double2 *tw3s = tw2s + 16 * 15;                               // [7][121]  rows t = 1..7
    double2 *twps = tw3s + 7 * kTw3Cols;                          // [481]
    WaveT *xs = reinterpret_cast<WaveT *>(twps + kTwpLen);        // [1920]
#else
    WaveT *xs = reinterpret_cast<WaveT *>(tw2s + 16 * 15);        // [1920]
#endif
    const int tid = threadIdx.x;
    // The tables (33 KB per CTA) are staged with 16-byte cp.async copies: all 17 copies of a thread are in flight at
    // once and land under the first frame's audio staging (one wait + barrier below covers both).  A load-then-store
    // loop keeps ONE load in flight per thread and iteration - 17 L2 round trips, as long as a whole frame takes,
    // which is half of a single-slot call (one frame per CTA) and 3 % of the 4096-slot batch.
    auto copy16 = [](void *dst, const void *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
#if FT8B_STFT_SMEMTAB
#pragma unroll
    for (int r = 0; r < (7 * kTw3Cols + kStft1920Threads - 1) / kStft1920Threads; ++r) {
        const int i = tid + r * kStft1920Threads;
        if (i < 7 * kTw3Cols) {
            const int t = i / kTw3Cols + 1, j = i - (t - 1) * kTw3Cols;
            copy16(&tw3s[i], &a.tw3[t * kTw3Stride + j]);
        }
    }
#pragma unroll
    for (int r = 0; r < (kTwpLen + kStft1920Threads - 1) / kStft1920Threads; ++r) {
        const int i = tid + r * kStft1920Threads;
        if (i < kTwpLen) copy16(&twps[i], &a.twpost[i]);
    }
#endif
#pragma unroll
    for (int r = 0; r < 512 / kStft1920Threads; ++r) {
        const int i = tid + r * kStft1920Threads;
        copy16(&logtab8[i], &a.logtab[i >> 3]);
    }
    const double2 *logtab = logtab8 + (tid & 7);                  // this lane's copy
#pragma unroll
    for (int r = 0; r < (16 * 15 + kStft1920Threads - 1) / kStft1920Threads; ++r) {
        const int i = tid + r * kStft1920Threads;
        if (i < 16 * 15) copy16(&tw2s[i], &a.tw2[i]);
    }
    const int64_t n_items = a.n_slots * a.frames_out;
    const int Fout = a.bin_hi - a.bin_lo;
    const double scale4 = 0.25 * a.scale;
    const bool full = (a.bin_lo == 0) && (a.bin_hi >= kM1920);   // bins 0..1919 all stored: no per-bin range checks
    auto frame_audio = [&](int64_t it) -> const WaveT * {
        const int64_t sl = it / a.frames_out;
        return reinterpret_cast<const WaveT *>(a.wave) + sl * a.slot_stride + (it - sl * a.frames_out + a.frame_lo) * a.hop;
    };
#ifndef FT8B_STFT_STAGE16
#define FT8B_STFT_STAGE16 1
#endif
#if FT8B_STFT_STAGE16
    // The next frame's 1920 samples are copied into `xs` with cp.async right after this frame's P1 has read the buffer,
    // and waited for just before the barrier that separates P2 from P3 - a whole radix-16 pass later, so the wait costs
    // nothing, and the barrier that is there anyway makes every thread's copies visible to all.  P1 of the next frame
    // then starts without a wait.  Frames that start on a 16-byte boundary are copied 16 bytes at a time with
    // cp.async.cg (no L1 allocation: the audio stream does not evict the window and twiddle tables); others pair by pair.
    auto stage_audio = [&](int64_t it) {
        if (VEC && it < n_items) {
            const char *src = reinterpret_cast<const char *>(frame_audio(it));
            if (a.copy16) {
                constexpr int n16 = (int)sizeof(WaveT) * kM1920 / 16;
#pragma unroll
                for (int t = 0; t < (n16 + 127) / 128; ++t) {
                    const int c = tid + 128 * t;
                    if (c < n16)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(reinterpret_cast<char *>(xs) + 16 * c)),
                                     "l"(src + 16 * c) : "memory");
                }
            } else {
                constexpr int kPair = 2 * (int)sizeof(WaveT);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int n = tid + 128 * t;
                    if (t < 7 || tid < 64) __pipeline_memcpy_async(reinterpret_cast<char *>(xs) + kPair * n, src + kPair * n, kPair);
                }
            }
        }
        __pipeline_commit();
    };
    stage_audio(blockIdx.x);
    __pipeline_wait_prior(0);           // tables + (VEC) the first frame's audio
    __syncthreads();
#else
    // Each thread copies exactly the sample pairs it will consume in P1 (n = tid + 128 t), so waiting on its own
    // cp.async groups is enough - no block barrier guards the staging buffer.
    auto stage_audio = [&](int64_t it) {
        if (VEC && it < n_items) {
            constexpr int kPair = 2 * (int)sizeof(WaveT);
            const char *src = reinterpret_cast<const char *>(frame_audio(it));
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int n = tid + 128 * t;
                if (t < 7 || tid < 64) __pipeline_memcpy_async(reinterpret_cast<char *>(xs) + kPair * n, src + kPair * n, kPair);
            }
        }
        __pipeline_commit();
    };
    stage_audio(blockIdx.x);
    __pipeline_wait_prior(0);           // the tables (this thread's first audio pairs ride along)
    __syncthreads();
#endif
    FT8B_PROF_DECL
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t slot = item / a.frames_out;
        const int64_t fr = item - slot * a.frames_out;
        const WaveT *x = VEC ? xs : frame_audio(item);
        double *out = a.mag + (slot * a.frames_out + fr) * (int64_t)Fout;
#if !FT8B_STFT_STAGE16
        if (VEC) __pipeline_wait_prior(0);     // this thread's pairs of this frame have landed
#endif
        FT8B_PROF(0);
        // ---------------- P1: radix 15, inputs z[j + 128 t] = (w x)[2n] + i (w x)[2n+1], zero for n >= 960
        {
            double2 u[15];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int n = tid + 128 * t;
                if (t < 7 || tid < 64) {
                    const double2 w = __ldg(reinterpret_cast<const double2 *>(a.win) + n);
                    if (VEC) {
                        const auto xv = load_pair(x, n);
                        u[t] = make_double2(w.x * xv.x, w.y * xv.y);
                    } else {
                        u[t] = make_double2(w.x * to_sample(x[2 * n]), w.y * to_sample(x[2 * n + 1]));
                    }
                } else {
                    u[t] = make_double2(0.0, 0.0);
                }
            }
            bf15<true>(u);                     // u[8..14] (the zero padding) are outputs only
            FT8B_PROF(1);
            // Barrier here rather than at the end of the loop: warps that finished the previous frame's P3 early
            // have already loaded and transformed their share of this frame; `s` (still read by slower warps in
            // P3) and the audio staging buffer are only touched after everybody has arrived.
            __syncthreads();
            FT8B_PROF(2);
#pragma unroll
            for (int pos = 0; pos < 15; ++pos) s[tid * 15 + out15(pos)] = u[pos];
        }
        stage_audio(item + gridDim.x);         // lands during P2/P3
        FT8B_PROF(1);
        __syncthreads();
        FT8B_PROF(2);
        // ---------------- P2: radix 16, Ns = 15
        // Butterfly j = 15 g + k runs on lane k of 16-lane group g (lane 15 of each group idles): the 8 lanes of
        // every quarter-warp then touch consecutive k, so the tw2s reads and the transposed stores below are
        // free of bank conflicts, and all four warps carry the same 30 butterflies.
        {
            double2 u[16];
            const int g = tid >> 4, k = tid & 15;
            const int j = 15 * g + k;
            if (k < 15) {
#pragma unroll
                for (int t = 0; t < 16; ++t) u[t] = s[j + 120 * t];
#pragma unroll
                for (int t = 1; t < 16; ++t) u[t] = cmul(u[t], tw2s[t * 15 + k]);
                bf16(u);
            }
            FT8B_PROF(3);
            __syncthreads();
            FT8B_PROF(4);
            if (k < 15) {
                const int j0 = 240 * g + k;
#pragma unroll
                for (int pos = 0; pos < 16; ++pos) s[j0 + 15 * out16(pos)] = u[pos];
            }
        }
        FT8B_PROF(3);
#if FT8B_STFT_STAGE16
        if (VEC) __pipeline_wait_prior(0);     // the next frame's audio (staged after P1) has landed: visible to all after the barrier
#endif
        __syncthreads();
        FT8B_PROF(4);
        // ---------------- P3: radix 8, Ns = 240, paired butterflies + untangling + dB store
        // Thread i in 0..118 owns butterflies (i+1, 239-i): outputs u1[t] = Z[k], k = i+1 + 240 t, and Z[M-k] from the
        // second.  Its twiddles are W^(t (240 - j1)) = W8^t conj(W^(t j1)): the conjugates of the first butterfly's
        // (one table load serves both, and only columns 0..120 of tw3 are ever touched), while the factor W8^t just
        // rotates the radix-8 outputs by one - so Z[M-k] = u2[(8-t) & 7] with no arithmetic for it.
        // Butterflies 0 and 120 pair with themselves (Z[240 t] with Z[240 ((8-t) & 7)], Z[120 + 240 t] with
        // Z[120 + 240 (7-t)]) and obey the same formulas with j1 = 0 / 120, so threads 119 and 120 run them through the
        // very same instruction stream.  Each of them produces all 8 bins of its butterfly as the "k" output; of the
        // mirrored outputs only the Nyquist bin (M - 0) is stored.
        if (tid <= 120) {
            double2 u1[8], u2[8];
            const bool mirror = tid < 119;                                     // the M-k outputs are ours to store
            const bool nyquist = tid == 119 && a.bin_hi > kM1920;              // ... or just bin M
            const int j1 = mirror ? tid + 1 : (tid == 119 ? 0 : 120);
            const int j2 = (tid == 119) ? 0 : 240 - j1;                        // row of `s` of the second butterfly
#pragma unroll
            for (int t = 0; t < 8; ++t) { u1[t] = s[j1 + 240 * t]; u2[t] = s[j2 + 240 * t]; }
#pragma unroll
            for (int t = 1; t < 8; ++t) {
#if FT8B_STFT_SMEMTAB
                const double2 w = tw3s[(t - 1) * kTw3Cols + j1];
#else
                const double2 w = __ldg(&a.tw3[t * kTw3Stride + j1]);
#endif
                u1[t] = cmul(u1[t], w);
                u2[t] = make_double2(u2[t].x * w.x + u2[t].y * w.y, u2[t].y * w.x - u2[t].x * w.y);   // * conj w
            }
            bf8(u1);
            bf8(u2);
            FT8B_PROF(5);
            if (full) {
                // bins 0..1919 are all stored: k = j1 + 240 t ascending from p1, M-k descending from p2,
                // every address an immediate offset from three per-thread base pointers
                const double2 *twp = a.twpost + j1;
                double *p1 = out + j1, *p2 = out + (kM1920 - j1);
                // Both outputs are stored unconditionally (a conditional store would fence each pair into its own
                // basic block and serialise the log chains): threads 119/120 aim the mirrored value at the slot
                // the same iteration overwrites with d1 right after.
                double *const m0 = (mirror || nyquist) ? p2 : p1, *const mt = mirror ? p2 : p1;
                const int mstep = mirror ? -240 : 240;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    // W3840^(k + 960) = -i W3840^k: bins k = j1 + 240 h and k + 960 share one post-twiddle load
#if FT8B_STFT_SMEMTAB
                    // k = j1 + 240 h; beyond 480 the table is mirrored: W^k = -i conj(W^(960-k)) = (-y, -x) of entry 960-k
                    // (get_plan builds entries 480..960 of the global table by this very rule, so both forms agree bit for bit)
                    double2 w;
                    if (h < 2) {
                        w = twps[j1 + 240 * h];
                    } else {
                        const double2 m = twps[960 - 240 * h - j1];
                        w = make_double2(-m.y, -m.x);
                    }
#else
                    const double2 w = __ldg(twp + 240 * h);
#endif
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int t = h + 4 * q;
                        double d1, d2;
                        untangle_pair(u1[t], u2[(8 - t) & 7], q ? mul_mi(w) : w, scale4, logtab, d1, d2);
                        if (t == 0) *m0 = d2; else mt[t * mstep] = d2;
                        p1[240 * t] = d1;
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) emit_pair(u1[t], u2[(8 - t) & 7], j1 + 240 * t, mirror || (t == 0 && nyquist), a, out, logtab);
            }
        }
        FT8B_PROF(6);
    }
    FT8B_PROF_END;
}

}  // namespace ft8b
