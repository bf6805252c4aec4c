/*
 * ft8_oracle.c - CPU restatement of the reference FT8 demodulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product package (ft8_demodulator_b200/) never links, imports or calls it.
 *
 * What it restates (all float64, the dtype the reference computes in for float64 audio):
 *   reference = Rintazero/ft8_demodulator, paths relative to src/ft8_tools/ft8_demodulator/
 *     spectrogram_analyse.py:19-66   calculate_spectrogram  (-> scipy.signal.spectrogram)
 *     ft8_decode.py:47-100           ft8_sync_score
 *     ft8_decode.py:102-149          ft8_find_candidates   (heapq semantics, see orc_find_candidates)
 *     ft8_decode.py:151-198          ft8_extract_symbol / ft8_extract_likelihood / ftx_normalize_logl
 *     ft8_decode.py:200-273          pack_bits / ft8_decode_candidate
 *     ft8_decode.py:288-394          decode_ft8_message    (masks + orchestration; no plotting/printing)
 *     ldpc_decoder.py:11-113         fast_tanh / fast_atanh / ldpc_check / bp_decode
 *     crc.py:11-54                   compute_crc / extract_crc
 *     ftx_types.py:14-47             FT8Waterfall.num_blocks / FT8Candidate.get_log_power
 *
 * The STFT arithmetic itself is a third-party dependency of the reference
 * (scipy.signal.spectrogram, scipy>=1.7 per requirements.txt:2; scipy 1.18.1 in the build
 * image, legacy _spectral_helper path, pocketfft).  Its published algorithm is restated in
 * orc_spectrogram_pos(): periodic Hann window, frames of nperseg samples every hop samples,
 * no padding/detrend, zero-padded nfft-point DFT, |X|^2 / (sum w)^2 ('spectrum' scaling,
 * two-sided so no doubling), then the reference's own 10*log10(1e-12 + |.|).
 *
 * Parity pin: tests/test_oracle_vs_reference.py (runs where /root/reference exists) and
 * tests/test_oracle_golden.py (committed .npz fixtures under tests/golden/, generated from the
 * unmodified reference by oracle/make_golden.py).  Stages downstream of the spectrogram are
 * bit-exact against the reference given the same spectrogram; the spectrogram agrees to
 * ~1e-11 dB (different FFT factorisation from pocketfft).
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off: NO fused multiply-add, numpy has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ldpc_tables.h"

#define ORC_LDPC_N 174
#define ORC_LDPC_M 83
#define ORC_LDPC_K 91
#define ORC_ND 58          /* ft8_decode.py:30 */
#define ORC_NUM_SYNC 3     /* :31 */
#define ORC_LEN_SYNC 7     /* :32 */
#define ORC_SYNC_OFFSET 36 /* :33 */

static const int ORC_COSTAS[7] = {3, 1, 4, 0, 6, 5, 2}; /* ft8_decode.py:42 */
static const int ORC_GRAY[8] = {0, 1, 3, 2, 5, 6, 4, 7}; /* ft8_decode.py:39 */

#define ORC_ERR_TIE (-2)   /* reference would raise TypeError (heap tuple tie) */
#define ORC_ERR_ARG (-1)

/* ------------------------------------------------------------------------------------ */
/* Window / frame geometry: spectrogram_analyse.py:32-43                                  */
/* ------------------------------------------------------------------------------------ */
int orc_plan(double fs, int bins_per_tone, int steps_per_symbol, int *nperseg, int *hop, int *nfft) {
    int sps = (int)(0.16 * fs);                         /* :32 int(FT8_SYMBOL_DURATION_S*fs) */
    int overlap = sps - sps / steps_per_symbol;         /* :33 */
    int n = (int)(fs / 6.25 * bins_per_tone);           /* :34 */
    if (overlap >= sps) overlap = sps - 1;              /* :42-43 */
    *nperseg = sps;
    *hop = sps - overlap;
    *nfft = n;
    return 0;
}

long orc_num_frames(long n, int nperseg, int hop) {
    if (n < nperseg) return 0;                          /* :37-39 */
    return (n - nperseg) / hop + 1;                     /* scipy _fft_helper: (N-noverlap)//step */
}

/* ------------------------------------------------------------------------------------ */
/* Mixed-radix DFT (any length).  Recursive decimation in time by the smallest prime      */
/* factor; generic O(p^2) butterflies.  Twiddles from one table of exp(-2*pi*i*q/N)       */
/* evaluated in long double.                                                              */
/* ------------------------------------------------------------------------------------ */
typedef struct { double re, im; } cplx;

static int smallest_factor(int n) {
    if (n % 2 == 0) return 2;
    for (int p = 3; (long)p * p <= n; p += 2)
        if (n % p == 0) return p;
    return n;
}

static void dft_rec(const cplx *in, int stride, cplx *out, int n, const cplx *tw, int N, cplx *tmp) {
    if (n == 1) { out[0] = in[0]; return; }
    int p = smallest_factor(n);
    int m = n / p;
    for (int q = 0; q < p; q++) dft_rec(in + (long)q * stride, stride * p, out + (long)q * m, m, tw, N, tmp);
    int step_n = N / n;  /* W_n^a = tw[a*step_n] */
    int step_p = N / p;
    for (int k = 0; k < m; k++) {
        for (int q = 0; q < p; q++) {
            cplx v = out[(long)q * m + k];
            cplx w = tw[(int)(((long)q * k % n) * step_n)];
            tmp[q].re = v.re * w.re - v.im * w.im;
            tmp[q].im = v.re * w.im + v.im * w.re;
        }
        for (int r = 0; r < p; r++) {
            double sr = 0.0, si = 0.0;
            for (int q = 0; q < p; q++) {
                cplx w = tw[(int)(((long)q * r % p) * step_p)];
                sr += tmp[q].re * w.re - tmp[q].im * w.im;
                si += tmp[q].re * w.im + tmp[q].im * w.re;
            }
            out[(long)r * m + k].re = sr;
            out[(long)r * m + k].im = si;
        }
    }
}

/*
 * dB spectrogram, non-negative frequency rows only (what decode_ft8_message keeps at
 * ft8_decode.py:323-325): rows k = 0 .. ceil(nfft/2)-1, layout mag[k*T + j] (freq, time)
 * exactly like the reference's `spectrogram[positive_freq_mask]`.
 *   spectrogram_analyse.py:46-56 (scipy call) and :59-60 (10*log10(1e-12+|S|)).
 */
static int spectrogram_pos_impl(const double *x, int cplx_in, long n, int nperseg, int hop, int nfft, double *mag) {
    long T = orc_num_frames(n, nperseg, hop);
    if (T <= 0 || nfft < nperseg) return ORC_ERR_ARG;
    int F = (nfft + 1) / 2;
    cplx *tw = (cplx *)malloc(sizeof(cplx) * nfft);
    cplx *buf = (cplx *)malloc(sizeof(cplx) * nfft);
    cplx *out = (cplx *)malloc(sizeof(cplx) * nfft);
    cplx *tmp = (cplx *)malloc(sizeof(cplx) * nfft);
    double *w = (double *)malloc(sizeof(double) * nperseg);
    if (!tw || !buf || !out || !tmp || !w) return ORC_ERR_ARG;
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int q = 0; q < nfft; q++) {
        long double a = -two_pi * (long double)q / (long double)nfft;
        tw[q].re = (double)cosl(a);
        tw[q].im = (double)sinl(a);
    }
    double wsum = 0.0;
    for (int i = 0; i < nperseg; i++) {  /* scipy get_window('hann', nperseg): periodic */
        w[i] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)nperseg);
        wsum += w[i];
    }
    double scale = 1.0 / (wsum * wsum);  /* scaling='spectrum' */
    for (long j = 0; j < T; j++) {
        if (cplx_in) {   /* complex audio, interleaved (re, im): same scipy call, two-sided either way */
            const double *fr = x + 2 * j * hop;
            for (int i = 0; i < nperseg; i++) { buf[i].re = w[i] * fr[2 * i]; buf[i].im = w[i] * fr[2 * i + 1]; }
        } else {
            const double *fr = x + j * hop;
            for (int i = 0; i < nperseg; i++) { buf[i].re = w[i] * fr[i]; buf[i].im = 0.0; }
        }
        for (int i = nperseg; i < nfft; i++) { buf[i].re = 0.0; buf[i].im = 0.0; }
        dft_rec(buf, 1, out, nfft, tw, nfft, tmp);
        for (int k = 0; k < F; k++) {
            double p = (out[k].re * out[k].re + out[k].im * out[k].im) * scale;
            mag[(long)k * T + j] = 10.0 * log10(1e-12 + fabs(p));
        }
    }
    free(tw); free(buf); free(out); free(tmp); free(w);
    return 0;
}

int orc_spectrogram_pos(const double *x, long n, int nperseg, int hop, int nfft, double *mag) {
    return spectrogram_pos_impl(x, 0, n, nperseg, hop, nfft, mag);
}
/* complex audio x[2n] = (re, im) pairs: the rows f >= 0 of the two-sided spectrogram (ft8_decode.py:323-325) */
int orc_spectrogram_pos_c(const double *x, long n, int nperseg, int hop, int nfft, double *mag) {
    return spectrogram_pos_impl(x, 1, n, nperseg, hop, nfft, mag);
}

/* ------------------------------------------------------------------------------------ */
/* Waterfall access: ftx_types.py:31-34 (num_blocks) and :45-47 (get_log_power)           */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const double *mag; /* [F][T] */
    int F, T, tosr, fosr;
} orc_wf;

static inline int wf_num_blocks(const orc_wf *wf) { return wf->T / wf->tosr; }

static inline double wf_power(const orc_wf *wf, int abs_time, int abs_freq, int time_off, int freq_off) {
    long r = (long)abs_freq + (long)freq_off * wf->fosr;
    long c = (long)abs_time + (long)time_off * wf->tosr;
    /* numpy negative indexing wraps; mirror it so behaviour matches for in-range |index| */
    if (r < 0) r += wf->F;
    if (c < 0) c += wf->T;
    return wf->mag[r * wf->T + c];
}

static inline int floordiv(int a, int b) { /* Python // for b > 0 */
    int q = a / b;
    if ((a % b) != 0 && (a < 0)) q -= 1;
    return q;
}

/* ft8_decode.py:47-100 */
double orc_sync_score_wf(const orc_wf *wf, int abs_time, int abs_freq) {
    double score = 0.0;
    int num = 0;
    int base = floordiv(abs_time, wf->tosr);  /* :59 */
    int nb = wf_num_blocks(wf);
    for (int m = 0; m < ORC_NUM_SYNC; m++) {
        int start = m * ORC_SYNC_OFFSET;
        for (int k = 0; k < ORC_LEN_SYNC; k++) {
            int block = start + k;
            int block_abs = base + block;
            if (block_abs < 0 || block_abs >= nb) continue;      /* :71-72 */
            int tone = ORC_COSTAS[k];
            double p = wf_power(wf, abs_time, abs_freq, block, tone);
            if (tone > 0) { score += p - wf_power(wf, abs_time, abs_freq, block, tone - 1); num++; }   /* :79-81 */
            if (tone < 7) { score += p - wf_power(wf, abs_time, abs_freq, block, tone + 1); num++; }   /* :83-85 */
            if (k > 0 && block_abs > 0) { score += p - wf_power(wf, abs_time, abs_freq, block - 1, tone); num++; } /* :88-90 */
            if (k < ORC_LEN_SYNC - 1 && block_abs + 1 < nb) { score += p - wf_power(wf, abs_time, abs_freq, block + 1, tone); num++; } /* :92-94 */
        }
    }
    if (num == 0 || isnan(score) || isinf(score)) return -INFINITY;  /* :97-98 */
    return score / (double)num;                                       /* :100 */
}

double orc_sync_score(const double *mag, int F, int T, int tosr, int fosr, int abs_time, int abs_freq) {
    orc_wf wf = {mag, F, T, tosr, fosr};
    return orc_sync_score_wf(&wf, abs_time, abs_freq);
}

/* Search ranges: ft8_decode.py:108-109 */
void orc_search_ranges(int F, int T, int tosr, int fosr, int *t0, int *t1, int *f0, int *f1) {
    int nb = T / tosr;
    *t0 = -10 * tosr;
    *t1 = nb * tosr - tosr * (ORC_ND + 1);
    *f0 = 0;
    *f1 = F - 7 * fosr;
}

/* All scores in scan order, [nt][nf] (test aid for the GPU score map). */
int orc_score_map(const double *mag, int F, int T, int tosr, int fosr, double *scores) {
    orc_wf wf = {mag, F, T, tosr, fosr};
    int t0, t1, f0, f1;
    orc_search_ranges(F, T, tosr, fosr, &t0, &t1, &f0, &f1);
    long i = 0;
    for (int t = t0; t < t1; t++)
        for (int f = f0; f < f1; f++) scores[i++] = orc_sync_score_wf(&wf, t, f);
    return 0;
}

/* ---- heapq on (key=-score, cand); Python's heapq._siftdown/_siftup, heapq.py ---- */
typedef struct { double key; int t, f; double score; } hitem;

static int h_lt(const hitem *a, const hitem *b, int *tie) {
    if (a->key == b->key) { *tie = 1; return 0; } /* tuple compare falls through to FT8Candidate -> TypeError */
    return a->key < b->key;
}
static void h_siftdown(hitem *h, int startpos, int pos, int *tie) {
    hitem item = h[pos];
    while (pos > startpos) {
        int parent = (pos - 1) >> 1;
        if (h_lt(&item, &h[parent], tie)) { h[pos] = h[parent]; pos = parent; continue; }
        break;
    }
    h[pos] = item;
}
static void h_siftup(hitem *h, int n, int pos, int *tie) {
    int startpos = pos;
    hitem item = h[pos];
    int child = 2 * pos + 1;
    while (child < n) {
        int right = child + 1;
        if (right < n && !h_lt(&h[child], &h[right], tie)) child = right;
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    h[pos] = item;
    h_siftdown(h, startpos, pos, tie);
}

/*
 * ft8_decode.py:102-149.  The heap holds (-score, cand), so h[0] is the BEST candidate and
 * `elif -score < candidates[0][0]: heapreplace` (:136-137) evicts the current best whenever a
 * new record shows up after the first K passers.  Reproduced literally.
 * Returns the number of candidates (<= K), ORC_ERR_TIE if an exact key tie was compared.
 */
int orc_find_candidates(const double *mag, int F, int T, int tosr, int fosr, int K, double min_score,
                        int *out_t, int *out_f, double *out_score, long *n_pass) {
    orc_wf wf = {mag, F, T, tosr, fosr};
    int t0, t1, f0, f1;
    orc_search_ranges(F, T, tosr, fosr, &t0, &t1, &f0, &f1);
    if (K < 0) return ORC_ERR_ARG;
    hitem *h = (hitem *)malloc(sizeof(hitem) * (K > 0 ? K : 1));
    int n = 0, tie = 0;
    long passed = 0;
    for (int t = t0; t < t1; t++) {
        for (int f = f0; f < f1; f++) {
            double s = orc_sync_score_wf(&wf, t, f);
            if (s == -INFINITY || s < min_score) continue;   /* :127-128 */
            passed++;
            hitem it = {-s, t, f, s};
            if (n < K) {                                      /* :134-135 heappush */
                h[n] = it;
                n++;
                h_siftdown(h, 0, n - 1, &tie);
            } else if (n > 0 && -s < h[0].key) {              /* :136-137 heapreplace */
                h[0] = it;
                h_siftup(h, n, 0, &tie);
            }
            /* K == 0: Python indexes candidates[0] on an empty list -> IndexError; we return none */
        }
    }
    /* :140 sorted(candidates, key=lambda x: x[0]) - stable insertion sort on the heap array */
    for (int i = 1; i < n; i++) {
        hitem it = h[i];
        int j = i - 1;
        while (j >= 0 && h[j].key > it.key) { h[j + 1] = h[j]; j--; }
        h[j + 1] = it;
    }
    for (int i = 0; i < n; i++) { out_t[i] = h[i].t; out_f[i] = h[i].f; out_score[i] = h[i].score; }
    free(h);
    if (n_pass) *n_pass = passed;
    return tie ? ORC_ERR_TIE : n;
}

/* ------------------------------------------------------------------------------------ */
/* LLR extraction: ft8_decode.py:151-188                                                  */
/* ------------------------------------------------------------------------------------ */
static inline double max4(double a, double b, double c, double d) {
    /* Python max(a,b,c,d): keeps the first maximal element */
    double m = a;
    if (b > m) m = b;
    if (c > m) m = c;
    if (d > m) m = d;
    return m;
}

void orc_extract_likelihood(const double *mag, int F, int T, int tosr, int fosr, int abs_time, int abs_freq,
                            double *log174) {
    orc_wf wf = {mag, F, T, tosr, fosr};
    int base = floordiv(abs_time, tosr);   /* :168 */
    int nb = wf_num_blocks(&wf);
    for (int k = 0; k < ORC_ND; k++) {
        int sym = k + (k < 29 ? 7 : 14);   /* :173 */
        int bit = 3 * k;
        int block = base + sym;
        if (block < 0 || block >= nb) {    /* :178-181 */
            log174[bit] = log174[bit + 1] = log174[bit + 2] = 0.0;
        } else {
            double s2[8];
            for (int j = 0; j < 8; j++) s2[j] = wf_power(&wf, abs_time, abs_freq, sym, ORC_GRAY[j]); /* :156-157,:185-186 */
            log174[bit + 0] = max4(s2[4], s2[5], s2[6], s2[7]) - max4(s2[0], s2[1], s2[2], s2[3]); /* :159 */
            log174[bit + 1] = max4(s2[2], s2[3], s2[6], s2[7]) - max4(s2[0], s2[1], s2[4], s2[5]); /* :160 */
            log174[bit + 2] = max4(s2[1], s2[3], s2[5], s2[7]) - max4(s2[0], s2[2], s2[4], s2[6]); /* :161 */
        }
    }
}

/* numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum),
 * which np.mean -> np.add.reduce uses for a contiguous float64 vector. */
static double np_pairwise_sum(const double *a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* ft8_decode.py:190-198 */
void orc_normalize_logl(double *log174) {
    double tmp[ORC_LDPC_N];
    double mean = np_pairwise_sum(log174, ORC_LDPC_N) / (double)ORC_LDPC_N;      /* np.mean :193 */
    for (int i = 0; i < ORC_LDPC_N; i++) { double d = log174[i] - mean; tmp[i] = d * d; }
    double variance = np_pairwise_sum(tmp, ORC_LDPC_N) / (double)ORC_LDPC_N;     /* :194 */
    double norm = sqrt(24.0 / variance);                                         /* :197 */
    for (int i = 0; i < ORC_LDPC_N; i++) log174[i] *= norm;                      /* :198 */
}

/* ------------------------------------------------------------------------------------ */
/* LDPC: ldpc_decoder.py                                                                  */
/* ------------------------------------------------------------------------------------ */
static inline double orc_fast_tanh(double x) {   /* :11-21 */
    /* np.clip(x, -4.97, 4.97): minimum(maximum(x, lo), hi); NaN propagates */
    if (x < -4.97) x = -4.97;
    if (x > 4.97) x = 4.97;
    double x2 = x * x;
    double a = x * (945.0 + x2 * (105.0 + x2));
    double b = 945.0 + x2 * (420.0 + x2 * 15.0);
    return a / b;
}
static inline double orc_fast_atanh(double x) {  /* :23-31 */
    double x2 = x * x;
    double a = x * (945.0 + x2 * (-735.0 + x2 * 64.0));
    double b = (945.0 + x2 * (-1050.0 + x2 * 225.0));
    return a / b;
}

int orc_ldpc_check(const uint8_t *cw) {           /* :33-52 */
    int errors = 0;
    for (int m = 0; m < ORC_LDPC_M; m++) {
        uint8_t x = 0;
        for (int i = 0; i < ORC_NUM_ROWS[m]; i++) x ^= cw[ORC_NM[m * 7 + i] - 1];
        if (x != 0) errors++;
    }
    return errors;
}

int orc_bp_decode(const double *codeword, int max_iterations, uint8_t *plain) {  /* :54-113 */
    double tov[ORC_LDPC_N][3];
    double toc[ORC_LDPC_M][7];
    memset(tov, 0, sizeof(tov));
    memset(toc, 0, sizeof(toc));
    memset(plain, 0, ORC_LDPC_N);
    int min_errors = ORC_LDPC_M;
    for (int iter = 0; iter < max_iterations; iter++) {
        int plain_sum = 0;
        for (int n = 0; n < ORC_LDPC_N; n++) {
            double msg = codeword[n] + ((tov[n][0] + tov[n][1]) + tov[n][2]);   /* :73 np.sum(tov,axis=1) */
            plain[n] = (msg > 0) ? 1 : 0;                                       /* :74 */
            plain_sum += plain[n];
        }
        if (plain_sum == 0) break;                                              /* :76-78 */
        int errors = orc_ldpc_check(plain);                                     /* :81 */
        if (errors < min_errors) {
            min_errors = errors;
            if (errors == 0) break;                                             /* :83-86 */
        }
        for (int m = 0; m < ORC_LDPC_M; m++) {                                  /* :89-97 */
            for (int ni = 0; ni < ORC_NUM_ROWS[m]; ni++) {
                int n = ORC_NM[m * 7 + ni] - 1;
                double Tnm = codeword[n];
                for (int mi = 0; mi < 3; mi++)
                    if ((ORC_MN[n * 3 + mi] - 1) != m) Tnm += tov[n][mi];
                toc[m][ni] = orc_fast_tanh(-Tnm / 2);
            }
        }
        for (int n = 0; n < ORC_LDPC_N; n++) {                                  /* :100-108 */
            for (int mi = 0; mi < 3; mi++) {
                int m = ORC_MN[n * 3 + mi] - 1;
                double Tmn = 1.0;
                for (int ni = 0; ni < ORC_NUM_ROWS[m]; ni++)
                    if ((ORC_NM[m * 7 + ni] - 1) != n) Tmn *= toc[m][ni];
                tov[n][mi] = -2 * orc_fast_atanh(Tmn);
            }
        }
    }
    return min_errors;
}

/* ------------------------------------------------------------------------------------ */
/* CRC-14 and bit packing: crc.py:11-54, ft8_decode.py:200-215                            */
/* ------------------------------------------------------------------------------------ */
unsigned orc_compute_crc(const uint8_t *message, int num_bits) {   /* crc.py:11-39 */
    const unsigned TOPBIT = 1u << 13;
    unsigned long remainder = 0;   /* Python ints do not wrap: keep wide, mask at the end */
    int idx_byte = 0;
    for (int idx_bit = 0; idx_bit < num_bits; idx_bit++) {
        if (idx_bit % 8 == 0) { remainder ^= ((unsigned long)message[idx_byte] << (14 - 8)); idx_byte++; }
        if (remainder & TOPBIT) remainder = (remainder << 1) ^ 0x2757u;
        else remainder = (remainder << 1);
        remainder &= 0xFFFFFFFFFFFFul; /* bits above 13 never feed back into bits <= 13; bound the width */
    }
    return (unsigned)(remainder & ((TOPBIT << 1) - 1));
}
unsigned orc_extract_crc(const uint8_t *a91) {                     /* crc.py:41-54 */
    return ((unsigned)(a91[9] & 0x07) << 11) | ((unsigned)a91[10] << 3) | ((unsigned)a91[11] >> 5);
}
void orc_pack_bits(const uint8_t *bits, int num_bits, uint8_t *packed) {   /* ft8_decode.py:200-215 */
    int nbytes = (num_bits + 7) / 8;
    memset(packed, 0, nbytes);
    unsigned mask = 0x80;
    int b = 0;
    for (int i = 0; i < num_bits; i++) {
        if (bits[i]) packed[b] |= mask;
        mask >>= 1;
        if (!mask) { mask = 0x80; b++; }
    }
}

/* Per-candidate record shared with the Python wrapper (ctypes.Structure, same field order). */
typedef struct {
    int32_t abs_time, abs_freq;
    double score;
    int32_t ok;             /* 1 = LDPC converged and CRC matched (ft8_decode.py:273) */
    int32_t ldpc_errors;    /* FT8DecodeStatus.ldpc_errors */
    int32_t crc_extracted;  /* 0 when LDPC failed (early return, :239-240) */
    int32_t crc_calculated;
    uint8_t payload[10];    /* FT8Message.payload; zeros unless ok */
    uint8_t plain[ORC_LDPC_N]; /* hard decisions returned by bp_decode */
    uint8_t _pad[8];
    double llr[ORC_LDPC_N]; /* normalised log174 fed to bp_decode */
} orc_record;

/* ft8_decode.py:225-273 */
int orc_decode_candidate(const double *mag, int F, int T, int tosr, int fosr, int abs_time, int abs_freq,
                         int max_iterations, orc_record *r) {
    r->abs_time = abs_time;
    r->abs_freq = abs_freq;
    r->ok = 0;
    r->crc_extracted = r->crc_calculated = 0;
    memset(r->payload, 0, 10);
    orc_extract_likelihood(mag, F, T, tosr, fosr, abs_time, abs_freq, r->llr);   /* :231 */
    orc_normalize_logl(r->llr);                                                  /* :233 */
    r->ldpc_errors = orc_bp_decode(r->llr, max_iterations, r->plain);            /* :236 */
    if (r->ldpc_errors > 0) return 0;                                            /* :239-240 */
    uint8_t a91[12], buf[12];
    orc_pack_bits(r->plain, ORC_LDPC_K, a91);                                    /* :243 */
    r->crc_extracted = (int32_t)orc_extract_crc(a91);                            /* :246 */
    memcpy(buf, a91, 10);
    buf[9] &= 0xF8; buf[10] = 0; buf[11] = 0;                                    /* :255-257 */
    r->crc_calculated = (int32_t)orc_compute_crc(buf, 82);                       /* :260 */
    if (r->crc_extracted != r->crc_calculated) return 0;                         /* :262-263 */
    memcpy(r->payload, a91, 10);
    r->payload[9] &= 0xF8;                                                       /* :269-271 */
    r->ok = 1;
    return 1;
}

/* ------------------------------------------------------------------------------------ */
/* decode_ft8_message: ft8_decode.py:288-394 (NaN for a limit means None)                 */
/* ------------------------------------------------------------------------------------ */
/* Row/column selection of ft8_decode.py:323-341.  Returns the contiguous [lo,hi) ranges. */
int orc_masks(double fs, int nperseg, int hop, int nfft, long T, double freq_min, double freq_max,
              double time_min, double time_max, int *k_lo, int *k_hi, long *j_lo, long *j_hi) {
    int Fpos = (nfft + 1) / 2;
    double d = 1.0 / fs;                 /* scipy: fftfreq(nfft, 1/fs) */
    double val = 1.0 / (nfft * d);       /* numpy fftfreq: results * val */
    int lo = 0, hi = Fpos;
    if (!isnan(freq_min) || !isnan(freq_max)) {
        double fmin = isnan(freq_min) ? 0 * val : freq_min;
        double fmax = isnan(freq_max) ? (Fpos - 1) * val : freq_max;
        lo = Fpos; hi = 0;
        for (int k = 0; k < Fpos; k++) {
            double f = k * val;
            if (f >= fmin && f <= fmax) { if (k < lo) lo = k; if (k + 1 > hi) hi = k + 1; }
        }
        if (hi < lo) { lo = 0; hi = 0; }
    }
    long jl = 0, jh = T;
    if (!isnan(time_min) || !isnan(time_max)) {
        double start = nperseg / 2.0;    /* scipy: arange(nperseg/2, N-nperseg/2+1, step)/fs */
        double t_first = (start + 0 * (double)hop) / fs;
        double t_last = (start + (double)(T - 1) * (double)hop) / fs;
        double tmin = isnan(time_min) ? t_first : time_min;
        double tmax = isnan(time_max) ? t_last : time_max;
        jl = T; jh = 0;
        for (long j = 0; j < T; j++) {
            double t = (start + (double)j * (double)hop) / fs;
            if (t >= tmin && t <= tmax) { if (j < jl) jl = j; if (j + 1 > jh) jh = j + 1; }
        }
        if (jh < jl) { jl = 0; jh = 0; }
    }
    *k_lo = lo; *k_hi = hi; *j_lo = jl; *j_hi = jh;
    return 0;
}

/*
 * Full decode of one waveform.  Writes up to max_records records for ALL candidates in
 * candidate (score-descending) order; the reference's return list is the subset with ok==1.
 * Returns the number of candidates, or a negative error.  T_out and F_out report the waterfall shape.
 */
static int decode_message_impl(const double *wave, int cplx_in, long n, double fs, int bins_per_tone, int steps_per_symbol,
                               int max_candidates, double min_score, int max_iterations, double freq_min,
                               double freq_max, double time_min, double time_max, orc_record *records, int max_records,
                               int *F_out, int *T_out) {
    int nperseg, hop, nfft;
    orc_plan(fs, bins_per_tone, steps_per_symbol, &nperseg, &hop, &nfft);
    long T = orc_num_frames(n, nperseg, hop);
    if (T <= 0) return ORC_ERR_ARG;   /* reference: IndexError at ft8_decode.py:346 */
    int Fpos = (nfft + 1) / 2;
    double *full = (double *)malloc(sizeof(double) * (size_t)Fpos * (size_t)T);
    if (!full) return ORC_ERR_ARG;
    int rc = spectrogram_pos_impl(wave, cplx_in, n, nperseg, hop, nfft, full);
    if (rc) { free(full); return rc; }
    int k_lo, k_hi;
    long j_lo, j_hi;
    orc_masks(fs, nperseg, hop, nfft, T, freq_min, freq_max, time_min, time_max, &k_lo, &k_hi, &j_lo, &j_hi);
    int F = k_hi - k_lo;
    int Tm = (int)(j_hi - j_lo);
    if (F_out) *F_out = F;
    if (T_out) *T_out = Tm;
    if (F <= 0 || Tm <= 0) { free(full); return 0; }
    double *mag = (double *)malloc(sizeof(double) * (size_t)F * (size_t)Tm);
    for (int k = 0; k < F; k++)
        memcpy(mag + (size_t)k * Tm, full + (size_t)(k + k_lo) * T + j_lo, sizeof(double) * (size_t)Tm);
    free(full);
    int K = max_candidates;
    int *ct = (int *)malloc(sizeof(int) * (K > 0 ? K : 1));
    int *cf = (int *)malloc(sizeof(int) * (K > 0 ? K : 1));
    double *cs = (double *)malloc(sizeof(double) * (K > 0 ? K : 1));
    int nc = orc_find_candidates(mag, F, Tm, steps_per_symbol, bins_per_tone, K, min_score, ct, cf, cs, NULL);
    if (nc >= 0) {
        for (int i = 0; i < nc && i < max_records; i++) {
            orc_decode_candidate(mag, F, Tm, steps_per_symbol, bins_per_tone, ct[i], cf[i], max_iterations, &records[i]);
            records[i].score = cs[i];
        }
    }
    free(ct); free(cf); free(cs); free(mag);
    return nc;
}

int orc_decode_message(const double *wave, long n, double fs, int bins_per_tone, int steps_per_symbol,
                       int max_candidates, double min_score, int max_iterations, double freq_min,
                       double freq_max, double time_min, double time_max, orc_record *records, int max_records,
                       int *F_out, int *T_out) {
    return decode_message_impl(wave, 0, n, fs, bins_per_tone, steps_per_symbol, max_candidates, min_score, max_iterations,
                               freq_min, freq_max, time_min, time_max, records, max_records, F_out, T_out);
}
/* the same for complex audio (n complex samples, interleaved) */
int orc_decode_message_c(const double *wave, long n, double fs, int bins_per_tone, int steps_per_symbol,
                         int max_candidates, double min_score, int max_iterations, double freq_min,
                         double freq_max, double time_min, double time_max, orc_record *records, int max_records,
                         int *F_out, int *T_out) {
    return decode_message_impl(wave, 1, n, fs, bins_per_tone, steps_per_symbol, max_candidates, min_score, max_iterations,
                               freq_min, freq_max, time_min, time_max, records, max_records, F_out, T_out);
}

/* ---- batch entry points (CPU baseline legs of bench.py): pthreads, dynamic work queue ---- */
#include <pthread.h>

typedef struct {
    int kind;            /* 0 = decode slots, 1 = BP words */
    long n_items;
    long next;           /* guarded by mu */
    pthread_mutex_t mu;
    /* decode */
    const double *waves; long n; double fs; int b, s, K; double min_score; int max_it;
    orc_record *records; int *n_cand;
    /* bp */
    const double *llr; uint8_t *plain; int32_t *errors;
    int err;
} orc_job;

static void *orc_worker(void *arg) {
    orc_job *J = (orc_job *)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        long i0 = J->next;
        long chunk = J->kind == 0 ? 1 : 64;
        J->next += chunk;
        pthread_mutex_unlock(&J->mu);
        if (i0 >= J->n_items) break;
        long i1 = i0 + chunk < J->n_items ? i0 + chunk : J->n_items;
        for (long i = i0; i < i1; i++) {
            if (J->kind == 0) {
                int nc = orc_decode_message(J->waves + i * J->n, J->n, J->fs, J->b, J->s, J->K, J->min_score, J->max_it,
                                            NAN, NAN, NAN, NAN, J->records + i * (long)J->K, J->K, NULL, NULL);
                J->n_cand[i] = nc;
                if (nc < 0) J->err = nc;
            } else {
                J->errors[i] = orc_bp_decode(J->llr + i * ORC_LDPC_N, J->max_it, J->plain + i * ORC_LDPC_N);
            }
        }
    }
    return NULL;
}

static int orc_run(orc_job *J, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_mutex_init(&J->mu, NULL);
    J->next = 0;
    J->err = 0;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    int started = 0;
    for (int i = 0; i < n_threads - 1; i++)
        if (pthread_create(&th[started], NULL, orc_worker, J) == 0) started++;
    orc_worker(J);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&J->mu);
    return J->err;
}

/* Batch of equal-length slots; records[i*K + c] for slot i, candidate c. */
int orc_decode_batch(const double *waves, long n_slots, long n, double fs, int bins_per_tone, int steps_per_symbol,
                     int max_candidates, double min_score, int max_iterations, orc_record *records, int *n_cand,
                     int n_threads) {
    orc_job J;
    memset(&J, 0, sizeof(J));
    J.kind = 0; J.n_items = n_slots; J.waves = waves; J.n = n; J.fs = fs; J.b = bins_per_tone; J.s = steps_per_symbol;
    J.K = max_candidates; J.min_score = min_score; J.max_it = max_iterations; J.records = records; J.n_cand = n_cand;
    return orc_run(&J, n_threads);
}

/* Batch BP over standalone LLR vectors (config C5 of BASELINE.json). */
int orc_bp_batch(const double *llr, long n_words, int max_iterations, uint8_t *plain, int32_t *errors, int n_threads) {
    orc_job J;
    memset(&J, 0, sizeof(J));
    J.kind = 1; J.n_items = n_words; J.llr = llr; J.max_it = max_iterations; J.plain = plain; J.errors = errors;
    return orc_run(&J, n_threads);
}

int orc_record_size(void) { return (int)sizeof(orc_record); }
