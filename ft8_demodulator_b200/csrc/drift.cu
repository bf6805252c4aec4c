// drift.cu - data-parallel pieces of the satellite pre-processor correct_frequency_drift
// (reference ft8_beacon_receiver/frequency_correction.py:118-659):
//   column_argmax : per spectrogram column, index of the strongest frequency row      (:222-224, :382-384)
//   dechirp       : multiply by exp(-2*pi*i*(c2*n^2 + c3*n^3)), n = sample index        (:352-355, :598-634)
// The regression / template-correlation logic around them is scalar host work on a few hundred points and stays
// in the Python mirror.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace ft8b {

// mag [n_slots][T][F] (time-major, as ft8b_stft_db writes it) -> idx [n_slots][T]; first maximum wins (np.argmax).
__global__ void column_argmax_kernel(const double *mag, int64_t n_cols, int F, int f_lo, int f_hi, int32_t *idx) {
    const int lane = threadIdx.x & 31;
    const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (col >= n_cols) return;
    const double *row = mag + col * (int64_t)F;
    double best = -INFINITY;
    int bi = 0x7fffffff;
    for (int it = 0, f = f_lo + lane; it < (f_hi - f_lo + 31) / 32; ++it, f += 32) {
        if (f < f_hi) {
            const double v = __ldg(row + f);
            // NaN never wins unless everything is NaN (np.argmax returns the first NaN; not reproduced)
            if (v > best || (v == best && f < bi)) { best = v; bi = f; }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) idx[col] = (bi == 0x7fffffff) ? 0 : bi - f_lo;
}

int column_argmax(const double *mag, int64_t n_cols, int F, int f_lo, int f_hi, int32_t *idx, cudaStream_t stream) {
    if (!mag || !idx || F < 1 || f_lo < 0 || f_hi > F || f_lo >= f_hi || n_cols < 0) FT8B_ARG_FAIL("ft8b_column_argmax: bad argument");
    if (n_cols == 0) return 0;
    const int threads = 256;
    const int64_t blocks = (n_cols * 32 + threads - 1) / threads;
    column_argmax_kernel<<<(unsigned)blocks, threads, 0, stream>>>(mag, n_cols, F, f_lo, f_hi, idx);
    FT8B_LAUNCH_CHECK("column_argmax_kernel");
    return 0;
}

// out[n] = in[n] * exp(-2*pi*i*(c2*n^2 + c3*n^3)).  The phase is reduced to one turn in float64 before sincospi.
template <typename InT>
__global__ void dechirp_kernel(const InT *in, int64_t n, double c2, double c3, double2 *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)i;
        double turns = x * x * (c2 + c3 * x);
        turns -= rint(turns);
        double sn, cs;
        sincospi(-2.0 * turns, &sn, &cs);
        const InT v = in[i];
        const double re = (double)v.x, im = (double)v.y;
        out[i] = make_double2(re * cs - im * sn, re * sn + im * cs);
    }
}

int dechirp(const void *in, int in_dtype, int64_t n, double c2, double c3, double2 *out, cudaStream_t stream) {
    if (!in || !out || n < 0) FT8B_ARG_FAIL("ft8b_dechirp: bad argument");
    if (in_dtype != FT8B_DTYPE_C64 && in_dtype != FT8B_DTYPE_C128) FT8B_ARG_FAIL("ft8b_dechirp: input must be complex64/complex128");
    if (n == 0) return 0;
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (in_dtype == FT8B_DTYPE_C64)
        dechirp_kernel<float2><<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<const float2 *>(in), n, c2, c3, out);
    else
        dechirp_kernel<double2><<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<const double2 *>(in), n, c2, c3, out);
    FT8B_LAUNCH_CHECK("dechirp_kernel");
    return 0;
}


// ------------------------------------------------------------------------------------------------------------
// Drift-hypothesis Costas search (extension named in BASELINE north_star for the satellite path; the reference has
// no equivalent, so this is NEW behaviour with payload-level checks only - SURVEY.md fact 9 / 8f-1).
// For hypothesis h the tone grid of symbol q (counted from the candidate's first symbol) is shifted by
// round(d_h * q) bins.  Each (abs_time, abs_freq) gets the best score over the hypotheses (first maximum wins, so put
// d = 0 first) and the index of that hypothesis.  The score of one hypothesis is ft8_sync_score's running sum
// (ft8_decode.py:47-100) on the shifted cells; with d = 0 it is exactly the reference score.  Cells whose own or
// neighbouring bins leave the waterfall are skipped like out-of-range symbols.
// ------------------------------------------------------------------------------------------------------------
constexpr int kMaxDriftHyp = 128;
constexpr int kNumCells = kNumSync * kLenSync;     // 21 Costas cells per candidate
constexpr int kDriftStrip = 128;                   // candidate frequencies (= threads) per CTA of the tiled kernel
constexpr int kMaxDriftShift = 64;                 // largest |round(d*q)| the tiled kernel stages; beyond: direct kernel
__device__ __constant__ double c_drift[kMaxDriftHyp];
__device__ __constant__ int c_cell_off[kMaxDriftHyp * kNumCells];   // byte offset of (hypothesis, cell) in the CTA's tile
__device__ __constant__ int c_costas[7] = {3, 1, 4, 0, 6, 5, 2};

// Direct formulation: one thread per (time, frequency), the definition evaluated hypothesis by hypothesis with the
// reference's running sum.  Any shift magnitude; used for |shift| > kMaxDriftShift and as the cross-check of the
// tiled kernel (tests/test_gpu_parity.py).
__global__ void sync_scores_drift_direct_kernel(const double *mag, int T, int F, int s, int b, int nb, int t0, int nt, int nf,
                                                int n_hyp, double *best_score, int32_t *best_hyp) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int ti = blockIdx.y;
    const int64_t slot = blockIdx.z;
    if (f >= nf) return;
    const int at = t0 + ti;
    int base = at / s;
    if ((at % s) != 0 && at < 0) --base;
    const double *m0 = mag + slot * (int64_t)T * F;
    double best = -INFINITY;
    int bh = 0;
    for (int h = 0; h < n_hyp; ++h) {
        const double d = c_drift[h];
        double score = 0.0;
        int num = 0;
        for (int m = 0; m < kNumSync; ++m) {
            for (int k = 0; k < kLenSync; ++k) {
                const int q = m * kSyncOffset + k;
                const int blk = base + q;
                if (blk < 0 || blk >= nb) continue;
                const int tone = c_costas[k];
                const int bin = f + (int)lrint(d * (double)q) + tone * b;
                if (bin - (tone > 0 ? b : 0) < 0 || bin + b >= F) continue;
                const double *col = m0 + (int64_t)(at + q * s) * F;
                const double p = __ldg(col + bin);
                if (tone > 0) { score = __dadd_rn(score, __dsub_rn(p, __ldg(col + bin - b))); ++num; }
                score = __dadd_rn(score, __dsub_rn(p, __ldg(col + bin + b))); ++num;
                if (k > 0 && blk > 0) { score = __dadd_rn(score, __dsub_rn(p, __ldg(col - (int64_t)s * F + bin))); ++num; }
                if (k < kLenSync - 1 && blk + 1 < nb) { score = __dadd_rn(score, __dsub_rn(p, __ldg(col + (int64_t)s * F + bin))); ++num; }
            }
        }
        double v = (num == 0 || !isfinite(score)) ? -INFINITY : __ddiv_rn(score, (double)num);
        if (v > best) { best = v; bh = h; }
    }
    const int64_t o = (slot * nt + ti) * (int64_t)nf + f;
    best_score[o] = best;
    best_hyp[o] = bh;
}

// Tiled formulation.  The score of (abs_time, abs_freq, hypothesis) is a sum of 21 cell contrasts
//   C_q(x) = [p - low] + (p - high) + [p - earlier] + [p - later],   p = mag[abs_time + q*s][x],
// evaluated at x = abs_freq + tone_q*b + shift(h, q).  C_q(x) does not depend on the hypothesis, only WHERE it is
// read does, so a CTA (one slot, one abs_time, 128 consecutive abs_freq) first forms the 21 rows of contrasts over
// the strip widened by the largest shift (5 loads + <= 7 adds per cell, once), and then every thread walks the
// hypotheses with 21 shared-memory loads and 21 adds each instead of 105 global loads and 168 flops.  The byte offset
// of (hypothesis, cell) inside the tile comes from constant memory (warp-uniform).  Cells whose symbol or bins fall
// outside the waterfall contribute 0 and are not counted (same rule as the direct kernel); only the first and last
// strip can lose cells in frequency, so the per-cell counts exist in the EDGE instantiation only.
// Summation: contrasts per cell, cells per Costas block, blocks - within 1e-12 of the running sum (test).
template <bool EDGE>
__device__ __forceinline__ void drift_walk(const char *tile_thread, const unsigned char *cnt_thread, int n_hyp, int n_time,
                                           double &best, int &bh) {
    best = -INFINITY;
    bh = 0;
    for (int h = 0; h < n_hyp; ++h) {
        const int *off = c_cell_off + h * kNumCells;
        double blk_sum[kNumSync];
        int n = 0;
#pragma unroll
        for (int m = 0; m < kNumSync; ++m) {
            double a = *reinterpret_cast<const double *>(tile_thread + off[m * kLenSync]);
            if (EDGE) n += cnt_thread[off[m * kLenSync] >> 3];
#pragma unroll
            for (int k = 1; k < kLenSync; ++k) {
                a = __dadd_rn(a, *reinterpret_cast<const double *>(tile_thread + off[m * kLenSync + k]));
                if (EDGE) n += cnt_thread[off[m * kLenSync + k] >> 3];
            }
            blk_sum[m] = a;
        }
        double v = __dadd_rn(__dadd_rn(blk_sum[0], blk_sum[1]), blk_sum[2]);
        if (EDGE) v = (n == 0 || !isfinite(v)) ? -INFINITY : __ddiv_rn(v, (double)n);
        else v = isfinite(v) ? v : -INFINITY;       // all hypotheses share n_time: one division after the loop
        if (v > best) { best = v; bh = h; }
    }
    if (!EDGE) best = (n_time == 0 || best == -INFINITY) ? -INFINITY : __ddiv_rn(best, (double)n_time);
}

#ifndef FT8B_DRIFT_MINB
#define FT8B_DRIFT_MINB 6
#endif
#ifndef FT8B_DRIFT_U
#define FT8B_DRIFT_U 5
#endif

// The symbols of one Costas cell (block m, position k) for a candidate starting at abs_time `at`.
struct DriftCell {
    int q, tone, n_cell;
    bool live, has_low, has_prev, has_next;
};
__device__ __forceinline__ DriftCell drift_cell(int qi, int base, int nb) {
    DriftCell c;
    const int m = qi / kLenSync, k = qi - m * kLenSync;
    c.q = m * kSyncOffset + k;
    const int blk = base + c.q;
    c.tone = c_costas[k];
    c.live = blk >= 0 && blk < nb;
    c.has_low = c.tone > 0;
    c.has_prev = k > 0 && blk > 0;
    c.has_next = k < kLenSync - 1 && blk + 1 < nb;
    c.n_cell = c.live ? (int)c.has_low + 1 + (int)c.has_prev + (int)c.has_next : 0;
    return c;
}

// One warp forms one row of cell contrasts, row[j] = C(x0 + j), j < wc; a lane takes kU cells per pass with all their
// loads issued up front (the tile is built from L2/L1 hits whose latency would otherwise be paid cell after cell).
// cnt (edge strips only): comparisons behind each contrast, 0 where the cell's bins leave the waterfall.
__device__ __forceinline__ void drift_contrast_row(const DriftCell &c, const double *col, int64_t step, int x0, int b, int F,
                                                   int wc, int lane, double *row, unsigned char *cnt) {
    constexpr int kU = FT8B_DRIFT_U;
    const int lo_b = c.has_low ? b : 0;
    for (int j0 = lane; j0 < wc; j0 += 32 * kU) {
        double p[kU], hi[kU], lo[kU], pv[kU], nx[kU];
        bool ok[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int x = x0 + j0 + 32 * u;
            ok[u] = c.live && j0 + 32 * u < wc && x - lo_b >= 0 && x + b < F;
            p[u] = hi[u] = lo[u] = pv[u] = nx[u] = 0.0;
            if (ok[u]) {
                p[u] = __ldg(col + x);
                hi[u] = __ldg(col + x + b);
                if (c.has_low) lo[u] = __ldg(col + x - b);
                if (c.has_prev) pv[u] = __ldg(col - step + x);
                if (c.has_next) nx[u] = __ldg(col + step + x);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int j = j0 + 32 * u;
            if (j >= wc) break;
            double v = 0.0;
            if (ok[u]) {
                v = __dsub_rn(p[u], hi[u]);
                if (c.has_low) v = __dadd_rn(__dsub_rn(p[u], lo[u]), v);
                if (c.has_prev) v = __dadd_rn(v, __dsub_rn(p[u], pv[u]));
                if (c.has_next) v = __dadd_rn(v, __dsub_rn(p[u], nx[u]));
            }
            row[j] = v;
            if (cnt) cnt[j] = ok[u] ? (unsigned char)c.n_cell : (unsigned char)0;
        }
    }
}

__global__ void __launch_bounds__(kDriftStrip, FT8B_DRIFT_MINB)
sync_scores_drift_tiled_kernel(const double *mag, int T, int F, int s, int b, int nb, int t0, int nt, int nf, int n_hyp,
                               int max_shift, double *best_score, int32_t *best_hyp) {
    extern __shared__ double tile[];                       // [21][wc] contrasts, then (EDGE) [21][wc] counts (bytes)
    const int wc = kDriftStrip + 2 * max_shift;
    unsigned char *cnt = reinterpret_cast<unsigned char *>(tile + kNumCells * wc);
    const int f0 = blockIdx.x * kDriftStrip;
    const int ti = blockIdx.y;
    const int64_t slot = blockIdx.z;
    const int at = t0 + ti;
    int base = at / s;
    if ((at % s) != 0 && at < 0) --base;
    const double *m0 = mag + slot * (int64_t)T * F;
    // every cell of this strip has all its bins inside the waterfall?
    const bool edge = (f0 - max_shift < 0) || (f0 + kDriftStrip - 1 + max_shift + 7 * b >= F);
    int n_time = 0;                                        // comparisons per score when no cell is lost in frequency
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int qi = 0; qi < kNumCells; ++qi) {               // rows are dealt to the four warps
        const DriftCell c = drift_cell(qi, base, nb);
        n_time += c.n_cell;
        if ((qi & 3) != warp) continue;
        drift_contrast_row(c, m0 + (int64_t)(at + c.q * s) * F, (int64_t)s * F, f0 - max_shift + c.tone * b, b, F, wc, lane,
                           tile + qi * wc, edge ? cnt + qi * wc : nullptr);
    }
    __syncthreads();
    const int f = f0 + threadIdx.x;
    if (f >= nf) return;
    double best;
    int bh;
    const char *tile_thread = reinterpret_cast<const char *>(tile + threadIdx.x);
    if (edge) drift_walk<true>(tile_thread, cnt + threadIdx.x, n_hyp, n_time, best, bh);
    else drift_walk<false>(tile_thread, nullptr, n_hyp, n_time, best, bh);
    const int64_t o = (slot * nt + ti) * (int64_t)nf + f;
    best_score[o] = best;
    best_hyp[o] = bh;
}

// Block-sum formulation (the one the library launches when it fits).  Within one Costas block the shift of a
// hypothesis is base + delta_k with base = round(d * 36m) and a 7-entry pattern delta_k = shift(36m + k) - base that only
// a handful of hypotheses-independent shapes can take (for |6d| <= 1.5: no step, or one step of +-1 at one of six
// positions).  So the sum of a block's seven contrasts is formed ONCE per (pattern, frequency),
//   B_m[pattern][x] = sum_k C_{m,k}(x + delta_k),
// and a hypothesis reads three block sums, B_0[..][f + base_0] + B_1[..][f + base_1] + B_2[..][f + base_2], instead of
// 21 contrasts: with 29 hypotheses the shared-memory loads per score drop from 21 to 3 + (patterns * 7 * (128 + 2
// max|base|) / 128) / 29 ~ 12, with 128 hypotheses to ~5.  The contrast tile only ever holds one block (7 rows).
// Same summation order as drift_walk: contrasts of a block in k order, then (B_0 + B_1) + B_2 - bit-identical scores.
constexpr int kMaxDriftPatterns = 96;
__device__ __constant__ int c_pat_off[kMaxDriftPatterns * kLenSync];     // element offset in the block's contrast tile: k*wc + G + delta_k
__device__ __constant__ int c_pat_begin[kNumSync + 1];                   // pattern rows of block m: [begin[m], begin[m+1])
__device__ __constant__ int c_bsum_off[kMaxDriftHyp * kNumSync];         // byte offset of (hypothesis, block) in the block-sum tile

__global__ void __launch_bounds__(kDriftStrip, 4)
sync_scores_drift_blocksum_kernel(const double *mag, int T, int F, int s, int b, int nb, int t0, int nt, int nf, int n_hyp,
                                  int max_base, int guard, double *best_score, int32_t *best_hyp) {
    extern __shared__ double tile[];
    const int wb = kDriftStrip + 2 * max_base;             // block-sum columns: frequency offset f - f0 + max_base + base
    const int margin = max_base + guard;
    const int wc = kDriftStrip + 2 * margin;               // contrast columns: bin offset + margin
    const int n_pat = c_pat_begin[kNumSync];
    double *bsum = tile;                                   // [n_pat][wb]
    double *con = bsum + n_pat * wb;                       // [7][wc], one Costas block at a time
    unsigned char *bcnt = reinterpret_cast<unsigned char *>(con + kLenSync * wc);      // [n_pat][wb]   (edge strips)
    unsigned char *ccnt = bcnt + n_pat * wb;                                           // [7][wc]
    const int f0 = blockIdx.x * kDriftStrip;
    const int ti = blockIdx.y;
    const int64_t slot = blockIdx.z;
    const int at = t0 + ti;
    int base = at / s;
    if ((at % s) != 0 && at < 0) --base;
    const double *m0 = mag + slot * (int64_t)T * F;
    const bool edge = (f0 - margin < 0) || (f0 + kDriftStrip - 1 + margin + 7 * b >= F);
    int n_time = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int m = 0; m < kNumSync; ++m) {
        for (int k = 0; k < kLenSync; ++k) {
            const DriftCell c = drift_cell(m * kLenSync + k, base, nb);
            n_time += c.n_cell;
            if ((k & 3) != warp) continue;
            drift_contrast_row(c, m0 + (int64_t)(at + c.q * s) * F, (int64_t)s * F, f0 - margin + c.tone * b, b, F, wc, lane,
                               con + k * wc, edge ? ccnt + k * wc : nullptr);
        }
        __syncthreads();
        for (int r = c_pat_begin[m] + warp; r < c_pat_begin[m + 1]; r += 4) {      // pattern rows dealt to the warps
            const int *off = c_pat_off + r * kLenSync;
            for (int j = lane; j < wb; j += 32) {
                double a = con[off[0] + j];
#pragma unroll
                for (int k = 1; k < kLenSync; ++k) a = __dadd_rn(a, con[off[k] + j]);
                bsum[r * wb + j] = a;
                if (edge) {
                    int n = 0;
#pragma unroll
                    for (int k = 0; k < kLenSync; ++k) n += ccnt[off[k] + j];
                    bcnt[r * wb + j] = (unsigned char)n;
                }
            }
        }
        __syncthreads();
    }
    const int f = f0 + threadIdx.x;
    if (f >= nf) return;
    const char *mine = reinterpret_cast<const char *>(bsum + threadIdx.x);
    const unsigned char *mine_cnt = bcnt + threadIdx.x;
    double best = -INFINITY;
    int bh = 0;
    for (int h = 0; h < n_hyp; ++h) {
        const int *off = c_bsum_off + h * kNumSync;
        double v = __dadd_rn(__dadd_rn(*reinterpret_cast<const double *>(mine + off[0]), *reinterpret_cast<const double *>(mine + off[1])),
                             *reinterpret_cast<const double *>(mine + off[2]));
        if (edge) {
            const int n = mine_cnt[off[0] >> 3] + mine_cnt[off[1] >> 3] + mine_cnt[off[2] >> 3];
            v = (n == 0 || !isfinite(v)) ? -INFINITY : __ddiv_rn(v, (double)n);
        } else {
            v = isfinite(v) ? v : -INFINITY;
        }
        if (v > best) { best = v; bh = h; }
    }
    if (!edge) best = (n_time == 0 || best == -INFINITY) ? -INFINITY : __ddiv_rn(best, (double)n_time);
    const int64_t o = (slot * nt + ti) * (int64_t)nf + f;
    best_score[o] = best;
    best_hyp[o] = bh;
}

static int drift_prepare(const double *mag, const double *d, const double *best_score, const int32_t *best_hyp, int n_hyp,
                         int T, int F, int s, int b) {
    if (!mag || !d || !best_score || !best_hyp) FT8B_ARG_FAIL("ft8b_sync_scores_drift: null pointer");
    if (n_hyp < 1 || n_hyp > kMaxDriftHyp) FT8B_ARG_FAIL("ft8b_sync_scores_drift: 1 <= n_hyp <= %d", kMaxDriftHyp);
    if (T < 0 || F < 0 || s < 1 || b < 1) FT8B_ARG_FAIL("ft8b_sync_scores_drift: bad shape");
    for (int h = 0; h < n_hyp; ++h)
        if (!std::isfinite(d[h]) || std::fabs(d[h]) > 1e6) FT8B_ARG_FAIL("ft8b_sync_scores_drift: hypothesis %d is not a finite drift", h);
    return 0;
}

int sync_scores_drift(const double *mag, int64_t n_slots, int T, int F, int s, int b, int n_hyp,
                      const double *drift_bins_per_symbol, double *best_score, int32_t *best_hyp, int mode,
                      cudaStream_t stream) {
    // mode 0: the library's choice (block sums when they fit and pay, else the contrast walk, else direct);
    // 1: contrast walk (or direct beyond its shift limit); 2: block sums or fail; 3: direct
    if (int rc = drift_prepare(mag, drift_bins_per_symbol, best_score, best_hyp, n_hyp, T, F, s, b)) return rc;
    int t0, nt, nf;
    ft8b_search_ranges(T, F, s, b, &t0, &nt, &nf);
    if (n_slots <= 0 || nt <= 0 || nf <= 0) return 0;
    if (nt > 65535 || n_slots > 65535) FT8B_ARG_FAIL("ft8b_sync_scores_drift: too many time offsets / slots per call");
    // shift(h, q) = round-half-even(d_h * q), the same expression the direct kernel and the LLR extraction evaluate
    int shift[kMaxDriftHyp * kNumCells];
    long max_shift = 0;
    for (int h = 0; h < n_hyp; ++h)
        for (int qi = 0; qi < kNumCells; ++qi) {
            const int q = (qi / kLenSync) * kSyncOffset + qi % kLenSync;
            const long sh = std::lrint(drift_bins_per_symbol[h] * (double)q);
            shift[h * kNumCells + qi] = (int)sh;
            if (std::labs(sh) > max_shift) max_shift = std::labs(sh);
        }
    if (mode == 3 || max_shift > kMaxDriftShift) {
        FT8B_CUDA_TRY(cudaMemcpyToSymbolAsync(c_drift, drift_bins_per_symbol, sizeof(double) * n_hyp, 0, cudaMemcpyHostToDevice, stream));
        dim3 grid((nf + 127) / 128, nt, (unsigned)n_slots);
        sync_scores_drift_direct_kernel<<<grid, 128, 0, stream>>>(mag, T, F, s, b, T / s, t0, nt, nf, n_hyp, best_score, best_hyp);
        FT8B_LAUNCH_CHECK("sync_scores_drift_direct_kernel");
        return 0;
    }
    if (mode != 1) {
        // block-sum plan: per block m the base shift and the pattern of the other six shifts relative to it
        int pat[kMaxDriftPatterns][kLenSync], pat_block[kMaxDriftPatterns], n_pat = 0, begin[kNumSync + 1];
        int row_of[kMaxDriftHyp * kNumSync], base_of[kMaxDriftHyp * kNumSync];
        long max_base = 0, guard = 0;
        bool fits = true;
        for (int m = 0; m < kNumSync && fits; ++m) {
            begin[m] = n_pat;
            for (int h = 0; h < n_hyp && fits; ++h) {
                const int *sh = shift + h * kNumCells + m * kLenSync;
                int d[kLenSync];
                for (int k = 0; k < kLenSync; ++k) {
                    d[k] = sh[k] - sh[0];
                    guard = std::max<long>(guard, std::labs((long)d[k]));
                }
                max_base = std::max<long>(max_base, std::labs((long)sh[0]));
                int r = begin[m];
                while (r < n_pat && std::memcmp(pat[r], d, sizeof(d)) != 0) ++r;
                if (r == n_pat) {
                    if (n_pat == kMaxDriftPatterns) { fits = false; break; }
                    std::memcpy(pat[n_pat], d, sizeof(d));
                    pat_block[n_pat++] = m;
                }
                row_of[h * kNumSync + m] = r;
                base_of[h * kNumSync + m] = sh[0];
            }
        }
        begin[kNumSync] = n_pat;
        const int wb = kDriftStrip + 2 * (int)max_base, wcb = kDriftStrip + 2 * (int)(max_base + guard);
        const size_t smem_b = ((size_t)n_pat * wb + (size_t)kLenSync * wcb) * (sizeof(double) + 1);
        // worth it when three CTAs fit an SM and it reads well under the shared memory of the 21-contrast walk: it runs
        // at 4 instead of 6 CTAs per SM and builds the pattern rows first, so it only pays for longer hypothesis lists
        // than the load count alone suggests (tools/drift_variant_sweep.py, +-3.33 Hz/s: 17 / 29 / 57 / 128 hypotheses = load ratio 0.73 / 0.68 /
        // 0.42 / 0.26 -> walk 23.4 / 31.6 / 53.9 / 111.9 us per slot, block sums 28.8 / 33.2 / 41.7 / 69.4)
        constexpr size_t kBlocksumSmemMax = 75 * 1024;
        const double loads_blocksum = 3.0 * n_hyp + 7.0 * n_pat * wb / kDriftStrip, loads_walk = 21.0 * n_hyp;
        if (fits && max_base + guard <= kMaxDriftShift && smem_b <= kBlocksumSmemMax && (mode == 2 || (n_hyp >= 40 && loads_blocksum < 0.8 * loads_walk))) {
            int pat_off[kMaxDriftPatterns * kLenSync], bsum_off[kMaxDriftHyp * kNumSync];
            for (int r = 0; r < n_pat; ++r)
                for (int k = 0; k < kLenSync; ++k) pat_off[r * kLenSync + k] = k * wcb + (int)guard + pat[r][k];
            for (int i = 0; i < n_hyp * kNumSync; ++i) bsum_off[i] = (row_of[i] * wb + (int)max_base + base_of[i]) * (int)sizeof(double);
            (void)pat_block;
            FT8B_CUDA_TRY(cudaMemcpyToSymbolAsync(c_pat_off, pat_off, sizeof(int) * n_pat * kLenSync, 0, cudaMemcpyHostToDevice, stream));
            FT8B_CUDA_TRY(cudaMemcpyToSymbolAsync(c_pat_begin, begin, sizeof(begin), 0, cudaMemcpyHostToDevice, stream));
            FT8B_CUDA_TRY(cudaMemcpyToSymbolAsync(c_bsum_off, bsum_off, sizeof(int) * n_hyp * kNumSync, 0, cudaMemcpyHostToDevice, stream));
            if (smem_b > 48 * 1024) {
                static std::once_flag once[64];
                int dev = 0;
                FT8B_CUDA_TRY(cudaGetDevice(&dev));
                cudaError_t e = cudaSuccess;
                std::call_once(once[dev & 63], [&] {
                    e = cudaFuncSetAttribute(sync_scores_drift_blocksum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlocksumSmemMax);
                });
                if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(sync_scores_drift_blocksum_kernel)");
            }
            dim3 grid((nf + kDriftStrip - 1) / kDriftStrip, nt, (unsigned)n_slots);
            sync_scores_drift_blocksum_kernel<<<grid, kDriftStrip, smem_b, stream>>>(mag, T, F, s, b, T / s, t0, nt, nf, n_hyp,
                                                                                     (int)max_base, (int)guard, best_score, best_hyp);
            FT8B_LAUNCH_CHECK("sync_scores_drift_blocksum_kernel");
            return 0;
        }
        if (mode == 2) FT8B_ARG_FAIL("ft8b_sync_scores_drift: the block-sum kernel does not cover this hypothesis list");
    }
    const int wc = kDriftStrip + 2 * (int)max_shift;
    for (int i = 0; i < n_hyp * kNumCells; ++i) shift[i] = ((i % kNumCells) * wc + (int)max_shift + shift[i]) * (int)sizeof(double);
    FT8B_CUDA_TRY(cudaMemcpyToSymbolAsync(c_cell_off, shift, sizeof(int) * n_hyp * kNumCells, 0, cudaMemcpyHostToDevice, stream));
    const size_t smem = (size_t)kNumCells * wc * (sizeof(double) + 1);      // <= 48 384 B at kMaxDriftShift
    dim3 grid((nf + kDriftStrip - 1) / kDriftStrip, nt, (unsigned)n_slots);
    sync_scores_drift_tiled_kernel<<<grid, kDriftStrip, smem, stream>>>(mag, T, F, s, b, T / s, t0, nt, nf, n_hyp, (int)max_shift,
                                                                        best_score, best_hyp);
    FT8B_LAUNCH_CHECK("sync_scores_drift_tiled_kernel");
    return 0;
}

}  // namespace ft8b
