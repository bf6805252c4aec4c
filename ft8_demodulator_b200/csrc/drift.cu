// drift.cu - data-parallel pieces of the satellite pre-processor correct_frequency_drift
// (reference ft8_beacon_receiver/frequency_correction.py:118-659):
//   column_argmax : per spectrogram column, index of the strongest frequency row      (:222-224, :382-384)
//   dechirp       : multiply by exp(-2*pi*i*(c2*n^2 + c3*n^3)), n = sample index        (:352-355, :598-634)
// The regression / template-correlation logic around them is scalar host work on a few hundred points and stays
// in the Python mirror.
#include <cmath>

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
__device__ __constant__ double c_drift[kMaxDriftHyp];
__device__ __constant__ int c_costas[7] = {3, 1, 4, 0, 6, 5, 2};

__global__ void sync_scores_drift_kernel(const double *mag, int T, int F, int s, int b, int nb, int t0, int nt, int nf,
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

int sync_scores_drift(const double *mag, int64_t n_slots, int T, int F, int s, int b, int n_hyp,
                      const double *drift_bins_per_symbol, double *best_score, int32_t *best_hyp, cudaStream_t stream) {
    if (!mag || !drift_bins_per_symbol || !best_score || !best_hyp) FT8B_ARG_FAIL("ft8b_sync_scores_drift: null pointer");
    if (n_hyp < 1 || n_hyp > kMaxDriftHyp) FT8B_ARG_FAIL("ft8b_sync_scores_drift: 1 <= n_hyp <= %d", kMaxDriftHyp);
    if (T < 0 || F < 0 || s < 1 || b < 1) FT8B_ARG_FAIL("ft8b_sync_scores_drift: bad shape");
    int t0, nt, nf;
    ft8b_search_ranges(T, F, s, b, &t0, &nt, &nf);
    if (n_slots <= 0 || nt <= 0 || nf <= 0) return 0;
    if (nt > 65535 || n_slots > 65535) FT8B_ARG_FAIL("ft8b_sync_scores_drift: too many time offsets / slots per call");
    FT8B_CUDA_TRY(cudaMemcpyToSymbolAsync(c_drift, drift_bins_per_symbol, sizeof(double) * n_hyp, 0, cudaMemcpyHostToDevice, stream));
    dim3 grid((nf + 127) / 128, nt, (unsigned)n_slots);
    sync_scores_drift_kernel<<<grid, 128, 0, stream>>>(mag, T, F, s, b, T / s, t0, nt, nf, n_hyp, best_score, best_hyp);
    FT8B_LAUNCH_CHECK("sync_scores_drift_kernel");
    return 0;
}

}  // namespace ft8b
