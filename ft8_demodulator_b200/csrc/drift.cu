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

}  // namespace ft8b
