// sync.cu - S2: Costas sync score map and candidate selection (float64).
//
// S2a replaces ft8_sync_score over the double loop of ft8_find_candidates (reference
// ft8_decode.py:47-100, 108-124).  S2b replaces the heap bookkeeping of ft8_decode.py:127-140 with
// its closed form: H = first K threshold passers in scan order; the best of H
// is replaced by the first occurrence of the maximum over the later passers when that is strictly
// greater; H sorted by descending score.
//
// S2a design: with t = t0 + rho + s*u (rho = time-oversampling phase) every waterfall column a
// candidate touches has the same phase, col = rho + s*g, so each phase is a 1-D problem in symbol
// units g.  One thread owns one abs_freq f and walks g = 0..nb-1 once, holding the 8 tone bins
// f+tau*b of columns g and g+1 in registers.  The 7 Costas cells of a block are spread over 7
// consecutive g, so 7 rotating accumulators carry the partial block sums S7(h) (h = g-k); a finished
// S7(h) is added into score[u] for the three blocks u = h+10, h-26, h-62 (m = 0,1,2 in the
// reference's order).  Frequency differences are shared between neighbouring tones and time
// differences between consecutive steps, ~39 float64 adds per (column, f) instead of 150 per score.
// Loads are coalesced (adjacent f = adjacent bins) and every value is fetched once per tone offset
// from L1/L2; the spectrogram is read from HBM once.
#include <cmath>
#include <cuda_pipeline.h>

#include "common.cuh"

namespace ft8b {

constexpr int kSyncThreads = 128;

__device__ __constant__ int kCostas[7] = {3, 1, 4, 0, 6, 5, 2};  // ft8_decode.py:42

struct SyncArgs {
    const double *mag;   // [n_slots][T][F]
    double *scores;      // [n_slots][nt][nf]
    int T, F, s, b;      // time_osr = s, freq_osr = b
    int nb;              // num_blocks = T / s                       (ftx_types.py:31-34)
    int nu;              // candidates per phase = nb - 49 ; nt = nu*s
    int nf;              // F - 7*b                                   (ft8_decode.py:109)
    SyncTile *tiles;     // [n_slots][s][tiles per phase] strip arg-maxima, or null (pair kernels only)
};

// number of comparisons ft8_sync_score makes for time index u (any phase): ft8_decode.py:62-94
__device__ __forceinline__ int sync_num_comparisons(int u, int nb) {
    int n = 0;
    for (int m = 0; m < kNumSync; ++m)
        for (int k = 0; k < kLenSync; ++k) {
            const int g = u - 10 + m * kSyncOffset + k;
            if (g < 0 || g >= nb) continue;
            n += (kCostas[k] > 0) + 1 /* tone < 7 always */ + ((k > 0 && g > 0) ? 1 : 0) +
                 ((k < kLenSync - 1 && g + 1 < nb) ? 1 : 0);
        }
    return n;
}

// One symbol step of the walk.  cur/nxt hold the 8 tone bins of columns g and g+1; HAS_NEXT is false only
// for the last column (no later neighbour).
template <bool HAS_NEXT>
__device__ __forceinline__ double sync_step(const double (&cur)[8], const double (&nxt)[8], double (&dtprev)[7],
                                            double (&acc)[6]) {
    // dhi[tau] = A[tau] - A[tau+1]:  p - hi = dhi[tau],  p - lo = -dhi[tau-1]
    double dhi[7], dt[7];
#pragma unroll
    for (int tau = 0; tau < 7; ++tau) dhi[tau] = __dsub_rn(cur[tau], cur[tau + 1]);
    // dt[tau] = A[g][tau] - A[g+1][tau] = (p - later); (p - earlier) at the next step is -dt[tau] (subtracted there)
    if (HAS_NEXT) {
#pragma unroll
        for (int tau = 0; tau < 7; ++tau) dt[tau] = __dsub_rn(cur[tau], nxt[tau]);
    }
    // cell sums in the reference's order: (p-lo) + (p-hi) + (p-earlier) + (p-later)   (ft8_decode.py:79-94)
    double G[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int tau = (k == 0) ? 3 : (k == 1) ? 1 : (k == 2) ? 4 : (k == 3) ? 0 : (k == 4) ? 6 : (k == 5) ? 5 : 2;
        double v = (tau > 0) ? __dadd_rn(-dhi[tau - 1], dhi[tau]) : dhi[0];
        if (k > 0) v = __dsub_rn(v, dtprev[tau]);              // (p - earlier) = -(previous step's p - later); 0 at first
        if (k < 6 && HAS_NEXT) v = __dadd_rn(v, dt[tau]);
        G[k] = v;
    }
    if (HAS_NEXT) {
#pragma unroll
        for (int tau = 0; tau < 7; ++tau) dtprev[tau] = dt[tau];
    }
    // rotate the block accumulators: acc[k] = S7(h = g-k) through cell k; S7(g-6) completes
    const double done = __dadd_rn(acc[5], G[6]);
#pragma unroll
    for (int k = 5; k >= 1; --k) acc[k] = __dadd_rn(acc[k - 1], G[k]);
    acc[0] = G[0];
    return done;
}

// add a finished block sum S7(h) into the scores it belongs to: u = h+10 (m=0), h-26 (m=1), h-62 (m=2)
__device__ __forceinline__ void sync_scatter(double *scp, int stride, int h, int nu, double v) {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const unsigned u = (unsigned)(h + 10 - m * kSyncOffset);
        if (u < (unsigned)nu) scp[u * stride] = __dadd_rn(scp[u * stride], v);
    }
}

// score / num_comparisons, -inf when there was no comparison or the sum is not finite (ft8_decode.py:97-100).
// nr = (n, RN(1/n)), (0, 0) when n = 0.  v / n in three FP64 instructions: q = v*r and one FMA residual step
// (Markstein); equal to v / n on all of 8.4e8 random dividends x n = 1..42 tried on the host.  Branch-free, so the
// unrolled store loops keep several of these chains in flight.
__device__ __forceinline__ double sync_normalise(double v, double2 nr) {
    const double q = __dmul_rn(v, nr.y);
    const double r = fma(fma(-nr.x, q, v), nr.y, q);
    const bool bad = (__double2hiint(nr.y) == 0) | (((unsigned)__double2hiint(v) & 0x7ff00000u) == 0x7ff00000u);
    return bad ? -INFINITY : r;
}

constexpr int kSyncStages = 5;   // spectrogram columns staged per warp: 2 in use + 3 in flight

template <int B>
__global__ void __launch_bounds__(kSyncThreads) sync_scores_kernel(const SyncArgs a) {
    extern __shared__ double sc[];   // [nu][kSyncThreads] score sums | ring[4 warps][kSyncStages][ring_w] | double2 ncmp[nu]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int f = blockIdx.x * kSyncThreads + tid;
    const int rho = blockIdx.y;
    const int64_t slot = blockIdx.z;
    const int nb = a.nb, nu = a.nu, s = a.s;
    const int bo = (B > 0) ? B : a.b;
    const int ring_w = 32 + 7 * bo;                     // bins a warp touches per column
    double *ring = sc + (size_t)nu * kSyncThreads + (size_t)wid * kSyncStages * ring_w;
    // per-u (num_comparisons, 1/num_comparisons) as doubles; (0, 0) when the candidate has no comparison at all
    double2 *ncmp = reinterpret_cast<double2 *>(sc + (size_t)nu * kSyncThreads + (size_t)(kSyncThreads / 32) * kSyncStages * ring_w);
    for (int u = tid; u < nu; u += kSyncThreads) {
        const int n = sync_num_comparisons(u, nb);
        ncmp[u] = make_double2((double)n, n > 0 ? __drcp_rn((double)n) : 0.0);
    }
    double *scp = sc + tid;
    for (int u = 0; u < nu; ++u) scp[u * kSyncThreads] = 0.0;
    __syncthreads();
    const int f_warp = blockIdx.x * kSyncThreads + wid * 32;
    if (f_warp >= a.nf) return;                          // whole warp beyond the last candidate frequency
    const bool active = f < a.nf;

    // column g of this phase is waterfall column rho + s*g (ftx_types.py:45-47 with time_osr = s)
    const double *col0 = a.mag + slot * (int64_t)a.T * a.F + (int64_t)rho * a.F + f_warp;
    const int64_t rowstep = (int64_t)s * a.F;
    const int f_room = a.F - f_warp;                     // bins left in the row from f_warp on
    const int last = nb - 1;
    double ca[8], cb[8], dtprev[7], acc[6];
#pragma unroll
    for (int k = 0; k < 7; ++k) dtprev[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = 0.0;
    // All addressing below is incremental (no per-step multiplies / modulo): staging source and ring slots advance
    // by constant strides and wrap with one compare.
    double *const ring_lane = ring + lane;
    double *const ring_end = ring_lane + kSyncStages * ring_w;
    const double *src = col0 + lane;                     // this lane's first bin of the next column to stage
    const bool p0 = lane < f_room;
    const bool p1 = (lane + 32 < ring_w) && (lane + 32 < f_room);
    // copy column `c` (if it exists) into ring slot `dstp`; exactly one commit group per call on every lane
    auto stage_to = [&](double *dstp, int c) {
        if (c <= last) {
            if (B > 0) {                                 // ring_w <= 60: at most two 8-byte copies per lane
                if (p0) __pipeline_memcpy_async(dstp, src, 8);
                if (p1) __pipeline_memcpy_async(dstp + 32, src + 32, 8);
            } else {
                for (int it = 0, i = lane; it < (ring_w + 31) / 32; ++it, i += 32)
                    if (i < ring_w && i < f_room) __pipeline_memcpy_async(dstp + (i - lane), src + (i - lane), 8);
            }
        }
        __pipeline_commit();
        src += rowstep;
    };
    // prologue: columns 0 .. kSyncStages-2 in flight, column 0 landed
#pragma unroll
    for (int c = 0; c < kSyncStages - 1; ++c) stage_to(ring_lane + c * ring_w, c);
    __pipeline_wait_prior(kSyncStages - 2);
    __syncwarp();
    if (active) {
#pragma unroll
        for (int tau = 0; tau < 8; ++tau) ca[tau] = ring_lane[tau * bo];
    }
    // One step (column gg): stage column gg+4 into slot SLOT_S, wait until column gg+1 (slot SLOT_N) has landed
    // (3 younger groups may still be in flight), read it, do the arithmetic, add the finished block sum
    // S7(h = gg-6) into score[u] for u = h+10, h-26, h-62 (m = 0, 1, 2).  SP points at &score[u = gg+4].
#define FT8B_SYNC_STEP(CUR, NXT, GG, SLOT_S_PTR, SLOT_N_PTR, SP, HASNEXT)                                     \
    do {                                                                                                      \
        __syncwarp();                                                                                         \
        stage_to((SLOT_S_PTR), (GG) + kSyncStages - 1);                                                       \
        __pipeline_wait_prior(kSyncStages - 2);                                                               \
        __syncwarp();                                                                                         \
        if (active) {                                                                                         \
            if (HASNEXT) {                                                                                    \
                const double *rp_ = (SLOT_N_PTR);                                                             \
                _Pragma("unroll") for (int tau = 0; tau < 8; ++tau) NXT[tau] = rp_[tau * bo];                 \
            }                                                                                                 \
            const double done_ = sync_step<HASNEXT>(CUR, NXT, dtprev, acc);                                   \
            if ((unsigned)((GG) + 4) < (unsigned)nu) (SP)[0] = __dadd_rn((SP)[0], done_);                     \
            if ((unsigned)((GG)-32) < (unsigned)nu)                                                           \
                (SP)[-kSyncOffset * kSyncThreads] = __dadd_rn((SP)[-kSyncOffset * kSyncThreads], done_);      \
            if ((unsigned)((GG)-68) < (unsigned)nu)                                                           \
                (SP)[-2 * kSyncOffset * kSyncThreads] = __dadd_rn((SP)[-2 * kSyncOffset * kSyncThreads], done_); \
        }                                                                                                     \
    } while (0)
    int g = 0;
    if (B > 0) {
        // Ten steps per trip (lcm of the 5 ring slots and the 2 register buffers): every ring slot, score slot and
        // buffer role is a compile-time constant, so a step is arithmetic plus a handful of address-free loads.
        constexpr int RW = 32 + 7 * (B > 0 ? B : 1);
        for (; g + 10 <= last; g += 10) {
            double *const sp0 = scp + (g + 4) * kSyncThreads;
#define FT8B_STEP10(CUR, NXT, O)                                                                              \
    FT8B_SYNC_STEP(CUR, NXT, g + (O), ring_lane + (((O) + 4) % kSyncStages) * RW,                             \
                   ring_lane + (((O) + 1) % kSyncStages) * RW, sp0 + (O) * kSyncThreads, true)
            FT8B_STEP10(ca, cb, 0); FT8B_STEP10(cb, ca, 1); FT8B_STEP10(ca, cb, 2); FT8B_STEP10(cb, ca, 3);
            FT8B_STEP10(ca, cb, 4); FT8B_STEP10(cb, ca, 5); FT8B_STEP10(ca, cb, 6); FT8B_STEP10(cb, ca, 7);
            FT8B_STEP10(ca, cb, 8); FT8B_STEP10(cb, ca, 9);
#undef FT8B_STEP10
        }
    }
    // remaining columns (and the generic freq_osr build): ring slots tracked with wrapping pointers
    double *dst = ring_lane + ((g + kSyncStages - 1) % kSyncStages) * ring_w;
    const double *rd = ring_lane + ((g + 1) % kSyncStages) * ring_w;
    double *sp = scp + (g + 4) * kSyncThreads;
#define FT8B_ADVANCE()                                                  \
    do {                                                                \
        dst += ring_w; if (dst == ring_end) dst = ring_lane;            \
        rd += ring_w; if (rd == ring_end) rd = ring_lane;               \
        sp += kSyncThreads;                                             \
    } while (0)
    for (; g + 2 <= last; g += 2) {
        FT8B_SYNC_STEP(ca, cb, g, dst, rd, sp, true); FT8B_ADVANCE();
        FT8B_SYNC_STEP(cb, ca, g + 1, dst, rd, sp, true); FT8B_ADVANCE();
    }
    if (g < last) {                                      // one leftover step with a successor
        FT8B_SYNC_STEP(ca, cb, g, dst, rd, sp, true); FT8B_ADVANCE();
#pragma unroll
        for (int tau = 0; tau < 8; ++tau) ca[tau] = cb[tau];
        ++g;
    }
    FT8B_SYNC_STEP(ca, cb, g, dst, rd, sp, false);       // g == last
#undef FT8B_ADVANCE
#undef FT8B_SYNC_STEP
    __pipeline_wait_prior(0);
    if (!active) return;
    // partial blocks that run past the last symbol: h = nb-6 .. nb-1 live in acc[5] .. acc[0]
#pragma unroll
    for (int k = 5; k >= 0; --k) sync_scatter(scp, kSyncThreads, nb - 1 - k, nu, acc[k]);
    // normalise and store: score / num_comparisons, -inf when no comparison or non-finite (ft8_decode.py:97-100)
    double *out = a.scores + slot * (int64_t)(nu * s) * a.nf + (int64_t)rho * a.nf + f;
    const int64_t outstep = (int64_t)s * a.nf;
#pragma unroll 4
    for (int u = 0; u < nu; ++u) out[u * outstep] = sync_normalise(scp[u * kSyncThreads], ncmp[u]);
}

}  // namespace ft8b

#include "sync_pair.cuh"
#include "sync_lap.cuh"

namespace ft8b {

// ------------------------------------------------------------------------------------------
// S2b: selection.  One CTA per slot scans the score map in scan order (flat index = scan order).
// ------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;            // select_topk_kernel; select_candidates_kernel takes its own count
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelPerThread = 4;   // consecutive scores per thread per chunk

struct SelArgs {
    const double *scores;   // [n_slots][L]
    int64_t L;              // nt*nf
    int nf, t0, K, Kpad;
    double min_score;
    int32_t *cand_time, *cand_freq, *n_cand, *n_pass;
    double *cand_score;
    const SyncTile *tiles;  // [n_slots][n_tiles] strip arg-maxima from the sync kernel, or null: phase B scans the map
    int n_tiles;
    // Spectrogram the score map was made from, or null.  With it every score a DECISION or a RESULT hangs on is
    // re-evaluated as the reference's own running sum (ref_sync_score below): the returned cand_score, the order of the
    // list, threshold tests of scores within kSelTol of min_score, and the record replacement.
    const double *mag;      // [n_slots][T][F]
    int T, F, s, b, nb;
    // lists too long for shared memory (K > kSelSmemK): [n_slots][Kpad] scores / scan indices in global memory, else null
    double *big_s;
    long long *big_i;
};
constexpr int kSelSmemK = 4096;     // longest list the selection keeps in shared memory (64 KB)

// The sync kernels add per cell, then per Costas block, then per score (what lets neighbouring candidates share work);
// the reference keeps ONE running sum.  The two agree to ~1e-13 relative, not bit for bit.
constexpr double kSelTol = 1e-9;

// ft8_sync_score exactly as the reference evaluates it (ft8_decode.py:47-100): one running `score +=` over the Costas
// cells in (m, k) order, each cell adding p - lower tone, p - higher tone, p - earlier symbol, p - later symbol in that
// order, then ONE IEEE division by the comparison count.  Same bits as NumPy float64 on the same spectrogram.
__device__ double ref_sync_score(const double *mag, int F, int s, int b, int nb, int at, int af) {
    int base = at / s;                                     // Python floor division (:58)
    if ((at % s) != 0 && at < 0) --base;
    double score = 0.0;
    int n = 0;
    for (int m = 0; m < kNumSync; ++m) {
#pragma unroll
        for (int k = 0; k < kLenSync; ++k) {
            const int block = m * kSyncOffset + k;
            const int ba = base + block;
            if (ba < 0 || ba >= nb) continue;              // :68-69
            const int tone = kCostas[k];
            const double *cell = mag + (int64_t)(at + block * s) * F + af + tone * b;
            const double p = __ldg(cell);
            if (tone > 0) { score = __dadd_rn(score, __dsub_rn(p, __ldg(cell - b))); ++n; }                     // :76-78
            if (tone < 7) { score = __dadd_rn(score, __dsub_rn(p, __ldg(cell + b))); ++n; }                     // :80-82
            if (k > 0 && ba > 0) { score = __dadd_rn(score, __dsub_rn(p, __ldg(cell - (int64_t)s * F))); ++n; }           // :85-87
            if (k < kLenSync - 1 && ba + 1 < nb) { score = __dadd_rn(score, __dsub_rn(p, __ldg(cell + (int64_t)s * F))); ++n; }   // :89-91
        }
    }
    if (n == 0 || isnan(score) || isinf(score)) return -INFINITY;       // :94-95
    return __ddiv_rn(score, (double)n);
}

// The same value computed by a whole warp: the up to 105 spectrogram values of the score (21 Costas cells x {p, lower,
// higher, earlier, later}) are fetched in parallel, four per lane, then every lane forms the running sum in the
// reference's order from its neighbours' registers - the loads of one score overlap instead of following each other
// down a chain of 84 dependent additions (one thread per score: 1.2 ms per 4096 slots; this: see DESIGN.md).
__device__ double ref_sync_score_warp(const double *mag, int F, int s, int b, int nb, int at, int af, int lane) {
    constexpr int kC[7] = {3, 1, 4, 0, 6, 5, 2};
    int base = at / s;
    if ((at % s) != 0 && at < 0) --base;
    double r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int v = lane + 32 * q;                       // value index = 5 * cell + j
        double x = 0.0;
        if (v < 105) {
            const int cell = v / 5, j = v - 5 * cell;
            const int m = cell / 7, k = cell - 7 * m;
            const int block = m * kSyncOffset + k, ba = base + block;
            const int tone = kCostas[k];
            bool ok = ba >= 0 && ba < nb;
            int64_t col = at + block * s;
            int bin = af + tone * b;
            if (j == 1) { ok = ok && tone > 0; bin -= b; }
            else if (j == 2) { bin += b; }
            else if (j == 3) { ok = ok && k > 0 && ba > 0; col -= s; }
            else if (j == 4) { ok = ok && k < kLenSync - 1 && ba + 1 < nb; col += s; }
            if (ok) x = __ldg(mag + col * F + bin);
        }
        r[q] = x;
    }
    double score = 0.0;
    int n = 0;
#pragma unroll
    for (int cell = 0; cell < 21; ++cell) {
        const int m = cell / 7, k = cell % 7;
        const int ba = base + m * kSyncOffset + k;
        // every lane evaluates every shuffle (no divergence around them); the terms are only added where the reference adds
        const double p = __shfl_sync(0xffffffffu, r[(5 * cell) >> 5], (5 * cell) & 31);
        const double lo = __shfl_sync(0xffffffffu, r[(5 * cell + 1) >> 5], (5 * cell + 1) & 31);
        const double hi = __shfl_sync(0xffffffffu, r[(5 * cell + 2) >> 5], (5 * cell + 2) & 31);
        const double ea = __shfl_sync(0xffffffffu, r[(5 * cell + 3) >> 5], (5 * cell + 3) & 31);
        const double la = __shfl_sync(0xffffffffu, r[(5 * cell + 4) >> 5], (5 * cell + 4) & 31);
        if (ba < 0 || ba >= nb) continue;
        if (kC[k] > 0) { score = __dadd_rn(score, __dsub_rn(p, lo)); ++n; }
        { score = __dadd_rn(score, __dsub_rn(p, hi)); ++n; }
        if (k > 0 && ba > 0) { score = __dadd_rn(score, __dsub_rn(p, ea)); ++n; }
        if (k < kLenSync - 1 && ba + 1 < nb) { score = __dadd_rn(score, __dsub_rn(p, la)); ++n; }
    }
    if (n == 0 || isnan(score) || isinf(score)) return -INFINITY;
    return __ddiv_rn(score, (double)n);
}

__device__ __forceinline__ bool sel_pass(double v, double thr) {
    // ft8_decode.py:127: skip if score == -inf or score < min_score
    return (v != -INFINITY) && !(v < thr) && !isnan(v);
}
// ordering used for max/sort: higher score first, then earlier scan position
__device__ __forceinline__ bool sel_better(double sa, long long ia, double sb, long long ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

// Bitonic sort of the candidate list over Kpad entries.  BY_INDEX = false: (score descending, tag ascending) - the
// order of the result; BY_INDEX = true: tag ascending only (scan order = the order the reference pushed them in).
template <int THREADS, bool BY_INDEX>
__device__ __forceinline__ void sel_bitonic(double *hs, long long *hi, int Kpad, int tid) {
    for (int size = 2; size <= Kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < Kpad; i += THREADS) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = ((i & size) == 0);   // "up" = better elements first
                    const double si = hs[i], sj = hs[j];
                    const long long ii = hi[i], ij = hi[j];
                    const bool j_better = BY_INDEX ? (ij < ii) : sel_better(sj, ij, si, ii);
                    if (up ? j_better : !j_better) { hs[i] = sj; hs[j] = si; hi[i] = ij; hi[j] = ii; }
                }
            }
            __syncthreads();
        }
    }
}

// Exact score ties inside the list (two candidates whose only in-range Costas blocks are the same cells - it happens
// with time-cropped or short spectrograms - or digital silence).  The reference sorts its heap ARRAY with a stable
// key sort (ft8_decode.py:140), so tied candidates come out in heap-array order, and a tie that heapq itself compares
// raises (the tuple comparison falls through to FT8Candidate).  This replays CPython's heapq on the list: hs / hi hold
// the first n passers in push (scan) order on entry and the heap array on exit.  `replace`: a later record evicts the
// root (ft8_decode.py:136-137) - the sift path of heapreplace only has to be checked for compared ties, the
// arrangement below the root is unchanged because the new key is strictly the smallest.  Returns false when the
// reference would have raised (a compared tie); the arrays are then some permutation of the input.
__device__ bool sel_heap_replay(double *hs, long long *hi, int n, bool replace, double rec_s, long long rec_i) {
    for (int m = 1; m < n; ++m) {                           // heappush: heapq._siftdown(heap, 0, m)
        const double s = hs[m];
        const long long id = hi[m];
        int pos = m;
        bool tie = false;
        while (pos > 0) {
            const int parent = (pos - 1) >> 1;
            const double ps = hs[parent];
            if (s == ps) { tie = true; break; }             // equal keys: the tuple comparison reaches the candidates
            if (!(s > ps)) break;                           // newitem < parent  <=>  -s < -ps
            hs[pos] = ps; hi[pos] = hi[parent];
            pos = parent;
        }
        hs[pos] = s; hi[pos] = id;
        if (tie) return false;
    }
    if (replace) {                                          // heapreplace: heapq._siftup(heap, 0) walks the smaller-key children
        int child = 1;
        while (child < n) {
            const int right = child + 1;
            if (right < n) {
                if (hs[child] == hs[right]) return false;
                if (!(hs[child] > hs[right])) child = right;
            }
            child = 2 * child + 1;
        }
        hs[0] = rec_s; hi[0] = rec_i;
    }
    return true;
}
constexpr int kSelTagShift = 40;                            // hi = (heap position << 40) | scan index in the tie path
constexpr long long kSelIdxMask = (1LL << kSelTagShift) - 1;

// THREADS: 256 for short lists (K <= 64: four CTAs per SM hide each other's barrier and DRAM latencies - 0.235 ms per
// 4096 slots at K = 20 against 0.47 ms with one 1024-thread CTA per SM), 1024 for long ones (the sort and the
// compaction of thousands of candidates want the threads).
#ifndef FT8B_SEL_MINB
#define FT8B_SEL_MINB 5
#endif
// BIG: the list lives in the global scratch (K > kSelSmemK); a compile-time switch, so that the common case keeps plain
// shared-memory accesses (generic pointers here cost 0.03 ms per 4096 slots)
template <int THREADS, bool BIG = false>
__global__ void __launch_bounds__(THREADS, THREADS <= 256 ? FT8B_SEL_MINB : 1) select_candidates_kernel(const SelArgs a) {   // 5 x 256 threads per SM: 48 registers
    constexpr int kSelThreads = THREADS, kSelWarps = THREADS / 32;
    extern __shared__ unsigned char sel_smem[];
    const int64_t slot = blockIdx.x;
    // [Kpad] scores and scan indices: shared memory, or (K > kSelSmemK) this slot's rows of the global scratch
    double *hs = BIG ? a.big_s + slot * a.Kpad : reinterpret_cast<double *>(sel_smem);
    long long *hi = BIG ? a.big_i + slot * a.Kpad : reinterpret_cast<long long *>(reinterpret_cast<double *>(sel_smem) + a.Kpad);
    __shared__ int warp_cnt[32];
    __shared__ int s_base, s_done;
    __shared__ long long s_cut;
    __shared__ double red_s[32];
    __shared__ long long red_i[32];
    __shared__ long long red_c[32];
    __shared__ double s_rec_s;          // the later record (phase B): score and scan index
    __shared__ long long s_rec_i;
    __shared__ double s_ev_s;           // the element the record evicted in B3 (needed again if the list holds exact ties)
    __shared__ long long s_ev_i;
    __shared__ int s_replaced;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *sc = a.scores + slot * a.L;
    const double *mag = a.mag ? a.mag + slot * (int64_t)a.T * a.F : nullptr;
    const int K = a.K;

    for (int i = tid; i < a.Kpad; i += kSelThreads) { hs[i] = -INFINITY; hi[i] = 0x7fffffffffffffffLL; }
    if (tid == 0) { s_base = 0; s_done = 0; s_cut = -1; s_replaced = 0; }
    __syncthreads();

    // ---- phase A: first K passers in scan order (ordered compaction) ----
    const int64_t chunk = (int64_t)kSelThreads * kSelPerThread;
    int64_t pos = 0;
    long long total_pass = 0;   // only meaningful when the scan runs to the end (K never reached)
    for (; pos < a.L; pos += chunk) {
        const int64_t i0 = pos + (int64_t)tid * kSelPerThread;
        double v[kSelPerThread];
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < kSelPerThread; ++j) {
            v[j] = (i0 + j < a.L) ? __ldg(sc + i0 + j) : -INFINITY;
            // a score this close to the threshold is decided on the reference's own sum (rare: ~1e-9 of all scores)
            if (mag && fabs(v[j] - a.min_score) <= kSelTol * (1.0 + fabs(a.min_score)))
                v[j] = ref_sync_score(mag, a.F, a.s, a.b, a.nb, a.t0 + (int)((i0 + j) / a.nf), (int)((i0 + j) % a.nf));
            cnt += sel_pass(v[j], a.min_score) ? 1 : 0;
        }
        // exclusive prefix of cnt over the block (thread order = scan order)
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) warp_cnt[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = (lane < kSelWarps) ? warp_cnt[lane] : 0;
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += n;
            }
            warp_cnt[lane] = wi - w;          // exclusive warp offsets
            if (lane == 31) s_done = wi;      // chunk total (reuse s_done as scratch)
        }
        __syncthreads();
        const int base = s_base;
        int rank = base + warp_cnt[wid] + (incl - cnt);
#pragma unroll
        for (int j = 0; j < kSelPerThread; ++j) {
            if (sel_pass(v[j], a.min_score)) {
                if (rank < K) { hs[rank] = v[j]; hi[rank] = i0 + j; }
                if (rank == K - 1) s_cut = i0 + j;
                ++rank;
            }
        }
        const int chunk_total = s_done;
        __syncthreads();
        if (tid == 0) s_base = base + chunk_total;
        total_pass = (long long)base + chunk_total;
        __syncthreads();
        if (base + chunk_total >= K) { pos += chunk; break; }
    }
    const int n_first = (s_base < K) ? s_base : K;
    const long long cut = s_cut;    // scan index of the K-th passer, -1 if fewer than K passers
    // ---- phase B: running-record replacement over the passers after the K-th ----
    // B1: the later record = first occurrence of the maximum over the passers beyond the cut -> (s_rec_s, s_rec_i)
    long long n_pass_total = total_pass;
    const bool have_later = cut >= 0 && K > 0;
    if (have_later && a.tiles) {
        // The later record is the first occurrence of the map's maximum whenever that lies beyond the cut; if it lies
        // before the cut it is already in H as its best element and nothing later is strictly greater.  So the
        // maximum over the strip summaries decides, and the rest of the map is not read (n_pass is not available).
        double bs = -INFINITY;
        long long bi = 0x7fffffffffffffffLL;
        const SyncTile *tl = a.tiles + slot * a.n_tiles;
        for (int i = tid; i < a.n_tiles; i += kSelThreads) {
            const double v = tl[i].score;
            const long long vi = tl[i].idx;
            if (sel_better(v, vi, bs, bi)) { bs = v; bi = vi; }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, d);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (sel_better(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) { red_s[wid] = bs; red_i[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            bs = (lane < kSelWarps) ? red_s[lane] : -INFINITY;
            bi = (lane < kSelWarps) ? red_i[lane] : 0x7fffffffffffffffLL;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, bs, d);
                const long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
                if (sel_better(os, oi, bs, bi)) { bs = os; bi = oi; }
            }
            if (lane == 0) {
                const bool live = bi > cut && bi != 0x7fffffffffffffffLL;
                s_rec_s = live ? bs : -INFINITY;
                s_rec_i = live ? bi : 0x7fffffffffffffffLL;
            }
        }
        n_pass_total = -1;
    } else if (have_later) {
        double bs = -INFINITY;
        long long bi = 0x7fffffffffffffffLL;
        long long extra = 0;
        for (int64_t i = cut + 1 + tid; i < a.L; i += kSelThreads) {
            double v = __ldg(sc + i);
            if (mag && fabs(v - a.min_score) <= kSelTol * (1.0 + fabs(a.min_score)))
                v = ref_sync_score(mag, a.F, a.s, a.b, a.nb, a.t0 + (int)(i / a.nf), (int)(i % a.nf));
            if (sel_pass(v, a.min_score)) {
                ++extra;
                if (sel_better(v, i, bs, bi)) { bs = v; bi = i; }
            }
        }
        // block arg-max (first occurrence on ties) and count
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, d);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
            extra += __shfl_xor_sync(0xffffffffu, extra, d);
            if (sel_better(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) { red_s[wid] = bs; red_i[wid] = bi; red_c[wid] = extra; }
        __syncthreads();
        if (wid == 0) {
            bs = (lane < kSelWarps) ? red_s[lane] : -INFINITY;
            bi = (lane < kSelWarps) ? red_i[lane] : 0x7fffffffffffffffLL;
            extra = (lane < kSelWarps) ? red_c[lane] : 0;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, bs, d);
                const long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
                extra += __shfl_xor_sync(0xffffffffu, extra, d);
                if (sel_better(os, oi, bs, bi)) { bs = os; bi = oi; }
            }
            if (lane == 0) { s_rec_s = bs; s_rec_i = bi; red_c[0] = extra; }
        }
        __syncthreads();
        // passers counted so far: everything up to and including the cut chunk is not fully counted;
        // recompute as K-th passer rank (K) + later passers.
        n_pass_total = (long long)K + red_c[0];
    } else if (tid == 0) {
        s_rec_s = -INFINITY;
        s_rec_i = 0x7fffffffffffffffLL;
    }
    __syncthreads();
    // B2: with the spectrogram at hand, the scores of the list and of the later record as the reference computes them
    // (what is returned, what the replacement compares and what the final sort orders by): one warp per score, the
    // 105 spectrogram values of a score loaded in parallel, the running sum then formed in the reference's order.
    if (mag) {
        const long long rec_i = s_rec_i;
        const int n_items = n_first + (rec_i != 0x7fffffffffffffffLL ? 1 : 0);
        if (THREADS <= 256) {
            // short lists: one warp per score (latency of ONE score matters: 21 warps cover K = 20 in three rounds)
            for (int it = wid; it < n_items; it += kSelWarps) {
                const long long idx = (it < n_first) ? hi[it] : rec_i;
                const double e = ref_sync_score_warp(mag, a.F, a.s, a.b, a.nb, a.t0 + (int)(idx / a.nf), (int)(idx % a.nf), lane);
                if (lane == 0) {
                    if (it < n_first) hs[it] = e; else s_rec_s = e;
                }
            }
        } else {
            // long lists (K up to 4096): one THREAD per score - 1024 running sums in flight hide each other's load
            // latency (K = 2000: 63 warp rounds of ~6 us against two thread rounds)
            for (int it = tid; it < n_items; it += kSelThreads) {
                const long long idx = (it < n_first) ? hi[it] : rec_i;
                const double e = ref_sync_score(mag, a.F, a.s, a.b, a.nb, a.t0 + (int)(idx / a.nf), (int)(idx % a.nf));
                if (it < n_first) hs[it] = e; else s_rec_s = e;
            }
        }
        __syncthreads();
    }
    // B3: b = best of H (heap root); replaced when the later record is strictly greater (:136-137)
    if (tid == 0 && have_later && s_rec_i != 0x7fffffffffffffffLL && sel_pass(s_rec_s, a.min_score)) {
        int best = 0;
        for (int i = 1; i < n_first; ++i)
            if (sel_better(hs[i], hi[i], hs[best], hi[best])) best = i;
        if (s_rec_s > hs[best]) { s_ev_s = hs[best]; s_ev_i = hi[best]; s_replaced = 1; hs[best] = s_rec_s; hi[best] = s_rec_i; }
    }
    __syncthreads();

    // ---- phase C: sort H by (score desc, scan index asc) - bitonic over the occupied power-of-two prefix
    // (entries n_first .. Kpad-1 are padding: -inf, largest index)
    int P = 2;
    while (P < n_first) P <<= 1;
    sel_bitonic<THREADS, false>(hs, hi, P, tid);
    // ---- phase D (rare): exact ties in the list come out in the reference's heap-array order ----
    int tie = 0;
    for (int i = tid; i + 1 < n_first; i += kSelThreads) tie |= (hs[i] == hs[i + 1]) ? 1 : 0;
    if (__syncthreads_or(tie)) {
        const bool replaced = s_replaced != 0;
        if (replaced)                                       // back to the K elements that were pushed
            for (int i = tid; i < n_first; i += kSelThreads)
                if (hi[i] == s_rec_i) { hs[i] = s_ev_s; hi[i] = s_ev_i; }
        __syncthreads();
        sel_bitonic<THREADS, true>(hs, hi, P, tid);         // push order
        if (tid == 0) {
            if (sel_heap_replay(hs, hi, n_first, replaced, s_rec_s, s_rec_i)) {
                for (int i = 0; i < n_first; ++i) hi[i] |= (long long)i << kSelTagShift;
            } else if (replaced) {
                // the reference raises here; what is returned instead is the scan-order list of B3
                int best = 0;
                for (int i = 1; i < n_first; ++i)
                    if (sel_better(hs[i], hi[i], hs[best], hi[best])) best = i;
                hs[best] = s_rec_s; hi[best] = s_rec_i;
            }
        }
        __syncthreads();
        sel_bitonic<THREADS, false>(hs, hi, P, tid);
    }
    for (int i = tid; i < K; i += kSelThreads) {
        const int64_t o = slot * K + i;
        if (i < n_first) {
            const long long idx = hi[i] & kSelIdxMask;
            a.cand_time[o] = a.t0 + (int)(idx / a.nf);
            a.cand_freq[o] = (int)(idx % a.nf);
            a.cand_score[o] = hs[i];
        } else {
            a.cand_time[o] = 0; a.cand_freq[o] = 0; a.cand_score[o] = 0.0;
        }
    }
    if (tid == 0) {
        a.n_cand[slot] = n_first;
        if (a.n_pass) a.n_pass[slot] = (int32_t)(n_pass_total > 0x7fffffffLL ? 0x7fffffffLL : n_pass_total);
    }
}

// ------------------------------------------------------------------------------------------
// Strict top-K selection - OPT-IN EXTENSION, not the reference's behaviour (the reference keeps the
// first K passers in scan order plus one record replacement, ft8_decode.py:127-140; with one strong signal and
// min_score = 1 that list is the signal plus K-1 noise cells from the first time row).  Here: the K best passers by
// (score descending, scan position ascending), in that order.  One CTA per slot: radix select (8 bits per pass over an
// order-preserving 64-bit key) finds the K-th best score, one more ordered pass gathers everything above it plus the
// earliest ties, then the same bitonic sort as the reference-semantics kernel.
__device__ __forceinline__ unsigned long long sel_key(double v) {          // larger double <=> larger key
    const unsigned long long b = (unsigned long long)__double_as_longlong(__dadd_rn(v, 0.0));     // -0.0 -> +0.0: equal scores, equal keys
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(const SelArgs a) {
    extern __shared__ unsigned char sel_smem[];
    double *hs = reinterpret_cast<double *>(sel_smem);                   // [Kpad] scores
    long long *hi = reinterpret_cast<long long *>(hs + a.Kpad);          // [Kpad] scan indices
    __shared__ int hist[256];
    __shared__ int warp_cnt[32];
    __shared__ int s_bucket, s_need, s_base, s_tot, s_fill;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t slot = blockIdx.x;
    const double *sc = a.scores + slot * a.L;
    const int K = a.K;
    for (int i = tid; i < a.Kpad; i += kSelThreads) { hs[i] = -INFINITY; hi[i] = 0x7fffffffffffffffLL; }
    if (tid == 0) { s_need = K; s_fill = 0; s_base = 0; }

    // ---- radix select: after the pass at `shift`, `prefix` holds the top bits of the K-th best key ----
    unsigned long long prefix = 0, mask = 0;
    bool take_all = false;          // fewer than K passers: everything that passes is taken
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
        __syncthreads();
        for (int64_t i = tid; i < a.L; i += kSelThreads) {
            const double v = __ldg(sc + i);
            if (sel_pass(v, a.min_score)) {
                const unsigned long long k = sel_key(v);
                if ((k & mask) == prefix) atomicAdd(&hist[(int)(k >> shift) & 255], 1);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int need = s_need, cum = 0, b = 255;
            for (; b >= 0; --b) {
                if (cum + hist[b] >= need) break;
                cum += hist[b];
            }
            s_bucket = b;                        // -1: fewer than `need` keys carry this prefix (only on the first pass)
            if (b >= 0) s_need = need - cum;     // still to take from bucket b
        }
        __syncthreads();
        if (s_bucket < 0) { take_all = true; break; }
        prefix |= (unsigned long long)s_bucket << shift;
        mask |= 0xffull << shift;
    }
    const int need_ties = take_all ? 0 : s_need;                 // keys equal to the threshold to take, earliest first
    __syncthreads();

    // ---- ordered gather: key > threshold always, key == threshold while the tie budget lasts (scan order) ----
    const int64_t chunk = (int64_t)kSelThreads * kSelPerThread;
    for (int64_t pos = 0; pos < a.L; pos += chunk) {
        const int64_t i0 = pos + (int64_t)tid * kSelPerThread;
        double v[kSelPerThread];
        int cls[kSelPerThread];                                   // 0 skip, 1 take, 2 tie
        int ties = 0;
#pragma unroll
        for (int j = 0; j < kSelPerThread; ++j) {
            v[j] = (i0 + j < a.L) ? __ldg(sc + i0 + j) : -INFINITY;
            cls[j] = 0;
            if (sel_pass(v[j], a.min_score)) {
                const unsigned long long k = sel_key(v[j]);
                cls[j] = take_all ? 1 : (k > prefix ? 1 : (k == prefix ? 2 : 0));
            }
            ties += cls[j] == 2;
        }
        // exclusive prefix of the tie counts over the block (thread order = scan order)
        int incl = ties;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) warp_cnt[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int w = (lane < kSelWarps) ? warp_cnt[lane] : 0;
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += n;
            }
            warp_cnt[lane] = wi - w;
            if (lane == 31) s_tot = wi;
        }
        __syncthreads();
        int tie_rank = s_base + warp_cnt[wid] + (incl - ties);
#pragma unroll
        for (int j = 0; j < kSelPerThread; ++j) {
            bool take = cls[j] == 1;
            if (cls[j] == 2) { take = tie_rank < need_ties; ++tie_rank; }
            if (take) {
                const int o = atomicAdd(&s_fill, 1);              // any order: the sort below is a total order
                if (o < a.Kpad) { hs[o] = v[j]; hi[o] = i0 + j; }
            }
        }
        __syncthreads();
        if (tid == 0) s_base += s_tot;
        __syncthreads();
    }
    const int n_first = s_fill < K ? s_fill : K;

    // ---- sort by (score desc, scan index asc) - bitonic over Kpad ----
    for (int size = 2; size <= a.Kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < a.Kpad; i += kSelThreads) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = ((i & size) == 0);
                    const double si = hs[i], sj = hs[j];
                    const long long ii = hi[i], ij = hi[j];
                    const bool j_better = sel_better(sj, ij, si, ii);
                    if (up ? j_better : !j_better) { hs[i] = sj; hs[j] = si; hi[i] = ij; hi[j] = ii; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += kSelThreads) {
        const int64_t o = slot * K + i;
        if (i < n_first) {
            a.cand_time[o] = a.t0 + (int)(hi[i] / a.nf);
            a.cand_freq[o] = (int)(hi[i] % a.nf);
            a.cand_score[o] = hs[i];
        } else {
            a.cand_time[o] = 0; a.cand_freq[o] = 0; a.cand_score[o] = 0.0;
        }
    }
    if (tid == 0) a.n_cand[slot] = n_first;
}

// ------------------------------------------------------------------------------------------
// FT8B_SYNC_LAP=1: freq_osr 2 runs on the path-sum Laplacian kernel (sync_lap.cuh) instead of the difference walk of
// sync_pair.cuh.  An independent formulation of the same map (40 % fewer float64 additions, but more shared-memory /
// shuffle traffic and 15 % more warp-steps: 5.79 ms against 5.19 ms per 4096 slots, so it is not the default); kept as
// the cross-check of the walk in the tests.
static bool sync_uses_lap(int b) {
    const char *env_lap = getenv("FT8B_SYNC_LAP");
    return b == 2 && env_lap && env_lap[0] == '1';
}

// Shared memory the walk kernels need for nu time offsets per phase (their score sums live there), and what a CTA may
// have.  Captures longer than ~42 s (nu > 213 at freq_osr 2) do not fit: they run on sync_scores_direct_kernel below.
static size_t sync_walk_smem(int nu, int b, bool pair, bool lap) {
    if (lap) return sync_lap_smem_bytes(nu);
    if (pair) return b == 1 ? sync_pair_smem_bytes<1>(nu) : b == 2 ? sync_pair_smem_bytes<2>(nu) : sync_pair_smem_bytes<4>(nu);
    return sizeof(double) * ((size_t)nu * kSyncThreads + (size_t)(kSyncThreads / 32) * kSyncStages * (32 + 7 * b)) +
           sizeof(double2) * (size_t)nu;
}
static size_t sync_smem_limit() {
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        cudaGetLastError();              // host-only workspace queries (no device): the sm_100 figure, 227 KB
        return 227 * 1024;
    }
    return (size_t)max_smem;
}

// ft8_sync_score for EVERY position of the search range, one thread per score, each the reference's own running sum
// (ref_sync_score): no shared state, any number of time offsets - the path of captures too long for the walk kernels'
// score sums (and bit-equal to the reference on the same spectrogram, where the walks are within 1e-12).
__global__ void sync_scores_direct_kernel(const SyncArgs a, int t0) {
    const int64_t per_slot = (int64_t)a.nu * a.s * a.nf;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= per_slot) return;
    const int64_t slot = blockIdx.y;
    const int row = (int)(i / a.nf), f = (int)(i - (int64_t)row * a.nf);
    a.scores[slot * per_slot + i] = ref_sync_score(a.mag + slot * (int64_t)a.T * a.F, a.F, a.s, a.b, a.nb, t0 + row, f);
}

// strips per slot the pair kernels summarise (0 when the geometry runs on a kernel that leaves none)
int sync_tiles_per_slot(int T, int F, int s, int b) {
    int t0, nt, nf;
    ft8b_search_ranges(T, F, s, b, &t0, &nt, &nf);
    const char *env_single = getenv("FT8B_SYNC_SINGLE");
    const char *env_scan = getenv("FT8B_SELECT_FULLSCAN");     // "1": A/B switch, the selection scans the whole score map
    if (nt <= 0 || nf <= 0 || !(b == 1 || b == 2 || b == 4) || (env_single && env_single[0] == '1') ||
        (env_scan && env_scan[0] == '1'))
        return 0;
    if (sync_walk_smem(nt / s, b, true, sync_uses_lap(b)) > sync_smem_limit()) return 0;      // direct kernel: no strips
    if (sync_uses_lap(b)) return s * ((sync_lap_warps(nf) + 1) / 2) * (kLapThreads / 32);
    return s * ((nf + kPairCols - 1) / kPairCols) * (kPairThreads / 32);
}

int sync_scores(const double *mag, int64_t n_slots, int T, int F, int s, int b, double *scores, SyncTile *tiles,
                cudaStream_t stream) {
    if (!mag || !scores) FT8B_ARG_FAIL("ft8b_sync_scores: null pointer");
    if (T < 0 || F < 0 || s < 1 || b < 1) FT8B_ARG_FAIL("ft8b_sync_scores: bad shape");
    int t0, nt, nf;
    ft8b_search_ranges(T, F, s, b, &t0, &nt, &nf);
    if (n_slots <= 0 || nt <= 0 || nf <= 0) return 0;
    SyncArgs a;
    a.mag = mag; a.scores = scores; a.T = T; a.F = F; a.s = s; a.b = b;
    a.nb = T / s; a.nu = nt / s; a.nf = nf;
    a.tiles = tiles;
    int dev = 0, max_smem = 0;
    FT8B_CUDA_TRY(cudaGetDevice(&dev));
    FT8B_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (s > 65535 || n_slots > 65535) {
        set_error("ft8b_sync_scores: steps_per_symbol and slots per call must be <= 65535");
        return FT8B_ERR_UNSUPPORTED;
    }
    const char *env_single = getenv("FT8B_SYNC_SINGLE");            // "1": one candidate per thread for every freq_osr
    const bool single = env_single && env_single[0] == '1';
    const bool pair = (b == 1 || b == 2 || b == 4) && !single;      // two candidates per thread (sync_pair.cuh)
    if (sync_walk_smem(a.nu, b, pair, pair && sync_uses_lap(b)) > (size_t)max_smem) {
        // too many time offsets for the score sums of a walk (a capture of more than ~42 s): one thread per score
        if (tiles) FT8B_ARG_FAIL("sync_scores: strip summaries are only produced by the pair kernels");
        const int64_t per_slot = (int64_t)nt * nf;
        if (n_slots > 65535) {
            set_error("ft8b_sync_scores: slots per call must be <= 65535");
            return FT8B_ERR_UNSUPPORTED;
        }
        dim3 grid((unsigned)((per_slot + 255) / 256), (unsigned)n_slots);
        sync_scores_direct_kernel<<<grid, 256, 0, stream>>>(a, t0);
        FT8B_LAUNCH_CHECK("sync_scores_direct_kernel");
        return 0;
    }
    if (pair && sync_uses_lap(b)) {
        const size_t smem_l = sync_lap_smem_bytes(a.nu);
        if (smem_l > (size_t)max_smem) {
            set_error("ft8b_sync_scores: %d time offsets per phase need %zu bytes of shared memory (limit %d)", a.nu, smem_l, max_smem);
            return FT8B_ERR_UNSUPPORTED;
        }
        dim3 grid((sync_lap_warps(nf) + 1) / 2, s, (unsigned)n_slots);
#define FT8B_LAP_LAUNCH(TR, NBC)                                                                                                 \
    do {                                                                                                                         \
        FT8B_CUDA_TRY(cudaFuncSetAttribute(sync_scores_lap_kernel<TR, NBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem)); \
        sync_scores_lap_kernel<TR, NBC><<<grid, kLapThreads, smem_l, stream>>>(a);                                               \
    } while (0)
        // 93 symbol columns per phase = a 15-s slot at any sample rate (186 frames, time_osr 2): unrolled head / tail
        if (a.nb == 93) { if (tiles) FT8B_LAP_LAUNCH(true, 93); else FT8B_LAP_LAUNCH(false, 93); }
        else            { if (tiles) FT8B_LAP_LAUNCH(true, 0); else FT8B_LAP_LAUNCH(false, 0); }
#undef FT8B_LAP_LAUNCH
        FT8B_LAUNCH_CHECK("sync_scores_lap_kernel");
        return 0;
    }
    const size_t smem = pair ? (b == 1 ? sync_pair_smem_bytes<1>(a.nu) : b == 2 ? sync_pair_smem_bytes<2>(a.nu) : sync_pair_smem_bytes<4>(a.nu))
                             : sizeof(double) * ((size_t)a.nu * kSyncThreads + (size_t)(kSyncThreads / 32) * kSyncStages * (32 + 7 * b)) +
                                   sizeof(double2) * (size_t)a.nu;
    if (smem > (size_t)max_smem) {
        set_error("ft8b_sync_scores: %d time offsets per phase need %zu bytes of shared memory (limit %d)", a.nu, smem, max_smem);
        return FT8B_ERR_UNSUPPORTED;
    }
#define FT8B_SYNC_LAUNCH(KERNEL, THREADS, COLS)                                                                      \
    do {                                                                                                             \
        dim3 grid((nf + (COLS) - 1) / (COLS), s, (unsigned)n_slots);                                                 \
        FT8B_CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));     \
        KERNEL<<<grid, THREADS, smem, stream>>>(a);                                                                  \
    } while (0)
    if (pair) {
        switch (b) {
#define FT8B_PAIR_LAUNCH(BB)                                                                                          \
    do {                                                                                                             \
        auto kt = sync_scores_pair_kernel<BB, true>;                                                                 \
        auto kf = sync_scores_pair_kernel<BB, false>;                                                                \
        auto kern = tiles ? kt : kf;                                                                                 \
        FT8B_SYNC_LAUNCH(kern, kPairThreads, kPairCols);                                                             \
    } while (0)
            case 1: FT8B_PAIR_LAUNCH(1); break;
            case 2: FT8B_PAIR_LAUNCH(2); break;
            default: FT8B_PAIR_LAUNCH(4); break;
#undef FT8B_PAIR_LAUNCH
        }
    } else {
        if (tiles) FT8B_ARG_FAIL("sync_scores: strip summaries are only produced by the pair kernels");
        switch (b) {       // freq_osr as a compile-time constant turns the 8 tone offsets into immediates
            case 1: FT8B_SYNC_LAUNCH(sync_scores_kernel<1>, kSyncThreads, kSyncThreads); break;
            case 2: FT8B_SYNC_LAUNCH(sync_scores_kernel<2>, kSyncThreads, kSyncThreads); break;
            case 4: FT8B_SYNC_LAUNCH(sync_scores_kernel<4>, kSyncThreads, kSyncThreads); break;
            default: FT8B_SYNC_LAUNCH(sync_scores_kernel<0>, kSyncThreads, kSyncThreads); break;
        }
    }
#undef FT8B_SYNC_LAUNCH
    FT8B_LAUNCH_CHECK("sync_scores_kernel");
    return 0;
}

// ft8_sync_score at given positions, one thread per score, the reference's own running sum (ref_sync_score above)
__global__ void sync_score_at_kernel(const double *mag, int64_t n_slots, int T, int F, int s, int b, const int32_t *ct,
                                     const int32_t *cf, int n_per_slot, double *out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_slots * n_per_slot) return;
    const int64_t slot = i / n_per_slot;
    const int at = ct[i], af = cf[i];
    if (af < 0 || af + 7 * b >= F) { out[i] = nan(""); return; }
    out[i] = ref_sync_score(mag + slot * (int64_t)T * F, F, s, b, T / s, at, af);
}

int sync_score_at(const double *mag, int64_t n_slots, int T, int F, int s, int b, const int32_t *ct, const int32_t *cf,
                  int n_per_slot, double *score, cudaStream_t stream) {
    if (!mag || !ct || !cf || !score) FT8B_ARG_FAIL("ft8b_sync_score_at: null pointer");
    if (T < 0 || F < 0 || s < 1 || b < 1 || n_per_slot < 0 || n_slots < 0) FT8B_ARG_FAIL("ft8b_sync_score_at: bad shape");
    const int64_t n = n_slots * n_per_slot;
    if (n == 0) return 0;
    sync_score_at_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(mag, n_slots, T, F, s, b, ct, cf, n_per_slot, score);
    FT8B_LAUNCH_CHECK("sync_score_at_kernel");
    return 0;
}

int select_candidates(const double *scores, int64_t n_slots, int nt, int nf, int t0, int K, double min_score,
                      int32_t *cand_time, int32_t *cand_freq, double *cand_score, int32_t *n_cand, int32_t *n_pass,
                      const SyncTile *tiles, int n_tiles, const double *mag, int T, int F, int s, int b,
                      cudaStream_t stream) {
    if (tiles && n_pass) FT8B_ARG_FAIL("select_candidates: n_pass needs the full scan (no strip summaries)");
    if (!cand_time || !cand_freq || !cand_score || !n_cand) FT8B_ARG_FAIL("ft8b_select_candidates: null pointer");
    if (K < 0 || K > FT8B_MAX_CANDIDATES || (n_tiles < 0 && K > kSelSmemK)) {
        set_error("ft8b_select_candidates: K=%d outside [0,%d]", K, n_tiles < 0 ? kSelSmemK : FT8B_MAX_CANDIDATES);
        return FT8B_ERR_UNSUPPORTED;
    }
    if (n_slots <= 0) return 0;
    if (nt <= 0 || nf <= 0 || K == 0) {
        FT8B_CUDA_TRY(cudaMemsetAsync(n_cand, 0, sizeof(int32_t) * n_slots, stream));
        if (n_pass) FT8B_CUDA_TRY(cudaMemsetAsync(n_pass, 0, sizeof(int32_t) * n_slots, stream));
        return 0;
    }
    if (!scores) FT8B_ARG_FAIL("ft8b_select_candidates: null scores");
    SelArgs a;
    a.scores = scores; a.L = (int64_t)nt * nf; a.nf = nf; a.t0 = t0; a.K = K;
    int kp = 2;
    while (kp < K) kp <<= 1;
    a.Kpad = kp;
    a.min_score = min_score;
    a.cand_time = cand_time; a.cand_freq = cand_freq; a.cand_score = cand_score; a.n_cand = n_cand; a.n_pass = n_pass;
    a.tiles = tiles; a.n_tiles = n_tiles;
    a.mag = mag; a.T = T; a.F = F; a.s = s; a.b = b; a.nb = (mag && s > 0) ? T / s : 0;
    if (mag) {
        int t0c, ntc, nfc;
        ft8b_search_ranges(T, F, s, b, &t0c, &ntc, &nfc);
        if (t0c != t0 || ntc != nt || nfc != nf) FT8B_ARG_FAIL("select_candidates: score map does not belong to this spectrogram geometry");
    }
    a.big_s = nullptr; a.big_i = nullptr;
    if (K > kSelSmemK) {
        // a list this long (the reference takes any max_candidates; callers pass huge values for "no limit") lives in a
        // stream-ordered global scratch: same kernel, its two arrays just point there
        void *scratch = nullptr;
        const size_t bytes = (size_t)n_slots * kp * (sizeof(double) + sizeof(long long));
        FT8B_CUDA_TRY(cudaMallocAsync(&scratch, bytes, stream));
        a.big_s = reinterpret_cast<double *>(scratch);
        a.big_i = reinterpret_cast<long long *>(a.big_s + (size_t)n_slots * kp);
        select_candidates_kernel<1024, true><<<(unsigned)n_slots, 1024, 0, stream>>>(a);
        const cudaError_t le = cudaGetLastError();
        FT8B_CUDA_TRY(cudaFreeAsync(scratch, stream));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (le != cudaSuccess) return cuda_fail(le, "select_candidates_kernel");
        return 0;
    }
    const size_t smem = (size_t)kp * (sizeof(double) + sizeof(long long));
    if (n_tiles < 0) {       // strict top-K (opt-in extension)
        a.tiles = nullptr; a.n_tiles = 0;
        FT8B_CUDA_TRY(cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem > 48 * 1024 ? (int)smem : 48 * 1024));
        select_topk_kernel<<<(unsigned)n_slots, kSelThreads, smem, stream>>>(a);
        FT8B_LAUNCH_CHECK("select_topk_kernel");
        return 0;
    }
    if (K <= 64) {
        select_candidates_kernel<256><<<(unsigned)n_slots, 256, smem, stream>>>(a);
    } else {
        FT8B_CUDA_TRY(cudaFuncSetAttribute(select_candidates_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem > 48 * 1024 ? (int)smem : 48 * 1024));
        select_candidates_kernel<1024><<<(unsigned)n_slots, 1024, smem, stream>>>(a);
    }
    FT8B_LAUNCH_CHECK("select_candidates_kernel");
    return 0;
}

}  // namespace ft8b
