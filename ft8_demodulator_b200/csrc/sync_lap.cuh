// sync_lap.cuh - S2a for freq_osr 2 (the canonical geometry): the Costas score map as a LAPLACIAN OF COSTAS PATH SUMS.
// NOT the default kernel (FT8B_SYNC_LAP=1 selects it): measured 5.79 ms against the 5.19 ms of the difference walk per
// 4096 slots - see "Why / outcome" below.  It stays as an independent formulation the walk is cross-checked against.
//
// ft8_sync_score (ft8_decode.py:47-100) adds, for each of the 21 Costas cells of a candidate, up to four differences
// p - neighbour (one tone lower / higher, one symbol earlier / later).  Per Costas block h (7 cells, symbol columns
// h..h+6) that is
//     B[h](f) = 4 P[h](f) - P[h](f+b) - P[h](f-b) - P[h-1](f) - P[h+1](f) + Corr[h](f)
// where P[h](f) = sum_k A[h+k][f + b*C[k]] is the plain sum of the block's 7 cells (C = Costas pattern, A = this
// oversampling phase of the waterfall, columns outside [0, nb) contributing nothing): the lower / higher tone
// neighbours of the cells of candidate f are the cells of candidates f-b / f+b, and the earlier / later neighbours
// are the cells of blocks h-1 / h+1.  Corr[h] collects the comparisons the reference does NOT make and the cells the
// shifted path sums do not share:
//     + A[h+3][f-b] - A[h+3][f]          tone 0 (k = 3) has no lower neighbour
//     + A[h-1][c0] - A[h][c0]            k = 0 has no earlier neighbour  (c0 = f + b*C[0])
//     + A[h+7][c6] - A[h+6][c6]          k = 6 has no later neighbour    (c6 = f + b*C[6])
//     + A[0][c(j-1)] - A[0][c(j)]        block cut by the first column (h = -j, j = 1..6): its cell in column 0
//                                        has no earlier neighbour, and P[h+1] has one cell more in column 0
//     + A[nb-1][c(k+1)] - A[nb-1][c(k)]  block cut by the last column (h = nb-1-k, k = 0..5), likewise
// (each term only where its column exists).  A score is the sum of its three blocks divided by the comparison count.
// The CPU oracle states this in NumPy (score_map_laplacian) and tests/test_oracle_golden.py holds it to the
// reference-pinned running-sum map (2e-14 on dB-scale waterfalls).
//
// Why.  The walk of sync_pair.cuh forms the four differences per cell: 31.6 float64 adds per (column, candidate).
// Here a (column, candidate) costs 6 adds for P, 4.5 for Corr, 6.5 for B and 2.7 for the three scores a block feeds:
// under 20 - and every hot kernel of this library is bound by the FP64 pipe.  The scores differ from the reference's
// running sum in the last bits (as the walk's did); ft8b_find_candidates re-evaluates every score that decides or is
// returned in the reference's own order, so the candidate lists are unaffected.
// Outcome: the float64 additions fell by 40 % (58 -> 36 per thread and column; total instructions 112 -> 103) and the
// kernel got SLOWER: the neighbour exchange adds 4 shuffles per step to the 22 shared-memory operations the walk
// already issues, the halo costs 6.7 % more warps and the block flush 7.5 % more steps.  The Costas search is bound
// by the shared-memory / shuffle (MIO) path and by issue latency at two warps per scheduler, not by the FP64 pipe.
// Limit: cells that cancel algebraically (the k = 3 cell of P(f-b), the extra cells of P[h-1] / P[h+1]) still poison a
// sum when they are inf / NaN, so on a waterfall with non-finite cells a few more scores come out -inf than the
// reference voids (finite waterfalls: identical pattern).
//
// Layout.  Thread = two candidates one tone apart (fa, fb = fa + 2), as in sync_pair.cuh, so each is the other's
// f+b / f-b neighbour; the outer neighbours (P of fa-2, P of fb+2) live in the adjacent lanes and arrive by shuffle
// one step after they complete.  A 16-lane half-warp walks the 32 candidates of one parity class of the warp's window;
// the first and the last of them are a halo (their outer neighbour is in another warp) and are not stored: a warp
// stores 60 of its 64 candidates, warp w of a phase covers abs_freq [60 w, 60 w + 60).  CTA = 2 warps x one phase x
// one slot; score sums [nu][128] in shared memory, a private 6-column cp.async ring per warp (78 bins per column:
// the window, one tone below it and six above), the 9 tone bins of the next column prefetched into registers
// during the arithmetic of the current one.  The E chain (Corr) is a 7-deep register delay line: the main loop is
// unrolled over 8 steps so that it rotates by renaming, without moves.
#pragma once

namespace ft8b {

constexpr int kLapThreads = 64;
constexpr int kLapStore = 60;                        // candidates a warp stores
constexpr int kLapCols = 128;                        // score-sum columns per CTA (halo slots included)
constexpr int kLapStages = 6;                        // ring columns per warp
constexpr int kLapStrip = 78;                        // bins per column a warp touches: window 64 + one tone below (2) + tones 1..6 above (12)
constexpr int kLapPitch = 28;                        // plane pitch of the swizzle, as PairGeom<2>
constexpr int kLapCol = 4 * kLapPitch;               // doubles per ring column

// strip bin i sits at (i % 4) * pitch + i / 4: staging writes and tone reads of a half-warp both hit 16 banks
__host__ __device__ constexpr int lap_pos(int i) { return (i % 4) * kLapPitch + i / 4; }
// offset of tone slot tau (bin fa - 2 + 2 tau, tau = 0..8) from the lane's base position lap_pos(4 l + c)
__host__ __device__ constexpr int lap_tone(int tau) { return 2 * (tau & 1) * kLapPitch + tau / 2; }

inline size_t sync_lap_smem_bytes(int nu) {
    return sizeof(double) * ((size_t)nu * kLapCols + (size_t)(kLapThreads / 32) * kLapStages * kLapCol) + sizeof(double2) * (size_t)nu;
}
inline int sync_lap_warps(int nf) { return (nf + kLapStore - 1) / kLapStore; }

struct LapState {
    double pa[6], pb[6];      // path-sum accumulators: p[k] = P of block g-k through cell k (after step g)
    double ea[7], eb[7];      // Corr accumulators:     e[j] = Corr of block g-j so far
    double pa1, pb1;          // P of the block completed one step ago (block g-7 at step g)
    double dpa, dpb;          // P[g-7] - P[g-8]
    double dn_a, up_b;        // P[g-7] of fa - 2 (lane - 1's fb) and of fb + 2 (lane + 1's fa)
};

enum { kLapNormal = 0, kLapFirst = 1, kLapLast = 2, kLapVirtual = 3 };

// One symbol step.  cur: column g's tone slots; d3, d4, d5: column g minus column g-1 at tone slots 3, 4, 5 (the c6 / c0
// bins of the pair; the caller forms them before the buffer of column g-1 is overwritten by the prefetch of g+1).
// MODE kLapFirst / kLapLast: column 0 / column nb-1 (blocks cut by the edge get their extra terms); kLapVirtual: one of
// the seven columns g >= nb that only flush the blocks still in flight (cur is not read: every cell is zero).
// Returns B[g-7] of both candidates.
template <int MODE>
__device__ __forceinline__ void lap_step(const double (&cur)[9], double d3, double d4, double d5, LapState &st, double &Ba,
                                         double &Bb) {
    double pa_new, pb_new, ea_done, eb_done;
    if (MODE != kLapVirtual) {
        // cells of the pair: tone t of fa is slot t + 1, of fb slot t + 2; Costas C = {3, 1, 4, 0, 6, 5, 2}
        const double Aa[7] = {cur[4], cur[2], cur[5], cur[1], cur[7], cur[6], cur[3]};
        const double Ab[7] = {cur[5], cur[3], cur[6], cur[2], cur[8], cur[7], cur[4]};
        // path sums: P[g-6] completes
        pa_new = st.pa[5] + Aa[6];
        pb_new = st.pb[5] + Ab[6];
#pragma unroll
        for (int k = 5; k >= 1; --k) { st.pa[k] = st.pa[k - 1] + Aa[k]; st.pb[k] = st.pb[k - 1] + Ab[k]; }
        st.pa[0] = Aa[0];
        st.pb[0] = Ab[0];
        // Corr: time differences at the c6 / c0 bins (slot 4 is fa's c0 and fb's c6), tone-0 frequency differences
        const double fda = cur[0] - cur[1], fdb = cur[1] - cur[2];
        ea_done = st.ea[6] + d3;                                             // block g-7: + A[g][c6] - A[g-1][c6]
        eb_done = st.eb[6] + d4;
        st.ea[6] = st.ea[5]; st.eb[6] = st.eb[5];
        st.ea[5] = st.ea[4]; st.eb[5] = st.eb[4];
        st.ea[4] = st.ea[3]; st.eb[4] = st.eb[3];
        st.ea[3] = st.ea[2] + fda; st.eb[3] = st.eb[2] + fdb;                // block g-3: tone 0 has no lower neighbour
        st.ea[2] = st.ea[1]; st.eb[2] = st.eb[1];
        st.ea[1] = st.ea[0]; st.eb[1] = st.eb[0];
        st.ea[0] = -d4; st.eb[0] = -d5;                                      // block g: A[g-1][c0] - A[g][c0]
        if (MODE == kLapFirst) {
            // blocks h = -j (j = 1..6) start before the first column: + A[0][c(j-1)] - A[0][c(j)]
#pragma unroll
            for (int j = 1; j <= 6; ++j) { st.ea[j] += Aa[j - 1] - Aa[j]; st.eb[j] += Ab[j - 1] - Ab[j]; }
        }
        if (MODE == kLapLast) {
            // blocks h = nb-1-j (j = 0..5) run past the last column: + A[nb-1][c(j+1)] - A[nb-1][c(j)]
#pragma unroll
            for (int j = 0; j <= 5; ++j) { st.ea[j] += Aa[j + 1] - Aa[j]; st.eb[j] += Ab[j + 1] - Ab[j]; }
        }
    } else {
        pa_new = st.pa[5];
        pb_new = st.pb[5];
#pragma unroll
        for (int k = 5; k >= 1; --k) { st.pa[k] = st.pa[k - 1]; st.pb[k] = st.pb[k - 1]; }
        st.pa[0] = 0.0;
        st.pb[0] = 0.0;
        ea_done = st.ea[6] + d3;
        eb_done = st.eb[6] + d4;
#pragma unroll
        for (int j = 6; j >= 1; --j) { st.ea[j] = st.ea[j - 1]; st.eb[j] = st.eb[j - 1]; }
        st.ea[0] = -d4; st.eb[0] = -d5;
    }
    // B[h], h = g-7, as differences of neighbouring path sums: (P - P(f-2)) + (P - P(f+2)) + (P - P[h-1]) - (P[h+1] - P)
    // + Corr.  Same value as 4 P - ..., but nothing of the size of 4 P is ever formed (a constant waterfall - digital
    // silence - scores exactly 0 wherever all three blocks and their time neighbours are complete, as in the reference),
    // fa and fb share the difference between them, and each time difference serves two consecutive blocks.
    const double dpa = pa_new - st.pa1, dpb = pb_new - st.pb1;           // P[h+1] - P[h]
    const double t = st.pa1 - st.pb1;                                    // fa's (P - P(f+2)) and minus fb's (P - P(f-2))
    Ba = (((st.pa1 - st.dn_a) + t) + (st.dpa - dpa)) + ea_done;
    Bb = (((st.pb1 - st.up_b) - t) + (st.dpb - dpb)) + eb_done;
    st.dpa = dpa; st.dpb = dpb;
    st.pa1 = pa_new; st.pb1 = pb_new;
    // the outer neighbours of block g-6, needed one step from now (halo lanes get their own value back: never stored)
    st.dn_a = __shfl_up_sync(0xffffffffu, pb_new, 1, 16);
    st.up_b = __shfl_down_sync(0xffffffffu, pa_new, 1, 16);
}

// add B[h = g-7] into score[u] for u = g+3, g-33, g-69 (m = 0, 1, 2).  sp = &score[u = g+3] of fa; fb sits 32 doubles on.
// The m = 0 block is the first contribution a score gets (rows 0..2 have none and start from zero): stored, not added.
__device__ __forceinline__ void lap_scatter(double *sp, int g, int nu, double ba, double bb) {
    if ((unsigned)(g + 3) < (unsigned)nu) {
        sp[0] = ba;
        sp[32] = bb;
    }
#pragma unroll
    for (int m = 1; m < 3; ++m) {
        if ((unsigned)(g + 3 - m * kSyncOffset) < (unsigned)nu) {
            double *q = sp - m * kSyncOffset * kLapCols;
            q[0] += ba;
            q[32] += bb;
        }
    }
}

// NB: number of symbol columns per phase as a compile-time constant (93 for a 15-s slot: head, tail and flush steps are
// then unrolled with their modes resolved at compile time), or 0 for any other length (rolled generic steps).
template <bool TRACK, int NB>
__global__ void __launch_bounds__(kLapThreads) sync_scores_lap_kernel(const SyncArgs a) {
    extern __shared__ double sc[];   // [nu][128] score sums | ring[2 warps][kLapStages][kLapCol] | double2 ncmp[nu]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rho = blockIdx.y;
    const int64_t slot = blockIdx.z;
    const int nb = NB > 0 ? NB : a.nb, nu = a.nu, s = a.s;
    double *const ring = sc + (size_t)nu * kLapCols + (size_t)wid * kLapStages * kLapCol;
    double2 *ncmp = reinterpret_cast<double2 *>(sc + (size_t)nu * kLapCols + (size_t)(kLapThreads / 32) * kLapStages * kLapCol);
    for (int u = tid; u < nu; u += kLapThreads) {
        const int n = sync_num_comparisons(u, nb);
        ncmp[u] = make_double2((double)n, n > 0 ? __drcp_rn((double)n) : 0.0);
    }
    double *scp = sc + wid * 64 + lane;                  // fa's column; fb's is 32 further
    for (int u = 0; u < nu && u < 3; ++u) { scp[u * kLapCols] = 0.0; scp[u * kLapCols + 32] = 0.0; }   // see lap_scatter
    __syncthreads();
    const int wglob = blockIdx.x * (kLapThreads / 32) + wid;
    const int f_store = wglob * kLapStore;               // first candidate this warp stores
    SyncTile *const tile = TRACK ? a.tiles + (slot * s + rho) * (int64_t)(gridDim.x * (kLapThreads / 32)) + wglob : nullptr;
    if (f_store >= a.nf) {
        if (TRACK && lane == 0) { tile->score = -INFINITY; tile->idx = 0x7fffffffffffffffLL; }
        return;
    }
    const int l = lane & 15, c = lane >> 4;
    const int fa = f_store - 2 + 4 * l + c, fb = fa + 2;
    const int strip0 = f_store - 4;                      // bin of strip index 0 (tone slot 0 of the first halo candidate)

    // ---- staging: lane copies strip bins lane, lane+32, lane+64 (< 78) of a column into their swizzled ring positions.
    // Bins outside the row are replaced by the nearest valid one: below bin 0 only the k = 3 cell of the halo candidates
    // f = -2, -1 and the tone-0 correction of f = 0, 1 read them, and those two cancel; above the row only never-stored
    // candidates do.  Columns beyond the last one (prefetched, never used) are replaced by the last column.
    const double *col0 = a.mag + slot * (int64_t)a.T * a.F + (int64_t)rho * a.F;
    const int64_t rowstep = (int64_t)s * a.F;
    const int has2 = lane + 64 < kLapStrip;
    auto clampb = [&](int i) { const int b = strip0 + i; return b < 0 ? 0 : (b >= a.F ? a.F - 1 : b); };
    const int b0 = clampb(lane), b1 = clampb(lane + 32), b2 = clampb(has2 ? lane + 64 : lane);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned d0 = ring_s + 8u * lap_pos(lane), d1 = ring_s + 8u * lap_pos(lane + 32),
                   d2o = ring_s + 8u * lap_pos(has2 ? lane + 64 : lane);
    const double *src = col0;
    unsigned wr = 0;
    int staged = 0;
    const int last = nb - 1;
    auto stage_col = [&]() {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %6, 0;\n\t"
            "cp.async.ca.shared.global [%0], [%3], 8;\n\t"
            "cp.async.ca.shared.global [%1], [%4], 8;\n\t"
            "@p cp.async.ca.shared.global [%2], [%5], 8;\n\t}"
            :: "r"(d0 + wr), "r"(d1 + wr), "r"(d2o + wr), "l"(src + b0), "l"(src + b1), "l"(src + b2), "r"(has2)
            : "memory");
        src += (staged < last) ? rowstep : 0;
        ++staged;
        wr += 8u * kLapCol;
        if (wr == 8u * kLapStages * kLapCol) wr = 0;
    };

    LapState st;
#pragma unroll
    for (int k = 0; k < 6; ++k) { st.pa[k] = 0.0; st.pb[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < 7; ++k) { st.ea[k] = 0.0; st.eb[k] = 0.0; }
    st.pa1 = st.pb1 = st.dpa = st.dpb = st.dn_a = st.up_b = 0.0;
    double ra[9], rb[9];
    const double *rd = ring + lap_pos(4 * l + c);        // this lane's tone slot 0 in the ring column to read next
    const double *const rd_end = rd + kLapStages * kLapCol;
#define FT8B_LAP_LOAD(BUF)                                                                   \
    do {                                                                                     \
        _Pragma("unroll") for (int tau = 0; tau < 9; ++tau) BUF[tau] = rd[lap_tone(tau)];   \
        rd += kLapCol;                                                                       \
        if (rd == rd_end) rd -= kLapStages * kLapCol;                                        \
    } while (0)
    // prologue: columns 0..3 in flight as four groups, column 0 landed and in registers
    stage_col(); __pipeline_commit();
    stage_col(); __pipeline_commit();
    stage_col(); __pipeline_commit();
    stage_col(); __pipeline_commit();
    __pipeline_wait_prior(3);
    __syncwarp();
    FT8B_LAP_LOAD(ra);
    int g = 0;
    // A step on buffers (X = column g, Y = column g-1, overwritten by the prefetch of column g+1 once the three time
    // differences have been taken)
#define FT8B_LAP_STEP(X, Y, MODE, SP, G, LOADNEXT)                                           \
    do {                                                                                     \
        const double d3_ = X[3] - Y[3], d4_ = X[4] - Y[4], d5_ = X[5] - Y[5];                \
        if (LOADNEXT) FT8B_LAP_LOAD(Y);                                                      \
        double ba_, bb_;                                                                     \
        lap_step<MODE>(X, d3_, d4_, d5_, st, ba_, bb_);                                      \
        lap_scatter(SP, G, nu, ba_, bb_);                                                    \
    } while (0)
#pragma unroll
    for (int tau = 0; tau < 9; ++tau) rb[tau] = 0.0;     // "column -1"
    // head: column 0 (one column staged, as in the generic steps)
    __syncwarp();
    stage_col(); __pipeline_commit();                    // column 4
    __pipeline_wait_prior(2);                            // column 1 has landed
    __syncwarp();
    FT8B_LAP_STEP(ra, rb, kLapFirst, scp + 3 * kLapCols, 0, true);
    g = 1;                                               // column 1 is in rb
    // main loop: steps g .. g+7 all strictly inside (0, nb-1); two columns staged per pair of steps
#define FT8B_LAP_TRIP(X, Y, O)                                                               \
    do {                                                                                     \
        __syncwarp();                                                                        \
        stage_col(); stage_col(); __pipeline_commit();   /* columns g+O+4, g+O+5 */          \
        FT8B_LAP_STEP(X, Y, kLapNormal, sp0 + (O) * kLapCols, g + (O), true);                \
        __pipeline_wait_prior(1);                                                            \
        __syncwarp();                                                                        \
        FT8B_LAP_STEP(Y, X, kLapNormal, sp0 + ((O) + 1) * kLapCols, g + (O) + 1, true);      \
    } while (0)
    for (; g + 8 <= nb - 1; g += 8) {
        double *const sp0 = scp + (g + 3) * kLapCols;
        FT8B_LAP_TRIP(rb, ra, 0);
        FT8B_LAP_TRIP(rb, ra, 2);
        FT8B_LAP_TRIP(rb, ra, 4);
        FT8B_LAP_TRIP(rb, ra, 6);
    }
#undef FT8B_LAP_TRIP
    // tail: the remaining columns, then seven virtual ones that complete the blocks still in flight.  Column g is in rb.
    if (NB > 0) {
        // compile-time length: every remaining real column (at most NB-1 <= g+4 with g = 8n+1) was staged by the main loop
        constexpr int kG0 = 1 + 8 * ((NB - 2) / 8);
        static_assert(NB == 0 || NB - 1 <= kG0 + 3, "tail columns must already be staged by the main loop");
        __pipeline_wait_prior(0);
        __syncwarp();
        double *const sp0 = scp + (kG0 + 3) * kLapCols;
#pragma unroll
        for (int i = 0; kG0 + i < NB + 7; ++i) {
            const int gg = kG0 + i;
            double *const sp = sp0 + i * kLapCols;
            if (gg < NB - 1) {
                if (i % 2 == 0) FT8B_LAP_STEP(rb, ra, kLapNormal, sp, gg, true); else FT8B_LAP_STEP(ra, rb, kLapNormal, sp, gg, true);
            } else if (gg == NB - 1) {
                if (i % 2 == 0) FT8B_LAP_STEP(rb, ra, kLapLast, sp, gg, false); else FT8B_LAP_STEP(ra, rb, kLapLast, sp, gg, false);
            } else {
                // virtual column: zero cells; the first one still sees the last real column as its predecessor
                const double(&prev)[9] = ((NB - 1 - kG0) % 2 == 0) ? rb : ra;
                const bool first_v = gg == NB;
                double ba_, bb_;
                lap_step<kLapVirtual>(ra, first_v ? -prev[3] : 0.0, first_v ? -prev[4] : 0.0, first_v ? -prev[5] : 0.0, st, ba_, bb_);
                lap_scatter(sp, gg, nu, ba_, bb_);
            }
        }
    } else {
        // any other length: one rolled step, mode chosen at run time; one column staged per step
        bool cur_is_rb = true;
        for (; g < nb + 7; ++g) {
            __syncwarp();
            stage_col(); __pipeline_commit();            // column g+4 into the slot of column g-2
            __pipeline_wait_prior(2);                    // column g+1 has landed
            __syncwarp();
            double *const sp = scp + (g + 3) * kLapCols;
            if (g < nb - 1) {
                if (cur_is_rb) FT8B_LAP_STEP(rb, ra, kLapNormal, sp, g, true); else FT8B_LAP_STEP(ra, rb, kLapNormal, sp, g, true);
            } else if (g == nb - 1) {
                if (cur_is_rb) FT8B_LAP_STEP(rb, ra, kLapLast, sp, g, false); else FT8B_LAP_STEP(ra, rb, kLapLast, sp, g, false);
            } else {
                // after the last real step the buffers are no longer swapped: the last column stays where it is
                const bool first_v = g == nb;
                const double p3 = cur_is_rb ? ra[3] : rb[3], p4 = cur_is_rb ? ra[4] : rb[4], p5 = cur_is_rb ? ra[5] : rb[5];
                double ba_, bb_;
                lap_step<kLapVirtual>(ra, first_v ? -p3 : 0.0, first_v ? -p4 : 0.0, first_v ? -p5 : 0.0, st, ba_, bb_);
                lap_scatter(sp, g, nu, ba_, bb_);
                continue;
            }
            cur_is_rb = !cur_is_rb;
        }
    }
    __pipeline_wait_prior(0);
#undef FT8B_LAP_STEP
#undef FT8B_LAP_LOAD

    // normalise and store: score / num_comparisons, -inf when no comparison or non-finite (ft8_decode.py:97-100)
    const bool va = (l != 0) && fa < a.nf, vb = (l != 15) && fb < a.nf;
    double *oa = a.scores + slot * (int64_t)(nu * s) * a.nf + (int64_t)rho * a.nf + (va ? fa : f_store);
    double *ob = oa + (vb ? fb : f_store) - (va ? fa : f_store);
    const int64_t outstep = (int64_t)s * a.nf;
    double best = -INFINITY;
    int best_u2 = 0x7fffffff;        // 2u + (0 for fa, 1 for fb) of the first maximum: ascending = scan order within the thread
#pragma unroll 4
    for (int u = 0; u < nu; ++u) {
        const double2 nr = ncmp[u];
        const double ya = sync_normalise(scp[u * kLapCols], nr), yb = sync_normalise(scp[u * kLapCols + 32], nr);
        if (va) *oa = ya;
        if (vb) *ob = yb;
        oa += outstep;
        ob += outstep;
        if (TRACK) {
            const bool ga = va & (ya > best);
            best = ga ? ya : best;
            best_u2 = ga ? 2 * u : best_u2;
            const bool gb = vb & (yb > best);
            best = gb ? yb : best;
            best_u2 = gb ? 2 * u + 1 : best_u2;
        }
    }
    if (TRACK) {
        long long bi = 0x7fffffffffffffffLL;
        if (best_u2 != 0x7fffffff)
            bi = ((long long)(best_u2 >> 1) * s + rho) * a.nf + ((best_u2 & 1) ? fb : fa);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, best, d);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (os > best || (os == best && oi < bi)) { best = os; bi = oi; }
        }
        if (lane == 0) { tile->score = best; tile->idx = bi; }
    }
}

}  // namespace ft8b
