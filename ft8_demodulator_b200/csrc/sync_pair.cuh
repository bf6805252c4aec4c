// sync_pair.cuh - S2a for freq_osr B in {1, 2, 4}: the Costas score walk of sync.cu with TWO candidate
// frequencies per thread, fa and fb = fa + B (one tone apart).
//
// Why pairs.  A candidate at f reads the 8 tone bins f + tau*B; its neighbour one tone up reads 7 of the same bins.
// Walking both in one thread needs 9 bins per column instead of 16, and the frequency differences, time
// differences and second differences D2[t] = -(A[t-1]-A[t]) + (A[t]-A[t+1]) that both candidates use are formed
// once: 31 float64 adds per (column, candidate) instead of 41, 4.5 shared-memory loads instead of 8, and all the
// per-step overhead (cp.async staging, waits, score-slot predicates) is paid once per pair.  Every score is still
// the same sequence of IEEE additions as in sync_scores_kernel, so the two kernels agree bit for bit.
//
// Layout.  A warp owns 64 consecutive candidate frequencies; a lane walks fa = f_warp + PairGeom::lane_base(lane) and
// fb = fa + B.  CTA = 2 warps = 128 frequencies x one time-oversampling phase x one slot (the score sums,
// [nu][128] float64 in shared memory, cap the occupancy at 4 CTAs per SM either way).  Each warp streams its
// (64 + 7B)-bin strip through a private 6-column cp.async ring (two columns per commit group, four in flight) and
// prefetches the tone bins of column g+2 into registers while it does the arithmetic of step g.  Inside a ring column
// bin i sits at (i % 2B) * PITCH + i / 2B, with PITCH chosen so that both the staging writes (16 consecutive bins per
// half-warp) and the tone reads (16 lanes, fixed tone offset) touch 16 different banks.
// The kernel is bound by the shared-memory/L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 89 % of peak, FP64
// pipe 45 %), so wavefronts are what is counted here: per warp and column 18 for the tone reads, 5 for staging
// (12 before the bank fix and before lanes stopped zero-filling positions beyond the strip), ~7 for the score sums
// (the m = 0 block is stored, not added, and the sums are not zero-filled first: ~10 before).
#pragma once

namespace ft8b {

constexpr int kPairThreads = 64;
constexpr int kPairCols = 128;       // candidate frequencies per CTA
constexpr int kPairStages = 6;       // ring columns per warp, staged two at a time: 2 being read + 4 in flight
                                     // (8 would push a CTA past the 4-per-SM shared-memory budget)

template <int B>
struct PairGeom {
    static constexpr int kStrip = 64 + 7 * B;                               // bins a warp touches per column
    // plane pitch: >= 96 / 2B so that the 96 staging positions of a column are distinct.  Modulo 16 it is 8 (B = 1) or
    // 12 (B = 2): 16 consecutive bins - one 128-byte line of the spectrogram, staged by one half-warp - then land in 16
    // different 8-byte banks, and so do the tone reads of a half-warp (lane_base() below).  ncu: 2.0 instead of 4.0
    // wavefronts per cp.async on both sides.  B = 4 keeps conflict-free reads only.
    static constexpr int kPitch = (B == 1) ? 56 : (B == 2) ? 28 : 20;
    static constexpr int kCol = 2 * B * kPitch;                             // doubles per ring column
    __host__ __device__ static constexpr int pos(int i) { return (i % (2 * B)) * kPitch + i / (2 * B); }
    // offset of tone bin tau relative to the lane's base position pos((l / B) * 2B + l % B)
    __host__ __device__ static constexpr int tone(int tau) { return B * (tau & 1) * kPitch + tau / 2; }
    // fa - f_warp of a lane.  Candidates fa = 2B j + c (c < B) sit at c * kPitch + j: a half-warp takes one c
    // (B = 2: lanes 0-15 walk fa = 4 l, lanes 16-31 fa = 4 (l - 16) + 1) so that its 16 reads hit 16 banks.
    __host__ __device__ static constexpr int lane_base(int lane) {
        return (B == 2) ? (lane & 15) * 4 + (lane >> 4) : (lane / B) * 2 * B + lane % B;
    }
};

template <int B>
size_t sync_pair_smem_bytes(int nu) {
    return sizeof(double) * ((size_t)nu * kPairCols + (size_t)(kPairThreads / 32) * kPairStages * PairGeom<B>::kCol) +
           sizeof(double2) * (size_t)nu;
}

// One symbol step for the pair.  cur/nxt: the 9 tone bins fa + tau*B of columns g and g+1.  Returns the finished
// block sums S7(g-6) of fa and fb.
template <bool HAS_NEXT>
__device__ __forceinline__ void sync_pair_step(const double (&cur)[9], const double (&nxt)[9], double (&dtprev)[9],
                                               double (&acca)[6], double (&accb)[6], double &done_a, double &done_b) {
    double dhi[8], dt[9], d2[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) dhi[t] = __dsub_rn(cur[t], cur[t + 1]);
    if (HAS_NEXT) {
#pragma unroll
        for (int t = 0; t < 9; ++t) dt[t] = __dsub_rn(cur[t], nxt[t]);
    }
#pragma unroll
    for (int t = 1; t < 8; ++t) d2[t] = __dadd_rn(-dhi[t - 1], dhi[t]);      // (p - lo) + (p - hi) at bin t
    double ga[7], gb[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int tau = (k == 0) ? 3 : (k == 1) ? 1 : (k == 2) ? 4 : (k == 3) ? 0 : (k == 4) ? 6 : (k == 5) ? 5 : 2;
        double va = (tau > 0) ? d2[tau] : dhi[0];          // Costas tone 0 has no lower neighbour (ft8_decode.py:79-82)
        double vb = (tau > 0) ? d2[tau + 1] : dhi[1];
        if (k > 0) { va = __dsub_rn(va, dtprev[tau]); vb = __dsub_rn(vb, dtprev[tau + 1]); }      // (p - earlier)
        if (k < 6 && HAS_NEXT) { va = __dadd_rn(va, dt[tau]); vb = __dadd_rn(vb, dt[tau + 1]); }  // (p - later)
        ga[k] = va;
        gb[k] = vb;
    }
    if (HAS_NEXT) {
#pragma unroll
        for (int t = 0; t < 9; ++t) dtprev[t] = dt[t];
    }
    done_a = __dadd_rn(acca[5], ga[6]);
    done_b = __dadd_rn(accb[5], gb[6]);
#pragma unroll
    for (int k = 5; k >= 1; --k) {
        acca[k] = __dadd_rn(acca[k - 1], ga[k]);
        accb[k] = __dadd_rn(accb[k - 1], gb[k]);
    }
    acca[0] = ga[0];
    accb[0] = gb[0];
}

// add the finished block sums S7(h = g-6) of step g into score[u] for u = h+10, h-26, h-62 (m = 0, 1, 2);
// sp = &score[u = g+4] of fa, fb sits 32 doubles further.  The m = 0 block is the first contribution a score gets
// (u >= 4; rows 0..3 have no m = 0 block and start from zero), so it is stored, not added: no load, no zero fill.
__device__ __forceinline__ void sync_pair_scatter(double *sp, int g, int nu, double da, double db) {
    if ((unsigned)(g + 4) < (unsigned)nu) {
        sp[0] = da;
        sp[32] = db;
    }
#pragma unroll
    for (int m = 1; m < 3; ++m) {
        if ((unsigned)(g + 4 - m * kSyncOffset) < (unsigned)nu) {
            double *q = sp - m * kSyncOffset * kPairCols;
            q[0] = __dadd_rn(q[0], da);
            q[32] = __dadd_rn(q[32], db);
        }
    }
}

// TRACK: every warp also leaves the arg-max of its 64-frequency strip (score, flat scan index; first occurrence
// in scan order on ties) in a.tiles[slot][rho][warp], so that the selection does not have to re-read the score map
// to find the running record of ft8_decode.py:134-137 (select_candidates_kernel, phase B).
template <int B, bool TRACK>
__global__ void __launch_bounds__(kPairThreads) sync_scores_pair_kernel(const SyncArgs a) {
    using G = PairGeom<B>;
    extern __shared__ double sc[];   // [nu][128] score sums | ring[2 warps][kPairStages][kCol] | double2 ncmp[nu]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rho = blockIdx.y;
    const int64_t slot = blockIdx.z;
    const int nb = a.nb, nu = a.nu, s = a.s;
    double *const ring = sc + (size_t)nu * kPairCols + (size_t)wid * kPairStages * G::kCol;
    double *const ring_end = ring + kPairStages * G::kCol;
    double2 *ncmp = reinterpret_cast<double2 *>(sc + (size_t)nu * kPairCols + (size_t)(kPairThreads / 32) * kPairStages * G::kCol);
    for (int u = tid; u < nu; u += kPairThreads) {
        const int n = sync_num_comparisons(u, nb);
        ncmp[u] = make_double2((double)n, n > 0 ? __drcp_rn((double)n) : 0.0);
    }
    // score sums of this thread: fa in column wid*64 + lane, fb 32 columns further (both conflict-free)
    double *scp = sc + wid * 64 + lane;
    for (int u = 0; u < nu && u < 4; ++u) { scp[u * kPairCols] = 0.0; scp[u * kPairCols + 32] = 0.0; }   // see sync_pair_scatter
    __syncthreads();
    const int f_warp = (blockIdx.x * (kPairThreads / 32) + wid) * 64;
    SyncTile *const tile = TRACK ? a.tiles + (slot * s + rho) * (int64_t)(gridDim.x * (kPairThreads / 32)) +
                                       blockIdx.x * (kPairThreads / 32) + wid
                                 : nullptr;
    if (f_warp >= a.nf) {
        if (TRACK && lane == 0) { tile->score = -INFINITY; tile->idx = 0x7fffffffffffffffLL; }
        return;
    }
    const int base = G::lane_base(lane);                 // fa - f_warp
    const int fa = f_warp + base, fb = fa + B;
    const int last = nb - 1;

    // ---- staging: lane copies strip bins lane, lane+32, lane+64 of a column into their swizzled ring positions.
    // No bounds handling inside the loop: a bin beyond the row (only candidates >= nf, which are never stored, read
    // it) is replaced by the strip's first bin, and columns beyond the last one (only ever loaded as the unused
    // successor of the last step) by the last column - every copy has a valid source, so each is ONE cp.async (a
    // run-time source size compiles to two predicated LDGSTS variants per copy), and the third copy is predicated
    // off for the lanes the strip has no bin for.
    const double *col0 = a.mag + slot * (int64_t)a.T * a.F + (int64_t)rho * a.F + f_warp;
    const int64_t rowstep = (int64_t)s * a.F;
    const int f_room = a.F - f_warp;
    const int has2 = lane + 64 < G::kStrip;              // the strip's bins 64..kStrip-1: the low lanes only
    const int b0 = lane < f_room ? lane : 0, b1 = lane + 32 < f_room ? lane + 32 : 0,
              b2 = (has2 && lane + 64 < f_room) ? lane + 64 : 0;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned d0 = ring_s + 8u * G::pos(lane), d1 = ring_s + 8u * G::pos(lane + 32),
                   d2o = ring_s + 8u * G::pos(has2 ? lane + 64 : lane);
    const double *src = col0;                            // first bin of the strip in the next column to stage
    unsigned wr = 0;                                     // byte offset of the ring slot of the next column to stage
    int staged = 0;                                      // next column to stage
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
        wr += 8u * G::kCol;
        if (wr == 8u * kPairStages * G::kCol) wr = 0;
    };
    auto stage_two = [&]() { stage_col(); stage_col(); __pipeline_commit(); };

    // ---- registers: three column buffers rotate (cur, nxt, the one being prefetched)
    double ra[9], rb[9], rc[9], dtprev[9], acca[6], accb[6];
#pragma unroll
    for (int k = 0; k < 9; ++k) dtprev[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { acca[k] = 0.0; accb[k] = 0.0; }
    const double *rd = ring + G::pos(base);              // this lane's tone-0 position in the ring slot to read next
    const double *const rd_end = rd + kPairStages * G::kCol;
#define FT8B_PAIR_LOAD(BUF)                                                                  \
    do {                                                                                     \
        _Pragma("unroll") for (int tau = 0; tau < 9; ++tau) BUF[tau] = rd[G::tone(tau)];     \
        rd += G::kCol;                                                                       \
        if (rd == rd_end) rd -= kPairStages * G::kCol;                                       \
    } while (0)
    // prologue: columns 0..5 in flight as three groups, columns 0 and 1 landed and in registers
    stage_two(); stage_two(); stage_two();
    __pipeline_wait_prior(kPairStages / 2 - 1);
    __syncwarp();
    FT8B_PAIR_LOAD(ra);
    FT8B_PAIR_LOAD(rb);

    // A trip of two steps (columns g, g+1): stage columns g+6, g+7 into the slots of columns g, g+1 (read during the
    // previous trip), wait until columns g+2, g+3 have landed (two younger groups may be in flight), then each step
    // prefetches the column after its successor while doing its arithmetic.
#define FT8B_PAIR_STEP(CUR, NXT, PRE, O)                                                     \
    do {                                                                                     \
        FT8B_PAIR_LOAD(PRE);                                                                 \
        double da_, db_;                                                                     \
        sync_pair_step<true>(CUR, NXT, dtprev, acca, accb, da_, db_);                        \
        sync_pair_scatter(sp0 + (O) * kPairCols, g + (O), nu, da_, db_);                     \
    } while (0)
#define FT8B_PAIR_TRIP(X, Y, Z, O)                                                           \
    do {                                                                                     \
        __syncwarp();                                                                        \
        stage_two();                                                                         \
        __pipeline_wait_prior(kPairStages / 2 - 1);                                                          \
        __syncwarp();                                                                        \
        FT8B_PAIR_STEP(X, Y, Z, O);                                                          \
        FT8B_PAIR_STEP(Y, Z, X, (O) + 1);                                                    \
    } while (0)
    int g = 0;
    for (; g + 6 <= last; g += 6) {                      // six steps: the three register buffers are back in place
        double *const sp0 = scp + (g + 4) * kPairCols;
        FT8B_PAIR_TRIP(ra, rb, rc, 0);
        FT8B_PAIR_TRIP(rc, ra, rb, 2);
        FT8B_PAIR_TRIP(rb, rc, ra, 4);
    }
#undef FT8B_PAIR_TRIP
#undef FT8B_PAIR_STEP
    // at most six columns are left and all of them are already in the ring or on their way
    __pipeline_wait_prior(0);
    __syncwarp();
    for (; g <= last; ++g) {
        double da, db;
        if (g < last) {
            FT8B_PAIR_LOAD(rc);                          // column g+2 (unused garbage when there is none)
            sync_pair_step<true>(ra, rb, dtprev, acca, accb, da, db);
        } else {
            sync_pair_step<false>(ra, rb, dtprev, acca, accb, da, db);
        }
        sync_pair_scatter(scp + (g + 4) * kPairCols, g, nu, da, db);
#pragma unroll
        for (int tau = 0; tau < 9; ++tau) { ra[tau] = rb[tau]; rb[tau] = rc[tau]; }
    }
#undef FT8B_PAIR_LOAD
    // partial blocks that run past the last symbol: h = nb-6 .. nb-1 live in acc[5] .. acc[0]
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        sync_scatter(scp, kPairCols, nb - 1 - k, nu, acca[k]);
        sync_scatter(scp + 32, kPairCols, nb - 1 - k, nu, accb[k]);
    }
    // normalise and store: score / num_comparisons, -inf when no comparison or non-finite (ft8_decode.py:97-100)
    // (a candidate beyond nf - 1 has nothing to store: its pointer is parked on the warp's first column, which lane 0
    // owns, and the predicate keeps it from writing)
    const bool va = fa < a.nf, vb = fb < a.nf;
    double *oa = a.scores + slot * (int64_t)(nu * s) * a.nf + (int64_t)rho * a.nf + (va ? fa : f_warp);
    double *ob = oa + (vb ? fb : f_warp) - (va ? fa : f_warp);
    const int64_t outstep = (int64_t)s * a.nf;
    double best = -INFINITY;
    int best_u2 = 0x7fffffff;        // 2u + (0 for fa, 1 for fb) of the first maximum: ascending = scan order within the thread
#pragma unroll 4
    for (int u = 0; u < nu; ++u) {
        const double2 nr = ncmp[u];
        const double ya = sync_normalise(scp[u * kPairCols], nr), yb = sync_normalise(scp[u * kPairCols + 32], nr);
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
