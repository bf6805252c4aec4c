// api.cu - extern "C" surface of libft8b200.so (see include/ft8b200.h) and the chunked pipeline.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ft8b {

std::atomic<int64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// implemented in stft.cu / sync.cu / decode.cu
int stft_db(const void *, int, int64_t, int64_t, int64_t, int, int, int, int, int, int64_t, int64_t, double *, cudaStream_t);
int release_plans();
int sync_tiles_per_slot(int, int, int, int);
int sync_scores(const double *, int64_t, int, int, int, int, double *, SyncTile *, cudaStream_t);
int sync_score_at(const double *, int64_t, int, int, int, int, const int32_t *, const int32_t *, int, double *, cudaStream_t);
int select_candidates(const double *, int64_t, int, int, int, int, double, int32_t *, int32_t *, double *, int32_t *,
                      int32_t *, const SyncTile *, int, const double *, int, int, int, int, cudaStream_t);
int decode_candidates(const double *, int64_t, int, int, int, int, const int32_t *, const int32_t *, const int32_t *, int,
                      int, uint8_t *, int32_t *, int32_t *, int32_t *, uint8_t *, double *, uint8_t *, const double *,
                      cudaStream_t);
int sync_scores_drift(const double *, int64_t, int, int, int, int, int, const double *, double *, int32_t *, int, cudaStream_t);
int ldpc_crc(const double *, int64_t, int, uint8_t *, int32_t *, int32_t *, int32_t *, uint8_t *, uint8_t *, cudaStream_t);
int extract_llr(const double *, int64_t, int, int, int, int, const int32_t *, const int32_t *, const int32_t *, int, int,
                double *, uint8_t *, int32_t *, cudaStream_t);
int normalize_llr(double *, int64_t, cudaStream_t);
int ldpc_crc_masked(const double *, const int32_t *, int, int64_t, int, uint8_t *, int32_t *, int32_t *, int32_t *, uint8_t *,
                    cudaStream_t);
int ldpc_check(const uint8_t *, int64_t, int32_t *, cudaStream_t);
int selfcheck_divide(const double *, const double *, int64_t, int64_t *, cudaStream_t);
int column_argmax(const double *, int64_t, int, int, int, int32_t *, cudaStream_t);
int release_host_paths();
int dechirp(const void *, int, int64_t, double, double, double2 *, cudaStream_t);

// ---- optional per-stage timing (cudaEvent pairs on the launching stream), used by bench.py ----
constexpr int kStages = 5;   // stft, sync, select, llr, ldpc
struct StageTimer {
    std::mutex mu;
    bool enabled = false;
    std::vector<cudaEvent_t> beg[kStages], end[kStages];
    size_t used[kStages] = {0, 0, 0, 0, 0};
    double ms[kStages] = {0, 0, 0, 0, 0};
    int64_t launches[kStages] = {0, 0, 0, 0, 0};
};
static StageTimer g_timer;

struct StageScope {
    int stage;
    cudaStream_t stream;
    cudaEvent_t e_end = nullptr;
    StageScope(int st, cudaStream_t s) : stage(st), stream(s) {
        if (!g_timer.enabled) return;
        std::lock_guard<std::mutex> lk(g_timer.mu);
        size_t &u = g_timer.used[stage];
        if (u == g_timer.beg[stage].size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            g_timer.beg[stage].push_back(a);
            g_timer.end[stage].push_back(b);
        }
        cudaEventRecord(g_timer.beg[stage][u], stream);
        e_end = g_timer.end[stage][u];
        ++u;
    }
    ~StageScope() {
        if (e_end) cudaEventRecord(e_end, stream);
    }
};

static int default_chunk_slots(int64_t per_slot_bytes) {
    // The kernels are bound by the FP64 pipe, not HBM, so L2 residency of a chunk is not the point; larger chunks
    // mean fewer launches and shorter tails (measured on 4096 slots: 188k / 206k / 214k slots/s at 128 / 487 / 2048).
    // Cap the chunk workspace at 8 GiB of the 180 GB.
    int64_t c = (8192ll << 20) / (per_slot_bytes > 0 ? per_slot_bytes : 1);
    if (c < 1) c = 1;
    if (c > 2048) c = 2048;
    return (int)c;
}

struct Geometry {
    int nperseg, hop, nfft, F, t0, nt, nf;
    int64_t T_all, T;
    int64_t mag_elems, score_elems;   // per slot
};

static int geometry(const ft8b_decode_params *p, int64_t n_samples, Geometry *g) {
    ft8b_plan(p->sample_rate, p->bins_per_tone, p->steps_per_symbol, &g->nperseg, &g->hop, &g->nfft);
    if (g->nperseg < 1 || g->hop < 1 || g->nfft < g->nperseg) FT8B_ARG_FAIL("decode: bad sample_rate/oversampling");
    g->T_all = ft8b_num_frames(n_samples, g->nperseg, g->hop);
    if (p->bin_lo < 0 || p->bin_hi < p->bin_lo || p->bin_hi > (g->nfft + 1) / 2) FT8B_ARG_FAIL("decode: bad bin range");
    if (p->frame_lo < 0 || p->frame_hi < p->frame_lo || p->frame_hi > g->T_all) FT8B_ARG_FAIL("decode: bad frame range");
    g->F = p->bin_hi - p->bin_lo;
    g->T = p->frame_hi - p->frame_lo;
    if (g->T > 0x7fffffff) FT8B_ARG_FAIL("decode: too many frames");
    ft8b_search_ranges((int)g->T, g->F, p->steps_per_symbol, p->bins_per_tone, &g->t0, &g->nt, &g->nf);
    g->mag_elems = g->T * (int64_t)g->F;
    g->score_elems = (g->nt > 0 && g->nf > 0) ? (int64_t)g->nt * g->nf : 0;
    return 0;
}

}  // namespace ft8b

using namespace ft8b;

extern "C" {

int ft8b_version(void) { return FT8B_VERSION; }
const char *ft8b_last_error(void) { return g_err; }
int ft8b_release(void) {
    release_host_paths();
    return release_plans();
}
int64_t ft8b_launch_count(void) { return g_launches.load(); }

int ft8b_stage_timing(int enable) {
    std::lock_guard<std::mutex> lk(g_timer.mu);
    g_timer.enabled = enable != 0;
    for (int s = 0; s < kStages; ++s) { g_timer.used[s] = 0; g_timer.ms[s] = 0; g_timer.launches[s] = 0; }
    return 0;
}

int ft8b_stage_times(double *ms, int64_t *launches) {
    if (!ms || !launches) FT8B_ARG_FAIL("ft8b_stage_times: null pointer");
    std::lock_guard<std::mutex> lk(g_timer.mu);
    for (int s = 0; s < kStages; ++s) {
        for (size_t i = 0; i < g_timer.used[s]; ++i) {
            FT8B_CUDA_TRY(cudaEventSynchronize(g_timer.end[s][i]));
            float t = 0.f;
            FT8B_CUDA_TRY(cudaEventElapsedTime(&t, g_timer.beg[s][i], g_timer.end[s][i]));
            g_timer.ms[s] += t;
            g_timer.launches[s] += 1;
        }
        g_timer.used[s] = 0;
        ms[s] = g_timer.ms[s];
        launches[s] = g_timer.launches[s];
    }
    return 0;
}

int ft8b_plan(double fs, int bins_per_tone, int steps_per_symbol, int *nperseg, int *hop, int *nfft) {
    if (!nperseg || !hop || !nfft || bins_per_tone < 1 || steps_per_symbol < 1 || !(fs > 0)) FT8B_ARG_FAIL("ft8b_plan: bad argument");
    const int sps = (int)(0.16 * fs);                      // spectrogram_analyse.py:32
    int overlap = sps - sps / steps_per_symbol;            // :33
    const int n = (int)(fs / 6.25 * bins_per_tone);        // :34
    if (overlap >= sps) overlap = sps - 1;                 // :42-43
    *nperseg = sps;
    *hop = sps - overlap;
    *nfft = n;
    return 0;
}

int64_t ft8b_num_frames(int64_t n, int nperseg, int hop) {
    if (nperseg < 1 || hop < 1 || n < nperseg) return 0;   // :37-39
    return (n - nperseg) / hop + 1;
}

int ft8b_search_ranges(int T, int F, int s, int b, int *t0, int *nt, int *nf) {
    if (!t0 || !nt || !nf || s < 1 || b < 1) FT8B_ARG_FAIL("ft8b_search_ranges: bad argument");
    const int nb = T / s;                                  // ftx_types.py:31-34
    *t0 = -10 * s;                                         // ft8_decode.py:108
    *nt = nb * s - s * (kND + 1) - *t0;
    *nf = F - 7 * b;                                       // :109
    return 0;
}

int ft8b_stft_db(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride, int nperseg,
                 int hop, int nfft, int bin_lo, int bin_hi, int64_t frame_lo, int64_t frame_hi, double *mag, void *stream) {
    return stft_db(wave, wave_dtype, n_slots, n_samples, slot_stride, nperseg, hop, nfft, bin_lo, bin_hi, frame_lo,
                   frame_hi, mag, (cudaStream_t)stream);
}

int ft8b_sync_scores(const double *mag, int64_t n_slots, int T, int F, int s, int b, double *scores, void *stream) {
    return sync_scores(mag, n_slots, T, F, s, b, scores, nullptr, (cudaStream_t)stream);
}

int ft8b_select_candidates(const double *scores, int64_t n_slots, int nt, int nf, int t0, int K, double min_score,
                           int32_t *cand_time, int32_t *cand_freq, double *cand_score, int32_t *n_cand, int32_t *n_pass,
                           void *stream) {
    return select_candidates(scores, n_slots, nt, nf, t0, K, min_score, cand_time, cand_freq, cand_score, n_cand, n_pass,
                             nullptr, 0, nullptr, 0, 0, 0, 0, (cudaStream_t)stream);
}

int ft8b_find_candidates(const double *mag, const double *scores, int64_t n_slots, int T, int F, int s, int b, int K,
                         double min_score, int32_t *cand_time, int32_t *cand_freq, double *cand_score, int32_t *n_cand,
                         int32_t *n_pass, void *stream) {
    if (!mag) FT8B_ARG_FAIL("ft8b_find_candidates: null spectrogram");
    if (T < 0 || F < 0 || s < 1 || b < 1) FT8B_ARG_FAIL("ft8b_find_candidates: bad shape");
    int t0, nt, nf;
    ft8b_search_ranges(T, F, s, b, &t0, &nt, &nf);
    return select_candidates(scores, n_slots, nt, nf, t0, K, min_score, cand_time, cand_freq, cand_score, n_cand, n_pass,
                             nullptr, 0, mag, T, F, s, b, (cudaStream_t)stream);
}

int ft8b_sync_score_at(const double *mag, int64_t n_slots, int T, int F, int s, int b, const int32_t *cand_time,
                       const int32_t *cand_freq, int n_per_slot, double *score, void *stream) {
    return sync_score_at(mag, n_slots, T, F, s, b, cand_time, cand_freq, n_per_slot, score, (cudaStream_t)stream);
}

int ft8b_select_topk(const double *scores, int64_t n_slots, int nt, int nf, int t0, int K, double min_score,
                     int32_t *cand_time, int32_t *cand_freq, double *cand_score, int32_t *n_cand, void *stream) {
    // n_tiles = -1 selects the strict top-K kernel
    return select_candidates(scores, n_slots, nt, nf, t0, K, min_score, cand_time, cand_freq, cand_score, n_cand, nullptr,
                             nullptr, -1, nullptr, 0, 0, 0, 0, (cudaStream_t)stream);
}

int ft8b_decode_candidates(const double *mag, int64_t n_slots, int T, int F, int s, int b, const int32_t *cand_time,
                           const int32_t *cand_freq, const int32_t *n_cand, int K, int max_iterations, uint8_t *ok,
                           int32_t *ldpc_errors, int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload,
                           double *llr_out, uint8_t *plain_out, void *stream) {
    return decode_candidates(mag, n_slots, T, F, s, b, cand_time, cand_freq, n_cand, K, max_iterations, ok, ldpc_errors,
                             crc_extracted, crc_calculated, payload, llr_out, plain_out, nullptr, (cudaStream_t)stream);
}

int ft8b_decode_candidates_drift(const double *mag, int64_t n_slots, int T, int F, int s, int b, const int32_t *cand_time,
                                 const int32_t *cand_freq, const int32_t *n_cand, const double *cand_drift, int K,
                                 int max_iterations, uint8_t *ok, int32_t *ldpc_errors, int32_t *crc_extracted,
                                 int32_t *crc_calculated, uint8_t *payload, void *stream) {
    if (!cand_drift) FT8B_ARG_FAIL("ft8b_decode_candidates_drift: null cand_drift");
    return decode_candidates(mag, n_slots, T, F, s, b, cand_time, cand_freq, n_cand, K, max_iterations, ok, ldpc_errors,
                             crc_extracted, crc_calculated, payload, nullptr, nullptr, cand_drift, (cudaStream_t)stream);
}

int ft8b_sync_scores_drift(const double *mag, int64_t n_slots, int T, int F, int s, int b, int n_hyp,
                           const double *drift_bins_per_symbol, double *best_score, int32_t *best_hyp, void *stream) {
    return sync_scores_drift(mag, n_slots, T, F, s, b, n_hyp, drift_bins_per_symbol, best_score, best_hyp, 0,
                             (cudaStream_t)stream);
}

int ft8b_sync_scores_drift_variant(const double *mag, int64_t n_slots, int T, int F, int s, int b, int n_hyp,
                                   const double *drift_bins_per_symbol, double *best_score, int32_t *best_hyp, int variant,
                                   void *stream) {
    if (variant < 0 || variant > 3) FT8B_ARG_FAIL("ft8b_sync_scores_drift_variant: variant %d", variant);
    return sync_scores_drift(mag, n_slots, T, F, s, b, n_hyp, drift_bins_per_symbol, best_score, best_hyp, variant,
                             (cudaStream_t)stream);
}

int ft8b_ldpc_crc(const double *llr, int64_t n_words, int max_iterations, uint8_t *plain, int32_t *ldpc_errors,
                  int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload, uint8_t *ok, void *stream) {
    return ldpc_crc(llr, n_words, max_iterations, plain, ldpc_errors, crc_extracted, crc_calculated, payload, ok,
                    (cudaStream_t)stream);
}

int ft8b_extract_llr(const double *mag, int64_t n_slots, int T, int F, int s, int b, const int32_t *cand_time,
                     const int32_t *cand_freq, const int32_t *n_cand, int K, int normalize, double *llr, uint8_t *scratch_u8,
                     int32_t *scratch_i32, void *stream) {
    return extract_llr(mag, n_slots, T, F, s, b, cand_time, cand_freq, n_cand, K, normalize, llr, scratch_u8, scratch_i32,
                       (cudaStream_t)stream);
}

int ft8b_normalize_llr(double *llr, int64_t n_words, void *stream) { return normalize_llr(llr, n_words, (cudaStream_t)stream); }

int ft8b_ldpc_check(const uint8_t *bits, int64_t n_words, int32_t *errors, void *stream) {
    return ldpc_check(bits, n_words, errors, (cudaStream_t)stream);
}

int ft8b_selfcheck_divide(const double *a, const double *b, int64_t n, int64_t *out2, void *stream) {
    return selfcheck_divide(a, b, n, out2, (cudaStream_t)stream);
}

int ft8b_column_argmax(const double *mag, int64_t n_cols, int F, int f_lo, int f_hi, int32_t *idx, void *stream) {
    return column_argmax(mag, n_cols, F, f_lo, f_hi, idx, (cudaStream_t)stream);
}

int ft8b_dechirp(const void *in, int in_dtype, int64_t n, double c2, double c3, void *out, void *stream) {
    return dechirp(in, in_dtype, n, c2, c3, reinterpret_cast<double2 *>(out), (cudaStream_t)stream);
}

// Chunking of a batch: per-slot workspace (spectrogram + score map + strip summaries of the sync kernel), and chunks of
// equal size (4096 slots at a 2045-slot limit run as 3 x 1366, not 2045 + 2045 + 6).
struct ChunkPlan {
    int64_t per_slot;   // bytes
    int chunk, n_tiles;
    int64_t limit;      // min(chunk limit, n_slots): no batch of <= n_slots slots runs a larger chunk
};
static ChunkPlan chunk_plan(const ft8b_decode_params *p, const Geometry &g, int64_t n_slots) {
    ChunkPlan c;
    c.n_tiles = sync_tiles_per_slot((int)g.T, g.F, p->steps_per_symbol, p->bins_per_tone);
    c.per_slot = 8 * (g.mag_elems + g.score_elems) + (int64_t)sizeof(SyncTile) * c.n_tiles;
    const int64_t n = n_slots > 0 ? n_slots : 1;
    int64_t lim = p->chunk_slots > 0 ? p->chunk_slots : default_chunk_slots(c.per_slot);
    if (lim > n) lim = n;
    c.limit = lim;
    const int64_t n_chunks = (n + lim - 1) / lim;
    c.chunk = (int)((n + n_chunks - 1) / n_chunks);
    return c;
}

int64_t ft8b_decode_workspace_bytes(const ft8b_decode_params *p, int64_t n_slots, int64_t n_samples) {
    if (!p) return FT8B_ERR_ARG;
    Geometry g;
    if (geometry(p, n_samples, &g)) return FT8B_ERR_ARG;
    const ChunkPlan c = chunk_plan(p, g, n_slots);
    // Chunk spectrogram + score map + strip summaries, then the LLR vectors of every candidate of the batch.
    // The chunk term is sized for the LARGEST chunk any batch of at most n_slots slots can be cut into
    // (min(limit, n_slots) slots), not for the balanced chunk of exactly n_slots: a workspace sized for n_slots
    // therefore also serves every shorter batch decoded with the same parameters (the tail piece of
    // ft8b_decode_slots_host: 128-slot pieces with chunk_slots = 100 run as 2 x 64, a 72..100-slot tail as one chunk).
    const int64_t n = n_slots > 0 ? n_slots : 0;
    return c.per_slot * (c.limit < 1 ? 1 : c.limit) + 8ll * kLdpcN * p->max_candidates * n + 256;
}

int ft8b_decode_slots(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride,
                      const ft8b_decode_params *p, void *workspace, int64_t workspace_bytes, int32_t *cand_time,
                      int32_t *cand_freq, double *cand_score, int32_t *n_cand, uint8_t *ok, int32_t *ldpc_errors,
                      int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // crc_extracted / crc_calculated may be NULL (not wanted); every other array is required
    if (!p || !wave || !workspace || !cand_time || !cand_freq || !cand_score || !n_cand || !ok || !ldpc_errors || !payload)
        FT8B_ARG_FAIL("ft8b_decode_slots: null pointer");
    // complex audio (the reference takes it too: two-sided spectrogram, rows f >= 0 kept - ft8_decode.py:316-325) runs a
    // full-length complex transform in the STFT stage; everything after the spectrogram is the same
    if (wave_dtype != FT8B_DTYPE_F32 && wave_dtype != FT8B_DTYPE_F64 && wave_dtype != FT8B_DTYPE_I16 &&
        wave_dtype != FT8B_DTYPE_C64 && wave_dtype != FT8B_DTYPE_C128)
        FT8B_ARG_FAIL("ft8b_decode_slots: bad dtype");
    if (p->max_candidates < 0 || p->max_candidates > FT8B_MAX_CANDIDATES) {
        set_error("ft8b_decode_slots: max_candidates=%d outside [0,%d]", p->max_candidates, FT8B_MAX_CANDIDATES);
        return FT8B_ERR_UNSUPPORTED;
    }
    Geometry g;
    int rc = geometry(p, n_samples, &g);
    if (rc) return rc;
    if (n_slots <= 0) return 0;
    const int K = p->max_candidates;
    const size_t elt = wave_dtype == FT8B_DTYPE_F32 ? 4 : wave_dtype == FT8B_DTYPE_I16 ? 2 : wave_dtype == FT8B_DTYPE_C128 ? 16 : 8;
    if (g.T <= 0 || g.F <= 0 || g.score_elems == 0 || K == 0) {
        // nothing to search: the reference returns an empty list.  Every output array is defined (zeros), as after a
        // search that found nothing - a direct C caller may pass uninitialised buffers.
        const size_t nK = (size_t)n_slots * (size_t)K;
        FT8B_CUDA_TRY(cudaMemsetAsync(n_cand, 0, sizeof(int32_t) * n_slots, stream));
        if (nK > 0) {
            FT8B_CUDA_TRY(cudaMemsetAsync(cand_time, 0, sizeof(int32_t) * nK, stream));
            FT8B_CUDA_TRY(cudaMemsetAsync(cand_freq, 0, sizeof(int32_t) * nK, stream));
            FT8B_CUDA_TRY(cudaMemsetAsync(cand_score, 0, sizeof(double) * nK, stream));
            FT8B_CUDA_TRY(cudaMemsetAsync(ok, 0, nK, stream));
            FT8B_CUDA_TRY(cudaMemsetAsync(ldpc_errors, 0, sizeof(int32_t) * nK, stream));
            if (crc_extracted) FT8B_CUDA_TRY(cudaMemsetAsync(crc_extracted, 0, sizeof(int32_t) * nK, stream));
            if (crc_calculated) FT8B_CUDA_TRY(cudaMemsetAsync(crc_calculated, 0, sizeof(int32_t) * nK, stream));
            FT8B_CUDA_TRY(cudaMemsetAsync(payload, 0, 10 * nK, stream));
        }
        return 0;
    }
    const ChunkPlan cp = chunk_plan(p, g, n_slots);
    const int chunk = cp.chunk;
    const int64_t need = cp.per_slot * chunk + 8ll * kLdpcN * K * n_slots;
    if (workspace_bytes < need) FT8B_ARG_FAIL("ft8b_decode_slots: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    double *mag = reinterpret_cast<double *>(workspace);
    double *scores = mag + g.mag_elems * chunk;
    SyncTile *tiles = reinterpret_cast<SyncTile *>(scores + g.score_elems * chunk);   // [chunk][n_tiles], when the sync kernel makes them
    double *llr = reinterpret_cast<double *>(tiles + (int64_t)cp.n_tiles * chunk);     // [n_slots][K][174]
    if (cp.n_tiles == 0) tiles = nullptr;
    for (int64_t s0 = 0; s0 < n_slots; s0 += chunk) {
        const int64_t nb = (n_slots - s0 < chunk) ? (n_slots - s0) : chunk;
        const char *w = reinterpret_cast<const char *>(wave) + (size_t)s0 * slot_stride * elt;
        {
            StageScope t(0, stream);
            rc = stft_db(w, wave_dtype, nb, n_samples, slot_stride, g.nperseg, g.hop, g.nfft, p->bin_lo, p->bin_hi,
                         p->frame_lo, p->frame_hi, mag, stream);
        }
        if (rc) return rc;
        {
            StageScope t(1, stream);
            rc = sync_scores(mag, nb, (int)g.T, g.F, p->steps_per_symbol, p->bins_per_tone, scores, tiles, stream);
        }
        if (rc) return rc;
        {
            StageScope t(2, stream);
            const bool strict = (p->flags & FT8B_FLAG_STRICT_TOPK) != 0;       // opt-in extension, see ft8b_select_topk
            // the spectrogram goes along: scores that decide or are returned are re-evaluated in the reference's own
            // summation order (select_candidates_kernel, ref_sync_score)
            rc = select_candidates(scores, nb, g.nt, g.nf, g.t0, K, p->min_score, cand_time + s0 * K, cand_freq + s0 * K,
                                   cand_score + s0 * K, n_cand + s0, nullptr, strict ? nullptr : tiles, strict ? -1 : cp.n_tiles,
                                   strict ? nullptr : mag, (int)g.T, g.F, p->steps_per_symbol, p->bins_per_tone, stream);
        }
        if (rc) return rc;
        {
            // LLRs of this chunk's candidates before its spectrogram is overwritten; ok/ldpc_errors serve as scratch
            StageScope t(3, stream);
            rc = extract_llr(mag, nb, (int)g.T, g.F, p->steps_per_symbol, p->bins_per_tone, cand_time + s0 * K,
                             cand_freq + s0 * K, n_cand + s0, K, 1, llr + s0 * K * kLdpcN, ok + s0 * K, ldpc_errors + s0 * K,
                             stream);
        }
        if (rc) return rc;
    }
    {
        // one launch over every candidate of the batch: warps stay busy although iteration counts differ
        StageScope t(4, stream);
        rc = ldpc_crc_masked(llr, n_cand, K, n_slots * (int64_t)K, p->max_iterations, ok, ldpc_errors, crc_extracted,
                             crc_calculated, payload, stream);
    }
    if (rc) return rc;
    return 0;
}

}  // extern "C"
