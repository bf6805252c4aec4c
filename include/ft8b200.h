/*
 * ft8b200.h - C ABI of libft8b200.so, the B200 (sm_100a) FT8 demodulation hot path.
 *
 * The reference (Rintazero/ft8_demodulator) is pure Python and has no FFI; its "operator
 * interface" for this path is the set of Python functions in src/ft8_tools/ft8_demodulator/.
 * Each entry point below replaces the arithmetic of the reference function(s) it cites; the
 * Python mirror in ft8_demodulator_b200/ft8_demodulator/ binds them with ctypes and keeps the
 * reference's names, argument meaning and error behaviour.
 *
 * Conventions
 *   - return 0 = OK, negative = argument / capability error (FT8B_ERR_*), positive = cudaError_t;
 *     ft8b_last_error() returns a thread-local description of the last failure.
 *   - entry points that take a `stream`: every data pointer is a DEVICE pointer owned by the
 *     caller (contiguous, 16-byte aligned); the library allocates nothing per call - except a
 *     stream-ordered scratch (cudaMallocAsync / cudaFreeAsync on `stream`) on two rare paths: frames
 *     of more than 14 336 transform points (sample rates from 96 kHz up) and candidate lists above
 *     4096 entries.
 *     Per-(device, nfft) window/twiddle tables are created on first use (mutex-guarded,
 *     synchronous) and live until ft8b_release().  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream); every call is asynchronous on it.
 *   - ft8b_decode_slots_host is the one entry point on HOST pointers (audio in, results out):
 *     the library owns streams and device buffers there, and the call is synchronous.
 *   - arithmetic is IEEE float64 throughout ("f64"): the reference computes in float64 for
 *     float64 audio, and bit-exact candidate/bit/payload parity needs it.  float32 audio is
 *     widened on load (exact).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Layouts (row-major, last index fastest)
 *   wave    [n_slots][slot_stride]            float32 / float64 / int16 samples (or complex64 / complex128
 *                                             baseband), n_samples used per slot
 *   mag     [n_slots][T][F]                   dB spectrogram, time-major (the reference's
 *                                             FT8Waterfall.mag is [F][T]; the Python mirror transposes)
 *   scores  [n_slots][nt][nf]                 Costas sync scores in the reference's scan order
 *   cand_*  [n_slots][K]                      candidates, score-descending; n_cand[n_slots]
 *   llr     [n_words][174], plain [n_words][174] (one byte per bit), payload [n_words][10]
 */
#ifndef FT8B200_H
#define FT8B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FT8B_VERSION 100

#define FT8B_ERR_ARG (-1)          /* bad argument */
#define FT8B_ERR_UNSUPPORTED (-2)  /* size outside what the kernels cover (see message) */
#define FT8B_ERR_NO_DEVICE (-3)    /* no CUDA device / driver */

#define FT8B_DTYPE_F32 0
#define FT8B_DTYPE_F64 1
#define FT8B_DTYPE_I16 2   /* int16 PCM, normalised on load as float32(x) / 32767.0f (from_wave.py:62-66) */
#define FT8B_DTYPE_C64 3   /* complex64 (interleaved re,im float32): ft8b_stft_db (all nfft bins, natural FFT order) and the
                              decode entry points (bins 0 .. (nfft+1)/2 - 1 = the reference's f >= 0 rows, ft8_decode.py:323-325) */
#define FT8B_DTYPE_C128 4  /* complex128 */

#define FT8B_LDPC_N 174
#define FT8B_LDPC_M 83
#define FT8B_FLAG_STRICT_TOPK 1    /* ft8b_decode_params.flags: candidate list = the K best scores (ft8b_select_topk) */
#define FT8B_MAX_CANDIDATES 1048576   /* K limit of the selection (lists above 4096 entries live in a global scratch);
                                         ft8b_select_topk / FT8B_FLAG_STRICT_TOPK: 4096 */

int ft8b_version(void);
const char *ft8b_last_error(void);
/* Free the cached per-device tables (optional; for tests / orderly shutdown). */
int ft8b_release(void);

/* Window/FFT geometry of calculate_spectrogram (spectrogram_analyse.py:32-43):
 * nperseg = int(0.16*fs), hop = nperseg - noverlap, nfft = int(fs/6.25*bins_per_tone). Host only. */
int ft8b_plan(double sample_rate, int bins_per_tone, int steps_per_symbol, int *nperseg, int *hop, int *nfft);
/* Frames scipy produces: (n - nperseg)/hop + 1, or 0 when n < nperseg (spectrogram_analyse.py:37-39). */
int64_t ft8b_num_frames(int64_t n_samples, int nperseg, int hop);
/* Candidate search ranges of ft8_find_candidates (ft8_decode.py:108-109):
 * abs_time in [t0, t0+nt), abs_freq in [0, nf).  nt or nf <= 0 means "no candidates". */
int ft8b_search_ranges(int T, int F, int time_osr, int freq_osr, int *t0, int *nt, int *nf);

/*
 * S1  calculate_spectrogram (spectrogram_analyse.py:19-66 -> scipy.signal.spectrogram, window='hann',
 *     detrend=False, return_onesided=False, scaling='spectrum') fused with the 10*log10(1e-12+|S|) of
 *     :59-60 and with the row/column selection of decode_ft8_message (ft8_decode.py:323-341).
 *     Real input: writes bins [bin_lo, bin_hi) of 0..floor(nfft/2) and frames [frame_lo, frame_hi) as
 *     mag[slot][frame-frame_lo][bin-bin_lo]; the two-sided fftshifted array the reference returns is the
 *     mirror image of these rows (the Python mirror builds it).  Complex input (FT8B_DTYPE_C64/C128, the
 *     satellite pre-processor's case): bins are 0..nfft-1 in natural FFT order; the Python mirror fftshifts.
 */
int ft8b_stft_db(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride,
                 int nperseg, int hop, int nfft, int bin_lo, int bin_hi, int64_t frame_lo, int64_t frame_hi,
                 double *mag, void *stream);

/*
 * S2a ft8_sync_score for every (abs_time, abs_freq) of ft8_find_candidates' double loop
 *     (ft8_decode.py:47-100, 108-124).  scores[slot][abs_time - t0][abs_freq]; -inf where the
 *     reference returns -inf.
 */
int ft8b_sync_scores(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr, double *scores,
                     void *stream);

/*
 * S2b Candidate selection with the reference's exact heap semantics (ft8_decode.py:127-140):
 *     the first K threshold passers in scan order, the running best replaced by every later
 *     strictly greater score, result sorted by descending score.  K <= FT8B_MAX_CANDIDATES.
 *     Exact score ties: where the reference returns a list, tied candidates come out in its heap-array order (CPython's
 *     heapq replayed on the list; the reference sorts the heap with a stable key sort, :140); where its heapq compares
 *     two equal scores (the reference raises TypeError) they are ordered by scan position.
 *     n_pass (optional, may be NULL) receives the number of threshold passers per slot.
 */
int ft8b_select_candidates(const double *scores, int64_t n_slots, int nt, int nf, int t0, int K, double min_score,
                           int32_t *cand_time, int32_t *cand_freq, double *cand_score, int32_t *n_cand,
                           int32_t *n_pass, void *stream);
/*
 * ft8_find_candidates (ft8_decode.py:102-149) on a score map made by ft8b_sync_scores from `mag` (same T, F, osr): the
 * selection above, with the spectrogram at hand.  The sync kernels add per cell, per Costas block, per score - the same
 * terms as the reference's single running sum in a different association, equal to ~1e-13 relative but not bit for
 * bit.  Here every score a decision or a result hangs on is re-evaluated exactly as ft8_sync_score does it
 * (ft8_decode.py:57-100: one running sum in (m, k, lower/higher/earlier/later) order, one IEEE division): the returned
 * cand_score, the order of the list, the threshold test of any score within 1e-9 (relative to 1 + |min_score|) of
 * min_score, and the record replacement.  For the same spectrogram the list, its order and its scores then equal the
 * reference's bit for bit, unless two DIFFERENT map cells compete for the map's maximum within 1e-13.
 * ft8b_decode_slots selects this way.
 */
int ft8b_find_candidates(const double *mag, const double *scores, int64_t n_slots, int T, int F, int time_osr, int freq_osr,
                         int K, double min_score, int32_t *cand_time, int32_t *cand_freq, double *cand_score,
                         int32_t *n_cand, int32_t *n_pass, void *stream);
/*
 * ft8_sync_score (ft8_decode.py:47-100) for given candidates, evaluated exactly as the reference does (one running sum
 * in (m, k, lower/higher/earlier/later) order, one IEEE division; -inf when nothing is compared or the sum is not
 * finite): score[slot][i] for (cand_time, cand_freq)[slot][i], i < n_per_slot.  Any abs_time; 0 <= abs_freq and
 * abs_freq + 7*freq_osr < F is the caller's to guarantee (the reference indexes out of range there), violating
 * candidates get NaN.  Bit-equal to the reference on the same spectrogram.
 */
int ft8b_sync_score_at(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr,
                       const int32_t *cand_time, const int32_t *cand_freq, int n_per_slot, double *score, void *stream);
/*
 * OPT-IN EXTENSION (not the reference's behaviour; SURVEY Appendix D "may add an opt-in strict top-K"): the K best
 * threshold passers by (score descending, scan position ascending), in that order - what a reader of
 * ft8_find_candidates would expect it to return.  Same outputs as ft8b_select_candidates.  In the pipeline it is
 * chosen with FT8B_FLAG_STRICT_TOPK in ft8b_decode_params.flags; the default stays the reference's list.
 */
int ft8b_select_topk(const double *scores, int64_t n_slots, int nt, int nf, int t0, int K, double min_score,
                     int32_t *cand_time, int32_t *cand_freq, double *cand_score, int32_t *n_cand, void *stream);

/*
 * S3  ft8_decode_candidate (ft8_decode.py:225-273) for every candidate of every slot:
 *     ft8_extract_likelihood (:164-188) + ftx_normalize_logl (:190-198) + bp_decode
 *     (ldpc_decoder.py:54-113, sum-product with the rational tanh/atanh of :11-31) + pack_bits (:200-215)
 *     + CRC-14 (crc.py:11-54).  One warp per candidate.  Entries c >= n_cand[slot] are zero-filled.
 *     Outputs are [n_slots][K](x...); llr_out and plain_out may be NULL.
 */
int ft8b_decode_candidates(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr,
                           const int32_t *cand_time, const int32_t *cand_freq, const int32_t *n_cand, int K,
                           int max_iterations, uint8_t *ok, int32_t *ldpc_errors, int32_t *crc_extracted,
                           int32_t *crc_calculated, uint8_t *payload, double *llr_out, uint8_t *plain_out,
                           void *stream);

/*
 * S3' bp_decode + CRC check on caller-supplied, already normalised LLR vectors
 *     (ldpc_decoder.py:54-113; ft8_decode.py:236-271).  plain/min_errors are what bp_decode returns;
 *     the CRC fields follow ft8_decode_candidate (zero unless min_errors == 0).  plain, ldpc_errors and
 *     ok are required; the other outputs may be NULL.
 */
int ft8b_ldpc_crc(const double *llr, int64_t n_words, int max_iterations, uint8_t *plain, int32_t *ldpc_errors,
                  int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload, uint8_t *ok, void *stream);

/*
 * ft8_extract_likelihood (ft8_decode.py:164-188), optionally followed by ftx_normalize_logl (:190-198),
 * for every candidate: llr[n_slots][K][174].  scratch_u8 / scratch_i32 are [n_slots][K] work arrays.
 */
int ft8b_extract_llr(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr,
                     const int32_t *cand_time, const int32_t *cand_freq, const int32_t *n_cand, int K, int normalize,
                     double *llr, uint8_t *scratch_u8, int32_t *scratch_i32, void *stream);
/* ftx_normalize_logl (ft8_decode.py:190-198) in place on llr[n_words][174]. */
int ft8b_normalize_llr(double *llr, int64_t n_words, void *stream);
/* ldpc_check (ldpc_decoder.py:33-52): errors[w] = unsatisfied parity checks of bits[w][174] (one byte per bit). */
int ft8b_ldpc_check(const uint8_t *bits, int64_t n_words, int32_t *errors, void *stream);
/*
 * Self-check of the branch-free float64 division the BP kernel uses for fast_tanh / fast_atanh
 * (ldpc_decoder.py:11-31) against IEEE division: out2[0] = number of i with a differing quotient a[i]/b[i],
 * out2[1] = number of i examined (operands the kernel would route to its exact fallback are skipped).
 * The kernel relies on out2[0] == 0 for b in [60, 2.1e4]; tests/test_gpu_parity.py sweeps that domain.
 */
int ft8b_selfcheck_divide(const double *a, const double *b, int64_t n, int64_t *out2, void *stream);

/*
 * Satellite pre-processor correct_frequency_drift (ft8_beacon_receiver/frequency_correction.py:118-659), its two
 * data-parallel steps; the regression / template-correlation logic around them is scalar host work.
 *   ft8b_column_argmax: idx[c] = argmax over rows [f_lo, f_hi) of spectrogram column c (first maximum, relative to
 *                       f_lo), mag laid out [n_cols][F] as ft8b_stft_db writes it            (:222-224, :382-384)
 *   ft8b_dechirp:       out[n] = in[n] * exp(-2*pi*i*(c2*n^2 + c3*n^3)), in complex64/complex128 (FT8B_DTYPE_C64/
 *                       C128), out complex128                                                 (:352-355, :598-634)
 */
int ft8b_column_argmax(const double *mag, int64_t n_cols, int F, int f_lo, int f_hi, int32_t *idx, void *stream);
int ft8b_dechirp(const void *in, int in_dtype, int64_t n, double c2, double c3, void *out, void *stream);

/*
 * Drift-hypothesis Costas search for the satellite path (EXTENSION named in BASELINE north_star; the reference has no
 * counterpart, results are checked at payload level only).  drift_bins_per_symbol is a HOST array of n_hyp (<= 128)
 * hypotheses: the tone grid of symbol q of a candidate is shifted by round(d*q) bins.  best_score/best_hyp
 * [n_slots][nt][nf]: maximum of ft8_sync_score (ft8_decode.py:47-100, evaluated on the shifted cells) over the
 * hypotheses and its index (first maximum; list d = 0 first).  ft8b_select_candidates then runs on best_score, and
 * ft8b_decode_candidates_drift extracts LLRs along each candidate's drift (cand_drift [n_slots][K], bins/symbol).
 * The hypothesis list is staged in constant memory in stream order: calls carrying DIFFERENT lists must not be in
 * flight on different streams of one device at the same time (every other entry point is free of such state).
 * Three kernels compute the same map (same best_hyp; scores within 1e-12 of the reference's running sum):
 *   FT8B_DRIFT_DIRECT    one thread per (time, frequency), every hypothesis evaluated from global memory with the
 *                        reference's running sum; any shift; the cross-check of the other two
 *   FT8B_DRIFT_WALK      the 21 hypothesis-independent cell contrasts of a 128-frequency strip are formed once in
 *                        shared memory, a hypothesis only chooses where to read them (|shift| <= 64 bins)
 *   FT8B_DRIFT_BLOCKSUM  additionally sums the seven contrasts of each Costas block once per distinct within-block
 *                        shift pattern, so a hypothesis reads 3 block sums instead of 21 contrasts
 * ft8b_sync_scores_drift picks BLOCKSUM when the hypothesis list has few enough patterns to fit and pay, else WALK,
 * else DIRECT; ft8b_sync_scores_drift_variant forces one (FT8B_DRIFT_AUTO = the same choice; BLOCKSUM fails with
 * FT8B_ERR_ARG when the list does not fit) and exists for the tests and the A/B bench.
 */
#define FT8B_DRIFT_AUTO 0
#define FT8B_DRIFT_WALK 1
#define FT8B_DRIFT_BLOCKSUM 2
#define FT8B_DRIFT_DIRECT 3
int ft8b_sync_scores_drift(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr, int n_hyp,
                           const double *drift_bins_per_symbol, double *best_score, int32_t *best_hyp, void *stream);
int ft8b_sync_scores_drift_variant(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr, int n_hyp,
                                   const double *drift_bins_per_symbol, double *best_score, int32_t *best_hyp, int variant,
                                   void *stream);
int ft8b_decode_candidates_drift(const double *mag, int64_t n_slots, int T, int F, int time_osr, int freq_osr,
                                 const int32_t *cand_time, const int32_t *cand_freq, const int32_t *n_cand,
                                 const double *cand_drift, int K, int max_iterations, uint8_t *ok, int32_t *ldpc_errors,
                                 int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload, void *stream);

/*
 * Whole path, decode_ft8_message (ft8_decode.py:288-394) for a batch of equal-length slots:
 * S1 -> S2a -> S2b -> LLR in chunks of `chunk_slots` slots (the float64 spectrogram and score map exist for one
 * chunk at a time: 4.2 MB per slot at 12 kHz, osr 2x2), then one LDPC/CRC launch over the candidates of the whole
 * batch.  `workspace` must hold ft8b_decode_workspace_bytes(p, n_slots, n_samples) bytes; that size is an upper bound
 * for every batch of AT MOST n_slots slots decoded with the same parameters (a workspace sized once for the largest
 * batch serves all shorter ones, e.g. a ragged last piece).
 * Outputs as in ft8b_select_candidates / ft8b_decode_candidates ([n_slots][K]...); every output array is fully
 * defined on return (unused candidate positions and "nothing to search" batches hold zeros).  crc_extracted and
 * crc_calculated may be NULL when not wanted (both entry points); every other pointer is required.
 */
typedef struct {
    double sample_rate;
    int32_t bins_per_tone, steps_per_symbol;
    int32_t max_candidates;
    int32_t max_iterations;
    double min_score;
    int32_t bin_lo, bin_hi;       /* rows kept (freq_min/freq_max mask), within [0, ceil(nfft/2)] */
    int64_t frame_lo, frame_hi;   /* columns kept (time_min/time_max mask) */
    int32_t chunk_slots;          /* 0 = library default */
    int32_t flags;                /* 0 = the reference's behaviour; FT8B_FLAG_* bits opt in to extensions */
} ft8b_decode_params;

int64_t ft8b_decode_workspace_bytes(const ft8b_decode_params *p, int64_t n_slots, int64_t n_samples);
int ft8b_decode_slots(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride,
                      const ft8b_decode_params *p, void *workspace, int64_t workspace_bytes, int32_t *cand_time,
                      int32_t *cand_freq, double *cand_score, int32_t *n_cand, uint8_t *ok, int32_t *ldpc_errors,
                      int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload, void *stream);

/*
 * The same path from HOST memory to HOST memory - the call a NumPy-only caller such as the reference makes (no CUDA
 * array library needed; this is what `e2e` in bench.py times).  wave and all nine result arrays are host pointers;
 * the library owns two streams, two device audio buffers, the chunk workspace and the device result arrays on
 * `device` (-1 = the calling thread's current device; cached between calls, grow-only, freed by ft8b_release).  The batch runs in pieces of `piece_slots`
 * slots (0 = 256): the host-to-device copy of piece i+1 overlaps the kernels of piece i.  Page-locked caller memory is
 * copied from directly; pageable memory is gathered into pinned staging buffers by persistent host threads first
 * (streaming stores; ft8b_set_host_staging).
 * The nine result arrays come back as ONE device-to-host copy into a pinned mirror and are then copied out on the host.
 * Synchronous: the result arrays are complete when the call returns.  Calls on the same device are serialised inside
 * the library (they share the cached buffers); calls on different devices run concurrently (one host thread per GPU).
 * On failure nothing of the call is left in flight: both streams are drained before the call returns, and after a
 * CUDA error the cached state of that device is dropped and rebuilt by the next call.
 */
int ft8b_decode_slots_host(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride,
                           const ft8b_decode_params *p, int device, int64_t piece_slots, int32_t *cand_time, int32_t *cand_freq,
                           double *cand_score, int32_t *n_cand, uint8_t *ok, int32_t *ldpc_errors,
                           int32_t *crc_extracted, int32_t *crc_calculated, uint8_t *payload);
/* Device-side duration (CUDA events: first copy enqueued .. last result byte on the host) of the calling thread's
 * most recent successful ft8b_decode_slots_host call, in ms. */
double ft8b_last_host_call_ms(void);
/* Host threads that gather pageable caller memory into the pinned staging buffers of ft8b_decode_slots_host
 * (persistent, created on first use, the calling thread included in the count).  threads <= 0: the default
 * (environment FT8B_STAGE_THREADS, else hardware threads - 2, at most 8); non_temporal != 0: write the staging buffers with
 * streaming stores (default; environment FT8B_STAGE_NT).  Inputs below FT8B_STAGE_MIN_KB (default 1024 KB) are
 * gathered by the calling thread alone.  Returns the thread count in effect. */
int ft8b_set_host_staging(int threads, int non_temporal);
/* Test hook (host only, no device needed): the staging gather itself - rows x row_bytes from a source of pitch
 * src_pitch bytes into a dense destination, by the staging threads.  Returns 0. */
int ft8b_selftest_host_staging(void *dst, const void *src, int64_t rows, int64_t row_bytes, int64_t src_pitch);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t ft8b_launch_count(void);
/* Per-stage device timing of ft8b_decode_slots (cudaEvent pairs on the launching stream).
 * ft8b_stage_timing(1) resets and enables, (0) disables.  ft8b_stage_times fills ms[5] / launches[5]
 * for the stages {stft, sync, select, llr, ldpc} accumulated since the reset (synchronises). */
int ft8b_stage_timing(int enable);
int ft8b_stage_times(double *ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* FT8B200_H */
