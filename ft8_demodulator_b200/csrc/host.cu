// host.cu - ft8b_decode_slots_host: decode_ft8_message (ft8_decode.py:288-394) for a batch of slots that live in HOST
// memory, results back in HOST memory.  This is the call a caller without any CUDA array library makes (the reference
// is NumPy-only): plain pointers in, plain pointers out, the library owns streams and device memory.
//
// The batch is cut into pieces; the host-to-device copy of piece i+1 (copy stream) runs under the kernels of piece i
// (compute stream) through two device audio buffers.  Pinned caller memory (cudaHostAlloc / cudaHostRegister, e.g. a
// pinned torch tensor) is copied from directly; pageable memory (NumPy) is first gathered into two pinned staging
// buffers by a few host threads, which overlaps as well because every device call here is asynchronous.
// All device state is cached per device and reused by later calls (grow-only); ft8b_release() frees it.
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define FT8B_HAVE_SSE2 1
#endif
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ft8b {

namespace {

struct HostPath {
    bool live = false;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};    // audio[k] filled (copy stream)
    cudaEvent_t done[2] = {nullptr, nullptr};     // kernels finished with audio[k] (compute stream)
    cudaEvent_t t_beg = nullptr, t_end = nullptr; // device-side span of one call (first copy .. results on the host)
    void *audio[2] = {nullptr, nullptr};
    size_t audio_bytes = 0;
    void *stage[2] = {nullptr, nullptr};          // pinned staging for pageable callers
    size_t stage_bytes = 0;
    void *work = nullptr;
    size_t work_bytes = 0;
    char *res = nullptr;                          // device result arrays of one call
    size_t res_bytes = 0;
    void *res_host = nullptr;                     // pinned mirror of res: ONE device-to-host copy per call
    size_t res_host_bytes = 0;
};

constexpr int kMaxDevices = 64;
thread_local double g_last_host_ms = 0.0;         // ft8b_last_host_call_ms: the calling thread's most recent call
// One lock per device: calls that target the same device share its cached buffers and are serialised; calls on
// different devices (one host thread per GPU) run side by side.  The staging thread pool is shared by all devices and
// has its own lock (g_pool_mu), held only while a piece is being gathered.
std::mutex g_dev_mu[kMaxDevices];
std::mutex g_pool_mu;
HostPath g_host[kMaxDevices];

int grow(void **ptr, size_t *have, size_t need, bool pinned) {
    if (*have >= need && *ptr) return 0;
    if (*ptr) {
        FT8B_CUDA_TRY(pinned ? cudaFreeHost(*ptr) : cudaFree(*ptr));
        *ptr = nullptr;
        *have = 0;
    }
    FT8B_CUDA_TRY(pinned ? cudaHostAlloc(ptr, need, cudaHostAllocDefault) : cudaMalloc(ptr, need));
    *have = need;
    return 0;
}

int init_path(HostPath &h) {
    if (h.live) return 0;
    FT8B_CUDA_TRY(cudaStreamCreateWithFlags(&h.compute, cudaStreamNonBlocking));
    FT8B_CUDA_TRY(cudaStreamCreateWithFlags(&h.copy, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        FT8B_CUDA_TRY(cudaEventCreateWithFlags(&h.ready[k], cudaEventDisableTiming));
        FT8B_CUDA_TRY(cudaEventCreateWithFlags(&h.done[k], cudaEventDisableTiming));
    }
    FT8B_CUDA_TRY(cudaEventCreate(&h.t_beg));
    FT8B_CUDA_TRY(cudaEventCreate(&h.t_end));
    h.live = true;
    return 0;
}

void free_path(HostPath &h) {
    if (!h.live) return;
    for (int k = 0; k < 2; ++k) {
        if (h.audio[k]) cudaFree(h.audio[k]);
        if (h.stage[k]) cudaFreeHost(h.stage[k]);
        if (h.ready[k]) cudaEventDestroy(h.ready[k]);
        if (h.done[k]) cudaEventDestroy(h.done[k]);
    }
    if (h.t_beg) cudaEventDestroy(h.t_beg);
    if (h.t_end) cudaEventDestroy(h.t_end);
    if (h.work) cudaFree(h.work);
    if (h.res) cudaFree(h.res);
    if (h.res_host) cudaFreeHost(h.res_host);
    if (h.compute) cudaStreamDestroy(h.compute);
    if (h.copy) cudaStreamDestroy(h.copy);
    h = HostPath();
}

// Copy for the staging path: 16-byte non-temporal stores (the pinned destination is read next by the DMA engine, not by
// a core, so it should not be pulled into the caches first: a cached store costs a read-for-ownership on top).
void copy_stream(char *dst, const char *src, size_t n) {
#ifndef FT8B_HAVE_SSE2
    std::memcpy(dst, src, n);      // hosts without SSE2 (aarch64, e.g. Grace): the C library's copy
#else
    const size_t head = std::min(n, (size_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    std::memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
    }
    _mm_sfence();
    std::memcpy(dst + i, src + i, n - i);
#endif
}

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

// Persistent host threads that gather pageable caller memory into the pinned staging buffers.  A job is a list of
// blocks (256 KB pieces of a row) handed out through an atomic counter (cores of a shared box are not equally free); the calling thread
// works too and returns when every block is done.  Created on first use, joined when the library is unloaded.
class StagePool {
public:
    ~StagePool() {
        if (!threads_.empty() && getpid() != pid_) {      // forked child: the threads do not exist here, nothing to join
            new std::vector<std::thread>(std::move(threads_));
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // rows x row_bytes from a strided source into a dense destination
    void gather(char *dst, const char *src, int64_t rows, size_t row_bytes, size_t src_pitch) {
        const size_t total = (size_t)rows * row_bytes;
        if (rows <= 0 || row_bytes == 0) return;
        const int want = workers_wanted(total, rows);
        ensure_threads(want);
        {
            std::lock_guard<std::mutex> lk(mu_);
            const int64_t segs = (int64_t)((row_bytes + kSegBytes - 1) / kSegBytes);      // blocks per row
            job_ = Job{dst, src, rows * segs, segs, row_bytes, src_pitch, nt_};
            next_.store(0, std::memory_order_relaxed);
            left_.store(rows * segs, std::memory_order_relaxed);
            active_ = want;
            ++generation_;
        }
        if (want > 0) cv_.notify_all();
        work();
        // the last rows may still be in flight on other threads
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return left_.load(std::memory_order_acquire) == 0 && busy_ == 0; });
        active_ = 0;
    }

private:
    static constexpr size_t kSegBytes = 256 << 10;
    struct Job {
        char *dst;
        const char *src;
        int64_t blocks, segs;
        size_t row_bytes, src_pitch;
        bool nt;
    };
public:
    int configure(int threads, int nt) {       // threads <= 0: default; nt < 0: default
        const int hc = (int)std::thread::hardware_concurrency();
        // measured on the 16-thread B200 box (tools/stage_sweep.py): 1/2/4/6/8/10/12/16 threads stage 15.7/28.4/46.0/56.4/
        // 61.5/60.8/58.8/57.6 k slots/s with streaming stores (cached stores: 12.2 ... 53.9 k at 8) - the memory system
        // saturates at 8 copiers, so that is the default (and it leaves cores to the other ranks of a multi-GPU box)
        const int dflt = std::max(1, std::min(hc > 2 ? hc - 2 : 1, 8));
        n_cfg_ = std::max(1, std::min(threads > 0 ? threads : env_int("FT8B_STAGE_THREADS", dflt), 64));
        nt_ = (nt >= 0 ? nt : env_int("FT8B_STAGE_NT", 1)) != 0;
        min_bytes_ = (size_t)std::max(0, env_int("FT8B_STAGE_MIN_KB", 1024)) << 10;
        return n_cfg_;
    }

private:
    int workers_wanted(size_t total, int64_t rows) {
        if (n_cfg_ < 0) configure(0, -1);
        if (total < min_bytes_) return 0;       // small inputs: waking threads costs more than it saves
        return (int)std::min<int64_t>(n_cfg_ - 1, (int64_t)(total / kSegBytes));
    }
    void ensure_threads(int n) {
        if (threads_.empty()) pid_ = getpid();
        while ((int)threads_.size() < n) {
            const int id = (int)threads_.size();
            threads_.emplace_back([this, id] { loop(id); });
        }
    }
    void work() {
        const Job j = job_;
        for (;;) {
            const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= j.blocks) break;
            const int64_t r = i / j.segs;
            const size_t o = (size_t)(i - r * j.segs) * kSegBytes, n = std::min(kSegBytes, j.row_bytes - o);
            char *d = j.dst + (size_t)r * j.row_bytes + o;
            const char *s = j.src + (size_t)r * j.src_pitch + o;
            if (j.nt) copy_stream(d, s, n);
            else std::memcpy(d, s, n);
            left_.fetch_sub(1, std::memory_order_release);
        }
    }
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (generation_ != seen && id < active_); });
                if (stop_) return;
                seen = generation_;
                ++busy_;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu_);
                --busy_;
            }
            done_cv_.notify_one();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> threads_;
    Job job_{};
    std::atomic<int64_t> next_{0}, left_{0};
    uint64_t generation_ = 0;
    int active_ = 0, busy_ = 0, n_cfg_ = -1;
    bool nt_ = true, stop_ = false;
    size_t min_bytes_ = 1u << 20;
    pid_t pid_ = 0;
};

StagePool g_stage_pool;        // used under g_pool_mu only

void gather_rows(char *dst, const char *src, int64_t rows, size_t row_bytes, size_t src_pitch) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_stage_pool.gather(dst, src, rows, row_bytes, src_pitch);
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int release_host_paths() {
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < kMaxDevices; ++d) {
        std::lock_guard<std::mutex> lk(g_dev_mu[d]);
        if (!g_host[d].live) continue;
        cudaSetDevice(d);
        free_path(g_host[d]);
    }
    if (have_cur) cudaSetDevice(cur);
    cudaGetLastError();
    return 0;
}

}  // namespace ft8b

using namespace ft8b;

namespace {

struct HostCall {
    const void *wave;
    int wave_dtype;
    int64_t n_slots, n_samples, slot_stride;
    const ft8b_decode_params *p;
    int64_t piece_slots;
    void *dst[9];      // cand_time, cand_freq, cand_score, n_cand, ok, ldpc_errors, crc_extracted, crc_calculated, payload
    bool pinned;
};

// The body of one host-memory call on an initialised HostPath.  May return at any point with work still queued on
// h.copy / h.compute: the caller (ft8b_decode_slots_host) is the single exit path that drains both streams before
// the cached buffers can be touched again.
int host_call(HostPath &h, const HostCall &c) {
    const ft8b_decode_params *p = c.p;
    const int64_t n_slots = c.n_slots, K = p->max_candidates;
    const size_t elt = c.wave_dtype == FT8B_DTYPE_F32 ? 4 : c.wave_dtype == FT8B_DTYPE_I16 ? 2 : c.wave_dtype == FT8B_DTYPE_C128 ? 16 : 8;
    const size_t row_bytes = (size_t)c.n_samples * elt, src_pitch = (size_t)c.slot_stride * elt;
    const int64_t piece = std::min<int64_t>(n_slots, c.piece_slots > 0 ? c.piece_slots : 256);
    // ft8b_decode_workspace_bytes(piece) covers every batch of at most `piece` slots, i.e. the shorter tail piece too
    // (whose single chunk may be larger than the balanced chunks of a full piece)
    const int64_t work_need = ft8b_decode_workspace_bytes(p, piece, c.n_samples);
    if (work_need < 0) return (int)work_need;     // geometry error, message already set
    int rc = 0;
    const int n_buf = n_slots > piece ? 2 : 1;
    const size_t piece_bytes = std::max<size_t>((size_t)piece * row_bytes, 16);
    if (h.audio_bytes < piece_bytes) {
        for (int k = 0; k < 2; ++k) {
            size_t have = h.audio[k] ? h.audio_bytes : 0;
            if ((rc = grow(&h.audio[k], &have, piece_bytes, false))) { h.audio_bytes = 0; return rc; }
        }
        h.audio_bytes = piece_bytes;
    }
    if (!c.pinned && h.stage_bytes < piece_bytes) {
        for (int k = 0; k < 2; ++k) {
            size_t have = h.stage[k] ? h.stage_bytes : 0;
            if ((rc = grow(&h.stage[k], &have, piece_bytes, true))) { h.stage_bytes = 0; return rc; }
        }
        h.stage_bytes = piece_bytes;
    }
    if ((rc = grow(&h.work, &h.work_bytes, (size_t)work_need, false))) return rc;
    // device result arrays [n_slots][K]
    const size_t nK = (size_t)n_slots * (size_t)K;
    size_t off[10];
    off[0] = 0;                                        // cand_time
    off[1] = off[0] + align256(4 * nK);                // cand_freq
    off[2] = off[1] + align256(4 * nK);                // cand_score
    off[3] = off[2] + align256(8 * nK);                // n_cand
    off[4] = off[3] + align256(4 * (size_t)n_slots);   // ok
    off[5] = off[4] + align256(nK);                    // ldpc_errors
    off[6] = off[5] + align256(4 * nK);                // crc_extracted
    off[7] = off[6] + align256(4 * nK);                // crc_calculated
    off[8] = off[7] + align256(4 * nK);                // payload
    off[9] = off[8] + align256(10 * nK);
    {
        void *r = h.res;
        rc = grow(&r, &h.res_bytes, std::max<size_t>(off[9], 256), false);
        h.res = static_cast<char *>(r);
        if (rc) return rc;
        if ((rc = grow(&h.res_host, &h.res_host_bytes, std::max<size_t>(off[9], 256), true))) return rc;
    }
    FT8B_CUDA_TRY(cudaEventRecord(h.t_beg, h.compute));
    FT8B_CUDA_TRY(cudaMemsetAsync(h.res, 0, off[9], h.compute));
    FT8B_CUDA_TRY(cudaStreamWaitEvent(h.copy, h.t_beg, 0));      // no copy of this call starts before its time origin

    // Piece sizes: `piece` slots each, the end of a long batch tapered (..., piece/2, piece/4, piece/4): the kernels of
    // the last piece are the only ones no copy hides, so that piece is kept short (0.55 -> 0.15 ms exposed at 128 slots).
    std::vector<int64_t> sizes;
    {
        int64_t left = n_slots;
        while (left > piece) { sizes.push_back(piece); left -= piece; }
        if (n_slots > 2 * piece) {
            while (left > 32) { const int64_t half = (left + 1) / 2; sizes.push_back(half); left -= half; }
        }
        if (left > 0) sizes.push_back(left);
    }
    bool staged[2] = {false, false};
    int64_t lo = 0;
    for (size_t i = 0; i < sizes.size(); lo += sizes[i], ++i) {
        const int64_t n = sizes[i];
        const int k = (int)(i % n_buf);
        const char *src = static_cast<const char *>(c.wave) + (size_t)lo * src_pitch;
        size_t pitch = src_pitch;
        if (!c.pinned) {
            // stage[k] is free once the copy that last read it has finished
            if (staged[k]) FT8B_CUDA_TRY(cudaEventSynchronize(h.ready[k]));
            gather_rows(static_cast<char *>(h.stage[k]), src, n, row_bytes, src_pitch);
            src = static_cast<const char *>(h.stage[k]);
            pitch = row_bytes;
            staged[k] = true;
        }
        if ((int)i >= n_buf) FT8B_CUDA_TRY(cudaStreamWaitEvent(h.copy, h.done[k], 0));     // audio[k] no longer read
        if (row_bytes > 0)
            FT8B_CUDA_TRY(cudaMemcpy2DAsync(h.audio[k], row_bytes, src, pitch, row_bytes, (size_t)n, cudaMemcpyHostToDevice, h.copy));
        FT8B_CUDA_TRY(cudaEventRecord(h.ready[k], h.copy));
        FT8B_CUDA_TRY(cudaStreamWaitEvent(h.compute, h.ready[k], 0));
        rc = ft8b_decode_slots(h.audio[k], c.wave_dtype, n, c.n_samples, c.n_samples, p, h.work, (int64_t)h.work_bytes,
                               reinterpret_cast<int32_t *>(h.res + off[0]) + lo * K, reinterpret_cast<int32_t *>(h.res + off[1]) + lo * K,
                               reinterpret_cast<double *>(h.res + off[2]) + lo * K, reinterpret_cast<int32_t *>(h.res + off[3]) + lo,
                               reinterpret_cast<uint8_t *>(h.res + off[4]) + lo * K, reinterpret_cast<int32_t *>(h.res + off[5]) + lo * K,
                               reinterpret_cast<int32_t *>(h.res + off[6]) + lo * K, reinterpret_cast<int32_t *>(h.res + off[7]) + lo * K,
                               reinterpret_cast<uint8_t *>(h.res + off[8]) + lo * K * 10, h.compute);
        if (rc) return rc;
        FT8B_CUDA_TRY(cudaEventRecord(h.done[k], h.compute));
    }
    const size_t len[9] = {4 * nK, 4 * nK, 8 * nK, 4 * (size_t)n_slots, nK, 4 * nK, 4 * nK, 4 * nK, 10 * nK};
    // Results.  Large result sets into page-locked caller arrays (a pinned torch tensor, cudaHostAlloc memory): nine
    // device-to-host copies straight into them - no mirror, no host memcpy, no first-touch page faults on fresh
    // destination pages (0.5 ms of a 4096-slot call).  Otherwise: ONE copy of the whole result block into the pinned
    // mirror (nine copies into pageable arrays would each be staged and waited for by the driver, and nine copies of a
    // single-slot result cost more than they save), then plain memcpy into the caller's arrays.
    bool direct = off[9] >= (1u << 20);
    for (int a = 0; a < 9 && direct; ++a) {
        if (!c.dst[a] || !len[a]) continue;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, c.dst[a]) != cudaSuccess) { cudaGetLastError(); direct = false; }
        else if (at.type != cudaMemoryTypeHost) direct = false;
    }
    if (direct) {
        for (int a = 0; a < 9; ++a)
            if (len[a] && c.dst[a]) FT8B_CUDA_TRY(cudaMemcpyAsync(c.dst[a], h.res + off[a], len[a], cudaMemcpyDeviceToHost, h.compute));
    } else {
        FT8B_CUDA_TRY(cudaMemcpyAsync(h.res_host, h.res, off[9], cudaMemcpyDeviceToHost, h.compute));
    }
    FT8B_CUDA_TRY(cudaEventRecord(h.t_end, h.compute));
    FT8B_CUDA_TRY(cudaStreamSynchronize(h.compute));
    FT8B_CUDA_TRY(cudaStreamSynchronize(h.copy));
    if (!direct)
        for (int a = 0; a < 9; ++a)
            if (len[a] && c.dst[a]) std::memcpy(c.dst[a], static_cast<const char *>(h.res_host) + off[a], len[a]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h.t_beg, h.t_end) == cudaSuccess) g_last_host_ms = ms;
    else cudaGetLastError();
    return 0;
}

}  // namespace

extern "C" int ft8b_decode_slots_host(const void *wave, int wave_dtype, int64_t n_slots, int64_t n_samples, int64_t slot_stride,
                                      const ft8b_decode_params *p, int device, int64_t piece_slots, int32_t *cand_time,
                                      int32_t *cand_freq, double *cand_score, int32_t *n_cand, uint8_t *ok,
                                      int32_t *ldpc_errors, int32_t *crc_extracted, int32_t *crc_calculated,
                                      uint8_t *payload) {
    // crc_extracted / crc_calculated may be NULL (not wanted), as in ft8b_decode_slots; every other array is required
    if (!p || (!wave && n_slots > 0) || !cand_time || !cand_freq || !cand_score || !n_cand || !ok || !ldpc_errors || !payload)
        FT8B_ARG_FAIL("ft8b_decode_slots_host: null pointer");
    if (wave_dtype != FT8B_DTYPE_F32 && wave_dtype != FT8B_DTYPE_F64 && wave_dtype != FT8B_DTYPE_I16 &&
        wave_dtype != FT8B_DTYPE_C64 && wave_dtype != FT8B_DTYPE_C128)
        FT8B_ARG_FAIL("ft8b_decode_slots_host: bad dtype");
    if (n_slots < 0 || n_samples < 0 || slot_stride < n_samples) FT8B_ARG_FAIL("ft8b_decode_slots_host: bad shape");
    if (p->max_candidates < 0 || p->max_candidates > FT8B_MAX_CANDIDATES) {
        set_error("ft8b_decode_slots_host: max_candidates=%d outside [0,%d]", p->max_candidates, FT8B_MAX_CANDIDATES);
        return FT8B_ERR_UNSUPPORTED;
    }
    if (n_slots == 0) return 0;
    int dev = 0, n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) {
        cudaGetLastError();
        set_error("ft8b_decode_slots_host: no CUDA device (there is no CPU fallback)");
        return FT8B_ERR_NO_DEVICE;
    }
    FT8B_CUDA_TRY(cudaGetDevice(&dev));
    struct DeviceScope {      // run on `device` (if given) and put the caller's device back afterwards
        int prev, cur;
        ~DeviceScope() { if (cur != prev) cudaSetDevice(prev); }
    } scope{dev, dev};
    if (device >= 0 && device != dev) {
        if (device >= n_dev) FT8B_ARG_FAIL("ft8b_decode_slots_host: device %d of %d", device, n_dev);
        FT8B_CUDA_TRY(cudaSetDevice(device));
        scope.cur = dev = device;
    }
    if (dev < 0 || dev >= kMaxDevices) FT8B_ARG_FAIL("ft8b_decode_slots_host: device ordinal %d", dev);

    HostCall c{wave, wave_dtype, n_slots, n_samples, slot_stride, p, piece_slots,
               {cand_time, cand_freq, cand_score, n_cand, ok, ldpc_errors, crc_extracted, crc_calculated, payload}, false};
    // is the caller's buffer page-locked (then the DMA engine reads it directly)?
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, wave) == cudaSuccess) {
        if (attr.type == cudaMemoryTypeDevice) FT8B_ARG_FAIL("ft8b_decode_slots_host: wave is a device pointer (use ft8b_decode_slots)");
        c.pinned = attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
    } else {
        cudaGetLastError();      // an address CUDA knows nothing about: pageable
    }

    std::lock_guard<std::mutex> lk(g_dev_mu[dev]);
    HostPath &h = g_host[dev];
    int rc = init_path(h);
    if (rc == 0) rc = host_call(h, c);
    if (rc != 0) {
        // Single failure exit: nothing of this call may still be queued when the cached buffers are reused.  Drain both
        // streams; after a CUDA error (rc > 0) the streams, events and buffers are in an unknown state, so the whole
        // cached path is dropped and the next call builds a fresh one.  The error message set by the failing step is kept.
        if (h.copy) cudaStreamSynchronize(h.copy);
        if (h.compute) cudaStreamSynchronize(h.compute);
        if (rc > 0) {
            h.live = true;       // free_path() releases whatever init_path() managed to create
            free_path(h);
        }
        cudaGetLastError();
    }
    return rc;
}

extern "C" int ft8b_set_host_staging(int threads, int non_temporal) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    return g_stage_pool.configure(threads, non_temporal != 0 ? 1 : 0);
}

extern "C" int ft8b_selftest_host_staging(void *dst, const void *src, int64_t rows, int64_t row_bytes, int64_t src_pitch) {
    if (!dst || !src || rows < 0 || row_bytes < 0 || src_pitch < row_bytes) FT8B_ARG_FAIL("ft8b_selftest_host_staging: bad argument");
    gather_rows(static_cast<char *>(dst), static_cast<const char *>(src), rows, (size_t)row_bytes, (size_t)src_pitch);
    return 0;
}

extern "C" double ft8b_last_host_call_ms(void) { return g_last_host_ms; }
