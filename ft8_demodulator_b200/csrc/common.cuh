// common.cuh - shared host-side plumbing of libft8b200 (error slot, launch counter, checks).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/ft8b200.h"

namespace ft8b {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

// FT8 protocol constants (ft8_decode.py:30-42)
constexpr int kND = 58;          // data symbols
constexpr int kNumSync = 3;      // Costas blocks
constexpr int kLenSync = 7;      // symbols per Costas block
constexpr int kSyncOffset = 36;  // symbols between Costas block starts
constexpr int kLdpcN = 174, kLdpcM = 83, kLdpcK = 91;

// arg-max of one warp strip of the score map, written by the sync kernel for the selection (sync_pair.cuh)
struct SyncTile {
    double score;
    long long idx;      // flat scan index (abs_time row * nf + abs_freq); LLONG_MAX when the strip is empty
};

}  // namespace ft8b

#define FT8B_CUDA_TRY(expr)                                         \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return ft8b::cuda_fail(_e, #expr);   \
    } while (0)

#define FT8B_LAUNCH_CHECK(name)                                     \
    do {                                                            \
        ft8b::g_launches.fetch_add(1, std::memory_order_relaxed);   \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return ft8b::cuda_fail(_e, name);    \
    } while (0)

#define FT8B_ARG_FAIL(...)               \
    do {                                 \
        ft8b::set_error(__VA_ARGS__);    \
        return FT8B_ERR_ARG;             \
    } while (0)
