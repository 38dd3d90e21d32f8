// common.cuh -- shared helpers for libpinsage_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pinsage_b200.h"

namespace pb200 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Returns PB200_OK or PB200_ERR_CUDA after a launch (no device sync).
int check_launch(const char* what);

#define PB_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            ::pb200::set_error(__VA_ARGS__);          \
            return PB200_ERR_INVALID_ARG;             \
        }                                             \
    } while (0)

#define PB_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            ::pb200::set_error("%s failed: %s", #call, cudaGetErrorString(_e));    \
            return PB200_ERR_CUDA;                                                 \
        }                                                                          \
    } while (0)

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- warp-level streaming top-k: lane r holds the r-th best (k <= 32) -------------
// Ordering: smaller `bad` first, ties by smaller id.  Entries start at (+inf, INT_MAX).
struct TopkLane {
    float bad;
    int id;
};

__device__ __forceinline__ bool better(float b1, int i1, float b2, int i2) {
    return b1 < b2 || (b1 == b2 && i1 < i2);
}

// All 32 lanes call with the same (b, id).  Inserts when it beats rank k-1.
__device__ __forceinline__ void topk_insert(TopkLane& e, float b, int id, int k, int lane) {
    const bool mine_better = (lane < k) && better(e.bad, e.id, b, id);
    const int pos = __popc(__ballot_sync(kFull, mine_better));
    if (pos >= k) return;  // warp-uniform
    const float pb = __shfl_up_sync(kFull, e.bad, 1);
    const int pi = __shfl_up_sync(kFull, e.id, 1);
    if (lane > pos) { e.bad = pb; e.id = pi; }
    else if (lane == pos) { e.bad = b; e.id = id; }
}

// Each lane offers one candidate (valid or not); candidates beating the current k-th
// best are inserted one by one (rare once the threshold has tightened).
__device__ __forceinline__ void topk_offer(TopkLane& e, float b, int id, bool valid, int k,
                                           int lane) {
    float tb = __shfl_sync(kFull, e.bad, k - 1);
    int ti = __shfl_sync(kFull, e.id, k - 1);
    unsigned m = __ballot_sync(kFull, valid && better(b, id, tb, ti));
    while (m) {
        const int src = __ffs(m) - 1;
        const float cb = __shfl_sync(kFull, b, src);
        const int ci = __shfl_sync(kFull, id, src);
        topk_insert(e, cb, ci, k, lane);
        tb = __shfl_sync(kFull, e.bad, k - 1);
        ti = __shfl_sync(kFull, e.id, k - 1);
        if (lane == src) valid = false;
        m = __ballot_sync(kFull, valid && better(b, id, tb, ti));
    }
}

}  // namespace pb200
