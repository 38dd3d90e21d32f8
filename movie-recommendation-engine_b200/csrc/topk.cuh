// topk.cuh -- internal interface of the fp32 exact-search kernels (topk.cu), shared with the
// tensor-core path (topk_tc.cu), which re-runs its uncertified queries through them.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace pb200 {

// out[r] = sum_c x[r,c]^2 (warp per row; the reduction order both search paths share)
int row_sqnorm_run(const float* x, int64_t n, int d, float* out, cudaStream_t stream);

// fp32 tile kernel + merge, all passes.  part_bad / part_ids: [nq, splits, 32] each.  With qsel
// set, slot i stands for query qsel[qsel_base + i] and only slots below *qsel_count - qsel_base
// run (device-side count: no host synchronisation).
int topk_fp32_run(const float* queries, int64_t nq, const float* items, int64_t nx, int dim, int k,
                  int metric, const float* qn, const float* xn, const int32_t* exclude_ids,
                  int32_t id_offset, float* out_scores, int32_t* out_ids, float* part_bad,
                  int32_t* part_ids, int splits, const int32_t* qsel, const int32_t* qsel_count,
                  int64_t qsel_base, cudaStream_t stream);

// One merge pass (k <= 32) of candidate lists vals/ids [nq, c] into out rows (slot -> qsel row).
int topk_merge_run(const float* vals, const int32_t* ids, int64_t nq, int c, int vals_are_bad, int largest,
                   int k, float* out_scores, int32_t* out_ids, const int32_t* qsel,
                   const int32_t* qsel_count, int64_t qsel_base, cudaStream_t stream);

}  // namespace pb200
