// ivf.cuh -- internal interface of the list-scan IVF kernel (ivf.cu), shared with the
// tensor-core path (topk_tc.cu), which re-runs its uncertified queries through it.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace pb200 {

// With qsel set, slot i stands for query qsel[qsel_base + i] and only slots below
// *qsel_count - qsel_base run (device-side count).  With part_bad / part_ids set
// ([nq][nprobe][32] each) one warp scans ONE probed list of one query and writes a partial
// list there instead of the outputs (merge with topk_merge_run); otherwise one warp per query.
// floor_dist / floor_id ([nq], optional): only candidates strictly worse than the floor are considered (k > 32 in passes).
int ivf_search_run(const float* queries, int64_t nq, int dim, const int32_t* probes, int nprobe,
                   const int32_t* list_offsets, const int32_t* list_ids, const float* list_vecs, int k,
                   float* out_dist, int32_t* out_ids, const int32_t* qsel, const int32_t* qsel_count,
                   int64_t qsel_base, float* part_bad, int32_t* part_ids, cudaStream_t stream,
                   const float* floor_dist = nullptr, const int32_t* floor_id = nullptr);

}  // namespace pb200
