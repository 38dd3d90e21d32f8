// ppr.cu -- N4: RandomWalkSampler.compute_ppr_matrix / precompute_top_neighbors
//                                                       (reference utils/random_walk.py:144-229).
//
// The reference runs, per source node, `num_iterations` in-place sweeps over ALL nodes in index order:
//     res = residual[v];  if res > 0:  ppr[v] += alpha res;  residual[n] += (1-alpha) res w_vn / sum_v  for
//     every out-edge (v, n) in edge order;  residual[v] = 0
// (dense float64 vectors; ppr[source] and residual[source] start at 1.0).  The sweep is sequential by
// definition -- a push to a higher index is consumed in the same sweep, one to a lower index in the next
// -- so the parallelism is across sources: one warp per source, dense vectors in HBM, nodes scanned 32 at
// a time (coalesced), active nodes processed one after the other in index order with the lanes spread
// over the node's edges.  Float64 adds happen in the reference's order except for parallel multi-edges
// to the same neighbour (two atomic adds in either order: <= 1 ulp).
#include "common.cuh"

namespace pb200 {

template <typename CumT>
__global__ void __launch_bounds__(256) ppr_push_kernel(const int64_t* __restrict__ row_ptr,
                                                       const int32_t* __restrict__ col,
                                                       const CumT* __restrict__ cum, double inv_scale,
                                                       int64_t num_nodes, int64_t vec_len,
                                                       const int32_t* __restrict__ sources, int64_t S, double alpha,
                                                       int iters, double* __restrict__ ppr_all,
                                                       double* __restrict__ res_all) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = wid; s < S; s += nw) {
        double* ppr = ppr_all + s * vec_len;
        volatile double* res = res_all + s * vec_len;
        for (int64_t i = lane; i < vec_len; i += 32) { ppr[i] = 0.0; res[i] = 0.0; }
        __syncwarp();
        if (lane == 0) { ppr[sources[s]] = 1.0; res[sources[s]] = 1.0; }
        __syncwarp();
        for (int it = 0; it < iters; ++it) {
            for (int64_t base = 0; base < vec_len; base += 32) {
                const int64_t mine = base + lane;
                unsigned mask = __ballot_sync(kFull, mine < vec_len && res[mine] > 0.0);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    const int64_t v = base + j;
                    const double r = res[v];                     // current value: earlier nodes of this chunk may have pushed here
                    if (lane == 0) ppr[v] += alpha * r;
                    if (v < num_nodes) {
                        const int64_t r0 = row_ptr[v], r1 = row_ptr[v + 1];
                        if (r1 > r0) {
                            const double total = (double)cum[r1 - 1] * inv_scale;
                            const double push = (1.0 - alpha) * r;
                            for (int64_t e = r0 + lane; e < r1; e += 32) {
                                const double w = ((double)cum[e] - (e > r0 ? (double)cum[e - 1] : 0.0)) * inv_scale;
                                atomicAdd(const_cast<double*>(res) + col[e], push * (w / total));
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) res[v] = 0.0;                 // after the pushes: a self loop's share is dropped, as in the reference
                    __threadfence_block();
                    __syncwarp();
                    // nodes after v in this chunk may have become active
                    mask = __ballot_sync(kFull, mine < vec_len && lane > j && res[mine] > 0.0);
                }
            }
        }
    }
}

// Per row: the k largest strictly positive scores, ties by smaller index (the reference's stable reverse sort
// of the (target, score) list in target order).  One warp per row, k passes.
__global__ void __launch_bounds__(256) topk_rows_f64_kernel(const double* __restrict__ scores, int64_t S,
                                                            int64_t n, int k, int32_t* __restrict__ out_ids,
                                                            double* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = wid; s < S; s += nw) {
        const double* row = scores + s * n;
        double last = INFINITY; long long last_id = -1;
        for (int j = 0; j < k; ++j) {
            double best = 0.0; long long best_id = -1;
            for (int64_t i = lane; i < n; i += 32) {
                const double v = row[i];
                const bool eligible = v > 0.0 && (v < last || (v == last && i > last_id));
                if (eligible && (v > best || (v == best && (best_id < 0 || i < best_id)))) { best = v; best_id = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(kFull, best, o);
                const long long oi = __shfl_xor_sync(kFull, best_id, o);
                if (oi >= 0 && (best_id < 0 || ov > best || (ov == best && oi < best_id))) { best = ov; best_id = oi; }
            }
            if (lane == 0) { out_ids[s * k + j] = (int32_t)best_id; out_scores[s * k + j] = best_id >= 0 ? best : 0.0; }
            if (best_id < 0) {
                for (int jj = j + 1 + lane; jj < k; jj += 32) { out_ids[s * k + jj] = -1; out_scores[s * k + jj] = 0.0; }
                break;
            }
            last = best; last_id = best_id;
        }
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_ppr_push(const int64_t* row_ptr, const int32_t* col, const void* cum, int cum_kind,
                              int quant_shift, int64_t num_nodes, int64_t vec_len, const int32_t* sources,
                              int64_t num_sources, double alpha, int num_iterations, double* ppr_out,
                              double* residual_ws, pb200_stream_t stream) {
    PB_REQUIRE(num_nodes >= 0 && vec_len >= num_nodes && num_sources >= 0 && num_iterations >= 0,
               "ppr_push: bad sizes");
    PB_REQUIRE(cum_kind == 0 || cum_kind == 1, "ppr_push: cum_kind must be 0 (u32 quanta) or 1 (f64)");
    if (num_sources == 0 || vec_len == 0) return PB200_OK;
    PB_REQUIRE(row_ptr && sources && ppr_out && residual_ws, "ppr_push: null pointer");
    int64_t blocks = ceil_div(num_sources, 8);
    if (blocks > (int64_t)kSMs * 8) blocks = (int64_t)kSMs * 8;
    if (cum_kind == 0)
        ppr_push_kernel<uint32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            row_ptr, col, static_cast<const uint32_t*>(cum), 1.0 / (double)(1u << (quant_shift > 0 ? quant_shift : 0)),
            num_nodes, vec_len, sources, num_sources, alpha, num_iterations, ppr_out, residual_ws);
    else
        ppr_push_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            row_ptr, col, static_cast<const double*>(cum), 1.0, num_nodes, vec_len, sources, num_sources, alpha,
            num_iterations, ppr_out, residual_ws);
    return check_launch("ppr_push_kernel");
}

extern "C" int pb200_topk_rows_f64(const double* scores, int64_t num_rows, int64_t row_len, int k, int32_t* out_ids,
                                   double* out_scores, pb200_stream_t stream) {
    PB_REQUIRE(num_rows >= 0 && row_len >= 0 && k >= 1, "topk_rows_f64: bad sizes");
    if (num_rows == 0) return PB200_OK;
    PB_REQUIRE(scores && out_ids && out_scores, "topk_rows_f64: null pointer");
    int64_t blocks = ceil_div(num_rows, 8);
    if (blocks > (int64_t)kSMs * 8) blocks = (int64_t)kSMs * 8;
    topk_rows_f64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(scores, num_rows, row_len, k, out_ids, out_scores);
    return check_launch("topk_rows_f64_kernel");
}
