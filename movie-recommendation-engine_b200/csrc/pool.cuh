// pool.cuh -- ragged neighbour-list preparation shared by the pooling kernel (P1-P4) and
// the fused conv kernel (G1).  One warp turns one row's padded list into a compacted
// (id, normalised weight) list in shared memory, applying the exact filtering / weight
// alignment rule of the reference class selected by `mode` (include/pinsage_b200.h).
#pragma once
#include "common.cuh"

namespace pb200 {

struct ListArgs {
    const int32_t* __restrict__ ids;      // [n, T]
    const float* __restrict__ weights;    // [n, T] or null
    const int32_t* __restrict__ list_len; // [n] or null (= T)
    const int32_t* __restrict__ weight_len;  // [n] or null (= list_len)
    int T;
    int mode;
    int64_t num_rows;  // rows of the gathered matrix (validity bound)
};

// Returns the number of valid neighbours (warp-uniform).  s_id / s_w: T entries each.
__device__ __forceinline__ int prepare_list(const ListArgs& a, int64_t row, int* s_id, float* s_w,
                                            int lane) {
    const int32_t* ids_row = a.ids + row * a.T;
    const float* w_row = a.weights ? a.weights + row * a.T : nullptr;
    int len = a.list_len ? a.list_len[row] : a.T;
    len = len < 0 ? 0 : (len > a.T ? a.T : len);
    int wlen = w_row ? (a.weight_len ? a.weight_len[row] : len) : 0;
    wlen = wlen < 0 ? 0 : (wlen > a.T ? a.T : wlen);
    int nv = 0;
    for (int base = 0; base < len; base += 32) {
        const int j = base + lane;
        const int id = j < len ? ids_row[j] : -1;
        // pinsage.py:124 (idx <= max_idx) / layers.py:109 (n < x.size(0)).  Negative ids index from the
        // end in the reference (python semantics): the host mirror rewrites them to id + num_rows (or
        // raises IndexError below -num_rows) before the lists reach the device
        // (neighbor_lists._wrap_negative); a negative id that still arrives here is padding.
        const bool valid = j < len && id >= 0 && (int64_t)id < a.num_rows;
        const unsigned m = __ballot_sync(kFull, valid);
        const int rank = nv + __popc(m & ((1u << lane) - 1u));
        if (valid) {
            s_id[rank] = id;
            float w = 1.0f;
            if (a.mode == PB200_POOL_PINSAGE)          // weight travels with its id; missing -> 1
                w = j < wlen ? w_row[j] : 1.0f;        // pinsage.py:126-129
            else if (a.mode == PB200_POOL_LAYERS)      // head of the weight list, layers.py:115
                w = rank < wlen ? w_row[rank] : 0.0f;
            else if (a.mode == PB200_POOL_AGGREGATOR)  // weights[:len], aggregators.py:74
                w = j < wlen ? w_row[j] : 0.0f;
            s_w[rank] = w;
        }
        nv += __popc(m);
    }
    __syncwarp();
    if (nv == 0 || a.mode == PB200_POOL_MAX) return nv;
    float s = 0.0f;
    for (int r = lane; r < nv; r += 32) s += s_w[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    const float uniform = 1.0f / (float)nv;
    for (int r = lane; r < nv; r += 32) {
        float w = s_w[r];
        if (a.mode == PB200_POOL_MEAN) w = uniform;
        else if (a.mode == PB200_POOL_PINSAGE) w = s > 0.0f ? w / s : w;   // pinsage.py:142-143
        else w = s == 0.0f ? uniform : w / s;          // layers.py:116-121, aggregators.py:78-84
        s_w[r] = w;
    }
    __syncwarp();
    return nv;
}

}  // namespace pb200
