// walk_bucket.cu -- direct-addressed sampling index ("bucket" leaf format, PB200_LEAF_BUCKET).
//
// The 8-ary tree index (walk_index.cu) still costs meta -> 1-3 index levels -> leaf: 3-5
// DEPENDENT loads per walk step and ~150 warp instructions.  ncu (profiles/r1_walk_compact_leaf_*)
// showed the walk kernel bound by issue slots + load latency, not by DRAM bytes.  This format
// cuts a step to TWO dependent loads (meta -> bucket) and ~35 instructions:
//
//   A row's weight axis [0, S) (S = row total in quanta) is cut into buckets of 2^s quanta
//   (s per row, 0..7).  Bucket j is one 32-byte block holding EVERY edge whose interval
//   [cum_{i-1}, cum_i) overlaps [j 2^s, (j+1) 2^s) -- at most 8 by the choice of s:
//       bytes  0.. 7   rel_i = min(cum_i - j 2^s, 2^s)  (1..128; unused slots hold 128)
//       bytes  8..15   neighbour id bits  0.. 7, one byte per slot
//       bytes 16..23   neighbour id bits  8..15
//       bytes 24..31   neighbour id bits 16..23
//   meta uint4 per node = {first bucket of the row, degree, S, s}.
//   A step draws t = floor(k53 S / 2^53) exactly as the flat search does, loads bucket
//   (t >> s) and takes the first slot with rel > (t & (2^s - 1)): the edge with the first
//   cum_i > t, i.e. the edge the flat search and the reference's inverse CDF select
//   (utils/random_walk.py:72-79).  The edge containing t overlaps the bucket containing t, so
//   it is always present.
//
//   s = the largest shift (<= 7) for which no bucket is overlapped by 9 edges: edges i..i+8
//   share a bucket iff (cum_i - 1) >> s == cum_{i+7} >> s, so
//   s = min(7, min_i msb((cum_i - 1) xor cum_{i+7})).  Needs every weight >= 1 quantum and node
//   ids < 2^24.  Size: ~S/2^s blocks per row (rating graphs: s = 4, ~11 B per edge) -- the price of
//   one sector per step.
//
//   PB200_LEAF_BUCKET32 (graphs with more than 2^24 nodes): the same with SIX slots per block --
//       bytes 0..5 rel_i, bytes 6..7 unused (128), bytes 8..31 six 32-bit neighbour ids;
//   the shift rule forbids 7 edges per bucket (xor with cum_{i+5}); ~16 B per edge on rating graphs.
//   Zero-weight edges or very heavy weights: the caller keeps the tree index.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace pb200 {

// one warp per row: shift, bucket count, zero-weight detection
template <int kSlots>
__global__ void __launch_bounds__(256) wbkt_plan_kernel(const int64_t* __restrict__ row_ptr,
                                                        const uint32_t* __restrict__ cum, int64_t N,
                                                        uint4* meta, unsigned long long* cnt,
                                                        unsigned long long* info) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long zeros = 0;
    for (int64_t v = wid; v < N; v += nw) {
        const int64_t r0 = row_ptr[v];
        const uint32_t deg = (uint32_t)(row_ptr[v + 1] - r0);
        const uint32_t S = deg ? cum[r0 + deg - 1] : 0u;
        int smin = 7;
        for (uint32_t i = lane; i < deg; i += 32) {
            const uint32_t b = cum[r0 + i];
            const uint32_t a = i ? cum[r0 + i - 1] : 0u;
            if (b == a) { ++zeros; continue; }
            if (i + kSlots < deg) {                          // edges i .. i + kSlots must not share a bucket
                const uint32_t x = (b - 1u) ^ cum[r0 + i + kSlots - 1];
                smin = min(smin, 31 - __clz((int)x));       // x != 0: cum[i + kSlots - 1] >= b > b - 1
            }
        }
        smin = __reduce_min_sync(kFull, smin);
        if (lane == 0) {
            meta[v] = make_uint4(0u, deg, S, (uint32_t)smin);
            cnt[v] = S ? (unsigned long long)((S - 1u) >> smin) + 1ull : 0ull;
        }
    }
    zeros = __reduce_add_sync(kFull, (unsigned)zeros);
    if (lane == 0 && zeros) atomicAdd(&info[1], zeros);
}

__global__ void wbkt_total_kernel(const unsigned long long* off, int64_t N, unsigned long long* info) {
    if (threadIdx.x == 0 && blockIdx.x == 0) info[0] = off[N];
}

// one thread per bucket
template <int kSlots>
__global__ void __launch_bounds__(256) wbkt_fill_kernel(const int64_t* __restrict__ row_ptr,
                                                        const int32_t* __restrict__ col,
                                                        const uint32_t* __restrict__ cum, int64_t N,
                                                        const unsigned long long* __restrict__ off,
                                                        uint4* meta, uint32_t* __restrict__ leaf,
                                                        unsigned long long total) {
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        // row = last v with off[v] <= g (rows without buckets share their successor's offset)
        int64_t lo = 0, hi = N;                       // invariant: off[lo] <= g < off[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (off[mid] <= g) lo = mid; else hi = mid;
        }
        const int64_t v = lo;
        const uint32_t j = (uint32_t)(g - off[v]);
        const uint4 m = meta[v];
        const int64_t r0 = row_ptr[v];
        const uint32_t deg = m.y, s = m.w;
        const uint32_t w = 1u << s;
        const uint64_t base = (uint64_t)j << s;       // bucket covers [base, base + w)
        if (j == 0) meta[v].x = (uint32_t)off[v];
        // first edge with cum > base
        uint32_t a = 0, b = deg - 1;
        while (a < b) {
            const uint32_t mid = a + ((b - a) >> 1);
            if ((uint64_t)cum[r0 + mid] > base) b = mid; else a = mid + 1;
        }
        uint32_t rel[8], id[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t e = a + q;
            bool valid = q < kSlots && e < deg;
            if (valid && q) valid = (uint64_t)cum[r0 + e - 1] < base + w;   // starts inside the bucket
            if (valid) {
                const uint64_t d = (uint64_t)cum[r0 + e] - base;
                rel[q] = d < w ? (uint32_t)d : w;
                id[q] = (uint32_t)col[r0 + e];
            } else {
                rel[q] = 128u; id[q] = 0u;
            }
        }
        uint32_t o[8];
        o[0] = rel[0] | (rel[1] << 8) | (rel[2] << 16) | (rel[3] << 24);
        o[1] = rel[4] | (rel[5] << 8) | (rel[6] << 16) | (rel[7] << 24);
        if (kSlots == 6) {
#pragma unroll
            for (int q = 0; q < 6; ++q) o[2 + q] = id[q];
        } else {
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                const int sh = 8 * pl;
                o[2 + 2 * pl] = ((id[0] >> sh) & 255u) | (((id[1] >> sh) & 255u) << 8) |
                                (((id[2] >> sh) & 255u) << 16) | (((id[3] >> sh) & 255u) << 24);
                o[3 + 2 * pl] = ((id[4] >> sh) & 255u) | (((id[5] >> sh) & 255u) << 8) |
                                (((id[6] >> sh) & 255u) << 16) | (((id[7] >> sh) & 255u) << 24);
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(leaf + g * 8ull);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

struct WbktWs { unsigned long long* cnt; void* temp; size_t temp_bytes, total; };
static WbktWs wbkt_carve(void* base, int64_t N) {
    WbktWs w{};
    cub::DeviceScan::ExclusiveSum(nullptr, w.temp_bytes, (unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (uint32_t)(N + 1));
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
    w.cnt = (unsigned long long*)take((size_t)(N + 1) * 8);
    w.temp = take(w.temp_bytes);
    w.total = off;
    return w;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_walk_bucket_workspace_bytes(int64_t num_nodes) {
    return wbkt_carve(nullptr, num_nodes > 0 ? num_nodes : 0).total;
}

extern "C" int pb200_walk_bucket_plan_ex(const int64_t* row_ptr, const void* cum, int64_t num_nodes,
                                         uint32_t* meta, uint64_t* info_out, void* workspace,
                                         size_t workspace_bytes, int leaf_format, pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(num_nodes > 0 && num_nodes < 2147483647ll && row_ptr && cum && meta && info_out && workspace,
               "walk_bucket_plan: bad arguments");
    PB_REQUIRE(leaf_format == PB200_LEAF_BUCKET || leaf_format == PB200_LEAF_BUCKET32, "walk_bucket_plan: unknown format");
    PB_REQUIRE((uintptr_t)meta % 16 == 0, "walk_bucket_plan: meta must be 16-byte aligned");
    WbktWs w = wbkt_carve(workspace, num_nodes);
    if (workspace_bytes < w.total) {
        set_error("walk_bucket_plan: workspace %zu B < required %zu B", workspace_bytes, w.total);
        return PB200_ERR_WORKSPACE;
    }
    PB_CUDA(cudaMemsetAsync(info_out, 0, 2 * sizeof(uint64_t), stream));
    PB_CUDA(cudaMemsetAsync(w.cnt + num_nodes, 0, sizeof(unsigned long long), stream));
    const unsigned blocks = (unsigned)(ceil_div(num_nodes, 8) < kSMs * 16 ? ceil_div(num_nodes, 8) : kSMs * 16);
    if (leaf_format == PB200_LEAF_BUCKET32)
        wbkt_plan_kernel<6><<<blocks, 256, 0, stream>>>(row_ptr, static_cast<const uint32_t*>(cum), num_nodes,
                                                        reinterpret_cast<uint4*>(meta), w.cnt,
                                                        reinterpret_cast<unsigned long long*>(info_out));
    else
        wbkt_plan_kernel<8><<<blocks, 256, 0, stream>>>(row_ptr, static_cast<const uint32_t*>(cum), num_nodes,
                                                        reinterpret_cast<uint4*>(meta), w.cnt,
                                                        reinterpret_cast<unsigned long long*>(info_out));
    int rc = check_launch("wbkt_plan_kernel");
    if (rc) return rc;
    size_t tb = w.temp_bytes;
    PB_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, w.cnt, w.cnt, (uint32_t)(num_nodes + 1), stream));
    count_launch(2);
    wbkt_total_kernel<<<1, 32, 0, stream>>>(w.cnt, num_nodes, reinterpret_cast<unsigned long long*>(info_out));
    return check_launch("wbkt_total_kernel");
}

extern "C" int pb200_walk_bucket_plan(const int64_t* row_ptr, const void* cum, int64_t num_nodes,
                                      uint32_t* meta, uint64_t* info_out, void* workspace,
                                      size_t workspace_bytes, pb200_stream_t stream) {
    return pb200_walk_bucket_plan_ex(row_ptr, cum, num_nodes, meta, info_out, workspace, workspace_bytes,
                                     PB200_LEAF_BUCKET, stream);
}

extern "C" int pb200_walk_bucket_fill_ex(const int64_t* row_ptr, const int32_t* col, const void* cum,
                                         int64_t num_nodes, const void* workspace, uint32_t* meta,
                                         uint32_t* leaf, uint64_t total_buckets, int leaf_format, pb200_stream_t stream) {
    PB_REQUIRE(num_nodes > 0 && row_ptr && col && cum && workspace && meta && leaf,
               "walk_bucket_fill: bad arguments");
    PB_REQUIRE(leaf_format == PB200_LEAF_BUCKET || leaf_format == PB200_LEAF_BUCKET32, "walk_bucket_fill: unknown format");
    PB_REQUIRE(leaf_format == PB200_LEAF_BUCKET32 || num_nodes <= (1 << 24),
               "walk_bucket_fill: PB200_LEAF_BUCKET holds 24-bit node ids (use PB200_LEAF_BUCKET32)");
    PB_REQUIRE(total_buckets < 4294967296ull, "walk_bucket_fill: more than 2^32 buckets");
    PB_REQUIRE((uintptr_t)leaf % 32 == 0 && (uintptr_t)meta % 16 == 0, "walk_bucket_fill: leaf/meta alignment");
    if (total_buckets == 0) return PB200_OK;
    WbktWs w = wbkt_carve(const_cast<void*>(workspace), num_nodes);
    const int64_t want = ceil_div((int64_t)total_buckets, 256);
    const unsigned blocks = (unsigned)(want < (int64_t)kSMs * 32 ? want : (int64_t)kSMs * 32);
    if (leaf_format == PB200_LEAF_BUCKET32)
        wbkt_fill_kernel<6><<<blocks, 256, 0, (cudaStream_t)stream>>>(
            row_ptr, col, static_cast<const uint32_t*>(cum), num_nodes, w.cnt, reinterpret_cast<uint4*>(meta), leaf,
            (unsigned long long)total_buckets);
    else
        wbkt_fill_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(
            row_ptr, col, static_cast<const uint32_t*>(cum), num_nodes, w.cnt, reinterpret_cast<uint4*>(meta), leaf,
            (unsigned long long)total_buckets);
    return check_launch("wbkt_fill_kernel");
}

extern "C" int pb200_walk_bucket_fill(const int64_t* row_ptr, const int32_t* col, const void* cum,
                                      int64_t num_nodes, const void* workspace, uint32_t* meta,
                                      uint32_t* leaf, uint64_t total_buckets, pb200_stream_t stream) {
    return pb200_walk_bucket_fill_ex(row_ptr, col, cum, num_nodes, workspace, meta, leaf, total_buckets,
                                     PB200_LEAF_BUCKET, stream);
}
