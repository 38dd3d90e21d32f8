// walk_index.cu -- sampling index over the CSR: an implicit 8-ary search tree per row.
//
// The flat walk step (walk_topt.cu) binary-searches a row's cumulative weights: ~log2(deg)
// DEPENDENT probes, each its own 32 B sector.  ncu on config C2 showed the kernel latency
// bound (long-scoreboard stalls, 30 % of HBM peak, 2.6x the algorithmic DRAM bytes).  This
// index turns a step into  meta -> [upper level(s)] -> leaf :
//   meta  uint4 per node  {leaf block offset, degree, row total, upper-level block offset}
//   leaf  64 B block      {8 cumulative weights (uint32, 0xFFFFFFFF padded), 8 neighbour ids}
//         or, "compact" (PB200_LEAF_COMPACT: < 2^24 nodes and every block spans <= 255 weight
//         quanta -- true for rating graphs), a 32 B block {8 x u8 (block separator - cum),
//         8 x u16 id low halves, 8 x u8 id high bytes}: ONE 256-bit load per step instead of two
//         (the walk kernel is bound by LSU wavefronts: one per lane and load) and half the
//         DRAM bytes; the separator is the parent key that selected the block (or the row total)
//   idx   32 B block      8 separator keys = last cumulative weight under each child block
// Every node of the tree is one 256-bit load (LDG.E.256).  The upper levels take
// ~E/7 * 4 B (about 30 MB at ML-25M scale) and stay L2-resident (evict_last), so a step costs
// one DRAM access (the 64 B leaf) instead of ~6, and 3-4 dependent loads instead of ~10.
// The selected edge is identical to the flat search: first edge with cum > t.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace pb200 {

__host__ __device__ __forceinline__ uint32_t upper_blocks(uint32_t nb0) {
    uint32_t tot = 0;
    while (nb0 > 1) { nb0 = (nb0 + 7) >> 3; tot += nb0; }
    return tot;
}

__global__ void widx_count_kernel(const int64_t* __restrict__ row_ptr, int64_t N,
                                  uint32_t* leaf_cnt, uint32_t* idx_cnt) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < N;
         v += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t deg = (uint32_t)(row_ptr[v + 1] - row_ptr[v]);
        const uint32_t nb0 = (deg + 7) >> 3;
        leaf_cnt[v] = nb0;
        idx_cnt[v] = upper_blocks(nb0);
    }
}

__global__ void widx_sizes_kernel(const uint32_t* leaf_off, const uint32_t* idx_off,
                                  const int64_t* row_ptr, int64_t N, int64_t* sizes) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (N == 0) { sizes[0] = 0; sizes[1] = 0; return; }
        const uint32_t deg = (uint32_t)(row_ptr[N] - row_ptr[N - 1]);
        const uint32_t nb0 = (deg + 7) >> 3;
        sizes[0] = (int64_t)leaf_off[N - 1] + nb0;
        sizes[1] = (int64_t)idx_off[N - 1] + upper_blocks(nb0);
    }
}

// largest (last cum - first cum) over all 8-edge leaf blocks: decides whether the compact leaf fits
__global__ void __launch_bounds__(256) widx_range_kernel(const int64_t* __restrict__ row_ptr,
                                                         const uint32_t* __restrict__ cum, int64_t N,
                                                         uint32_t* max_range) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t m = 0;
    for (int64_t v = wid; v < N; v += nw) {
        const int64_t r0 = row_ptr[v];
        const uint32_t deg = (uint32_t)(row_ptr[v + 1] - r0);
        const uint32_t nb0 = (deg + 7) >> 3;
        for (uint32_t b = lane; b < nb0; b += 32) {
            const uint32_t last = min(8u * b + 7u, deg - 1u);
            m = max(m, cum[r0 + last] - cum[r0 + 8u * b]);
        }
    }
    m = __reduce_max_sync(kFull, m);
    if (lane == 0 && m) atomicMax(max_range, m);
}

// one warp per row: meta, leaf blocks, upper levels (top level first)
__global__ void __launch_bounds__(256) widx_fill_kernel(const int64_t* __restrict__ row_ptr,
                                                        const int32_t* __restrict__ col,
                                                        const uint32_t* __restrict__ cum, int64_t N,
                                                        const uint32_t* __restrict__ leaf_off,
                                                        const uint32_t* __restrict__ idx_off,
                                                        uint4* meta, uint32_t* idx, uint32_t* leaf, int compact) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = wid; v < N; v += nw) {
        const int64_t r0 = row_ptr[v];
        const uint32_t deg = (uint32_t)(row_ptr[v + 1] - r0);
        const uint32_t nb0 = (deg + 7) >> 3;
        const uint32_t lo = leaf_off[v], io = idx_off[v];
        if (lane == 0)
            meta[v] = make_uint4(lo, deg, deg ? cum[r0 + deg - 1] : 0u, io);
        // leaves
        for (uint32_t i = lane; i < nb0 * 8; i += 32) {
            const uint32_t b = i >> 3, j = i & 7;
            if (compact) {
                uint8_t* blk = reinterpret_cast<uint8_t*>(leaf + ((size_t)lo + b) * 8);
                const uint32_t sep = cum[r0 + min(8u * b + 7u, deg - 1u)];
                const uint32_t id = i < deg ? (uint32_t)col[r0 + i] : 0xFFFFFFu;
                blk[j] = i < deg ? (uint8_t)(sep - cum[r0 + i]) : (uint8_t)0;   // padding: never >= a positive threshold
                reinterpret_cast<uint16_t*>(blk + 8)[j] = (uint16_t)(id & 0xFFFFu);
                blk[24 + j] = (uint8_t)(id >> 16);
            } else {
                uint32_t* blk = leaf + ((size_t)lo + b) * 16;
                blk[j] = i < deg ? cum[r0 + i] : 0xFFFFFFFFu;
                blk[8 + j] = i < deg ? (uint32_t)col[r0 + i] : 0xFFFFFFFFu;
            }
        }
        // upper levels: level l (>=1) has one key per block of level l-1; key j of level l is
        // the last cumulative weight covered by that block: cum[min(8^l (j+1), deg) - 1]
        uint32_t nbl[8];
        int L = 0;
        nbl[0] = nb0;
        while (nbl[L] > 1) { nbl[L + 1] = (nbl[L] + 7) >> 3; ++L; }   // L upper levels
        uint32_t off = io;
        for (int l = L; l >= 1; --l) {
            const uint32_t nkeys = nbl[l - 1];
            const uint64_t span = 1ull << (3 * l);
            for (uint32_t i = lane; i < nbl[l] * 8; i += 32) {
                uint32_t key = 0xFFFFFFFFu;
                if (i < nkeys) {
                    uint64_t last = span * (i + 1);
                    if (last > deg) last = deg;
                    key = cum[r0 + last - 1];
                }
                idx[((size_t)off) * 8 + i] = key;
            }
            off += nbl[l];
        }
    }
}

struct WidxWs { uint32_t *leaf_cnt, *idx_cnt, *leaf_off, *idx_off; void* temp; size_t temp_bytes, total; };
static WidxWs widx_carve(void* base, int64_t N) {
    WidxWs w{};
    cub::DeviceScan::ExclusiveSum(nullptr, w.temp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t)(N > 0 ? N : 1));
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
    const size_t n = (size_t)(N > 0 ? N : 1);
    w.leaf_cnt = (uint32_t*)take(n * 4); w.idx_cnt = (uint32_t*)take(n * 4);
    w.leaf_off = (uint32_t*)take(n * 4); w.idx_off = (uint32_t*)take(n * 4);
    w.temp = take(w.temp_bytes);
    w.total = off;
    return w;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_walk_index_workspace_bytes(int64_t num_nodes) {
    return widx_carve(nullptr, num_nodes).total;
}

extern "C" int pb200_walk_index_sizes(const int64_t* row_ptr, int64_t num_nodes, int64_t* sizes_out,
                                      void* workspace, size_t workspace_bytes,
                                      pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(row_ptr && sizes_out && workspace && num_nodes >= 0 && num_nodes < 2147483647ll,
               "walk_index_sizes: bad arguments");
    WidxWs w = widx_carve(workspace, num_nodes);
    if (workspace_bytes < w.total) {
        set_error("walk_index_sizes: workspace %zu B < required %zu B", workspace_bytes, w.total);
        return PB200_ERR_WORKSPACE;
    }
    if (num_nodes > 0) {
        const unsigned blocks = (unsigned)(ceil_div(num_nodes, 256) < kSMs * 8 ? ceil_div(num_nodes, 256) : kSMs * 8);
        widx_count_kernel<<<blocks, 256, 0, stream>>>(row_ptr, num_nodes, w.leaf_cnt, w.idx_cnt);
        int rc = check_launch("widx_count_kernel");
        if (rc) return rc;
        size_t tb = w.temp_bytes;
        PB_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, w.leaf_cnt, w.leaf_off, (uint32_t)num_nodes, stream));
        tb = w.temp_bytes;
        PB_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, w.idx_cnt, w.idx_off, (uint32_t)num_nodes, stream));
        count_launch(4);
    }
    widx_sizes_kernel<<<1, 32, 0, stream>>>(w.leaf_off, w.idx_off, row_ptr, num_nodes, sizes_out);
    return check_launch("widx_sizes_kernel");
}

extern "C" int pb200_walk_index_leaf_range(const int64_t* row_ptr, const void* cum, int64_t num_nodes,
                                           uint32_t* max_range_out, pb200_stream_t stream) {
    PB_REQUIRE(num_nodes >= 0 && max_range_out, "walk_index_leaf_range: bad arguments");
    PB_CUDA(cudaMemsetAsync(max_range_out, 0, sizeof(uint32_t), (cudaStream_t)stream));
    if (num_nodes == 0) return PB200_OK;
    PB_REQUIRE(row_ptr && cum, "walk_index_leaf_range: null pointer");
    const unsigned blocks = (unsigned)(ceil_div(num_nodes, 8) < kSMs * 16 ? ceil_div(num_nodes, 8) : kSMs * 16);
    widx_range_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(row_ptr, static_cast<const uint32_t*>(cum), num_nodes,
                                                               max_range_out);
    return check_launch("widx_range_kernel");
}

extern "C" int pb200_walk_index_build_ex(const int64_t* row_ptr, const int32_t* col, const void* cum,
                                         int64_t num_nodes, const void* workspace, uint32_t* meta,
                                         uint32_t* idx, uint32_t* leaf, int leaf_format, pb200_stream_t stream) {
    PB_REQUIRE(num_nodes >= 0, "walk_index_build: bad arguments");
    PB_REQUIRE(leaf_format == PB200_LEAF_WIDE || leaf_format == PB200_LEAF_COMPACT, "walk_index_build: unknown leaf format");
    PB_REQUIRE(leaf_format == PB200_LEAF_WIDE || num_nodes <= (1 << 24) - 1,
               "walk_index_build: the compact leaf holds 24-bit node ids");
    if (num_nodes == 0) return PB200_OK;
    PB_REQUIRE(row_ptr && workspace && meta, "walk_index_build: null pointer");
    PB_REQUIRE(((uintptr_t)meta % 16 == 0) && ((uintptr_t)idx % 32 == 0) && ((uintptr_t)leaf % 64 == 0),
               "walk_index_build: meta/idx/leaf must be 16/32/64-byte aligned");
    WidxWs w = widx_carve(const_cast<void*>(workspace), num_nodes);
    const unsigned blocks = (unsigned)(ceil_div(num_nodes, 8) < kSMs * 16 ? ceil_div(num_nodes, 8) : kSMs * 16);
    widx_fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        row_ptr, col, static_cast<const uint32_t*>(cum), num_nodes, w.leaf_off, w.idx_off,
        reinterpret_cast<uint4*>(meta), idx, leaf, leaf_format == PB200_LEAF_COMPACT);
    return check_launch("widx_fill_kernel");
}

extern "C" int pb200_walk_index_build(const int64_t* row_ptr, const int32_t* col, const void* cum,
                                      int64_t num_nodes, const void* workspace, uint32_t* meta,
                                      uint32_t* idx, uint32_t* leaf, pb200_stream_t stream) {
    return pb200_walk_index_build_ex(row_ptr, col, cum, num_nodes, workspace, meta, idx, leaf, PB200_LEAF_WIDE,
                                     stream);
}
