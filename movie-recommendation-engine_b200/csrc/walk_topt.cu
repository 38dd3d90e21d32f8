// walk_topt.cu -- K1: random-walk neighbour sampling, visit counting, top-T.
//
// Replaces RandomWalkSampler._single_walk / sample_neighbors / batch_sample_neighbors
// (reference utils/random_walk.py:52-142).  One warp per start node:
//   * lane l runs walks l, l+32, ...; every step draws one 53-bit uniform from
//     Philox4x32-10 (counter = (start, walk, step/2, epoch)) and picks the first edge whose
//     row-local cumulative weight exceeds u*total (integer-exact on quantised weights);
//   * visits go into a per-warp shared-memory hash table (key, count, first-visit index);
//   * the warp selects the top-T by (count desc, first visit asc) -- the reference's
//     Counter + stable sorted() order (random_walk.py:101-107) -- and emits
//     weight = count / sum(kept counts) (random_walk.py:113-115).
// Bound by HBM latency/bandwidth: per step ~ (row_ptr pair + log2(deg) prefix probes + col).
#include <type_traits>

#include "common.cuh"
#include "philox.cuh"

namespace pb200 {

constexpr int kEmpty = -1;

struct WalkParams {
    const int64_t* __restrict__ row_ptr;
    const int32_t* __restrict__ col;
    const void* __restrict__ cum;
    const int32_t* __restrict__ starts;
    const int32_t* __restrict__ trace_in;  // count-only mode: [n, V]
    const uint4* __restrict__ meta;        // indexed mode (walk_index.cu)
    const uint32_t* __restrict__ idx;
    const uint32_t* __restrict__ leaf;
    int64_t n;
    int64_t num_nodes;
    int W, L, T;
    int slots, slot_shift;  // hash table size (power of two) and 32 - log2(slots)
    uint32_t seed_lo, seed_hi, epoch;
    int32_t* __restrict__ out_ids;
    int32_t* __restrict__ out_counts;
    float* __restrict__ out_w;
    int32_t* __restrict__ out_nvalid;
    int32_t* __restrict__ trace_out;
};

__device__ __forceinline__ void table_insert(int32_t* keys, uint32_t* cnt, uint32_t* first,
                                             int slots, int shift, int node, uint32_t fs) {
    uint32_t h = ((uint32_t)node * 2654435761u) >> shift;
    for (;;) {
        const int prev = atomicCAS(&keys[h], kEmpty, node);
        if (prev == kEmpty || prev == node) {
            atomicAdd(&cnt[h], 1u);
            atomicMin(&first[h], fs);
            return;
        }
        h = (h + 1) & (uint32_t)(slots - 1);
    }
}

// first p in [r0, r1) with cum[p] > u * total
__device__ __forceinline__ int64_t pick_edge(const uint32_t* __restrict__ cum, int64_t r0,
                                             int64_t r1, uint64_t k53) {
    const uint64_t total = __ldg(cum + r1 - 1);
    // t = floor(k53 * total / 2^53), exact: the product is < 2^85
    const uint64_t lo64 = k53 * total;
    const uint64_t hi64 = __umul64hi(k53, total);
    const uint64_t t = (hi64 << 11) | (lo64 >> 53);
    int64_t lo = r0, hi = r1 - 1;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)__ldg(cum + mid) > t) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ int64_t pick_edge(const double* __restrict__ cum, int64_t r0,
                                             int64_t r1, uint64_t k53) {
    const double total = __ldg(cum + r1 - 1);
    const double x = ((double)k53 * (1.0 / 9007199254740992.0)) * total;
    int64_t lo = r0, hi = r1 - 1;  // clamp: x may round up to total
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cum + mid) > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ---- indexed step (walk_index.cu): every tree node is one 256-bit load ----
struct U8 { uint32_t v[8]; };

__device__ __forceinline__ U8 ld256_keep(const uint32_t* p) {   // upper levels: keep in L2
    U8 r;
    asm volatile("ld.global.nc.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                   "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ U8 ld256_stream(const uint32_t* p) { // leaves: streamed through L2
    U8 r;
    asm volatile("ld.global.nc.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                   "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t count_le(const U8& k, uint32_t t) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) c += (k.v[i] <= t);
    return c;
}

constexpr int kMaxUpper = 7;   // degree < 8^8

// Returns the next node (>= 0) or -1 at a dead end.  Same rule as pick_edge: first edge whose
// cumulative weight exceeds t = floor(k53 * total / 2^53).
__device__ __forceinline__ int indexed_step(const uint4* __restrict__ meta,
                                            const uint32_t* __restrict__ idx,
                                            const uint32_t* __restrict__ leaf, int cur,
                                            uint64_t k53) {
    const uint4 m = __ldg(meta + cur);   // {leaf block offset, degree, total, idx block offset}
    if (m.y == 0u) return -1;
    const uint64_t total = m.z;
    const uint64_t t64 = (__umul64hi(k53, total) << 11) | ((k53 * total) >> 53);
    const uint32_t t = (uint32_t)t64;
    uint32_t nbl[kMaxUpper + 1];
    nbl[0] = (m.y + 7u) >> 3;
    int L = 0;
#pragma unroll
    for (int l = 1; l <= kMaxUpper; ++l) {
        nbl[l] = (nbl[l - 1] + 7u) >> 3;
        L += (nbl[l - 1] > 1u);
    }
    uint32_t pos = 0, off = m.w;
#pragma unroll
    for (int l = kMaxUpper; l >= 1; --l) {
        if (l <= L) {
            const U8 k = ld256_keep(idx + ((size_t)(off + pos) << 3));
            pos = pos * 8u + count_le(k, t);
            off += nbl[l];
        }
    }
    const uint32_t* blk = leaf + ((size_t)(m.x + pos) << 4);
    const U8 keys = ld256_stream(blk);
    const U8 cols = ld256_stream(blk + 8);
    const uint32_t c = count_le(keys, t);
    uint32_t next = cols.v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) next = (c == (uint32_t)i) ? cols.v[i] : next;
    return (int)next;
}

enum WalkMode { kFlatU32 = 0, kFlatF64 = 1, kCountTrace = 2, kIndexed = 3 };

template <int kMode>
__global__ void __launch_bounds__(256) walk_topt_kernel(const WalkParams p) {
    constexpr bool kCountOnly = kMode == kCountTrace;
    using CumT = typename std::conditional<kMode == kFlatF64, double, uint32_t>::type;
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    int32_t* keys = smem + (size_t)warp * 3 * p.slots;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(keys + p.slots);
    uint32_t* first = cnt + p.slots;
    const int V = p.W * p.L;

    for (int64_t s = (int64_t)blockIdx.x * warps_per_block + warp; s < p.n;
         s += (int64_t)gridDim.x * warps_per_block) {
        for (int i = lane; i < p.slots; i += 32) {
            keys[i] = kEmpty; cnt[i] = 0u; first[i] = 0xFFFFFFFFu;
        }
        __syncwarp();

        if (kCountOnly) {
            const int32_t* tr = p.trace_in + s * V;
            for (int i = lane; i < V; i += 32) {
                const int v = tr[i];
                if (v >= 0) table_insert(keys, cnt, first, p.slots, p.slot_shift, v, (uint32_t)i);
            }
        } else {
            const CumT* __restrict__ cum = static_cast<const CumT*>(p.cum);
            const int start = p.starts[s];
            int32_t* tr = p.trace_out ? p.trace_out + s * V : nullptr;
            for (int wbase = 0; wbase < p.W; wbase += 32) {
                const int walk = wbase + lane;
                bool active = walk < p.W;
                int cur = start;
                Philox4 r;
                for (int l = 0; l < p.L; ++l) {
                    if (active && (l & 1) == 0)
                        r = philox4x32_10((uint32_t)start, (uint32_t)walk, (uint32_t)(l >> 1),
                                          p.epoch, p.seed_lo, p.seed_hi);
                    int next = -1;
                    if (active) {
                        const uint64_t k53 = (l & 1) ? uniform53(r.v[2], r.v[3])
                                                     : uniform53(r.v[0], r.v[1]);
                        if (kMode == kIndexed) {
                            next = indexed_step(p.meta, p.idx, p.leaf, cur, k53);
                        } else {
                            const int64_t r0 = __ldg(p.row_ptr + cur);
                            const int64_t r1 = __ldg(p.row_ptr + cur + 1);
                            if (r1 != r0) next = __ldg(p.col + pick_edge(cum, r0, r1, k53));
                        }
                        if (next < 0) {
                            active = false;  // dead end: random_walk.py:68-69
                        } else {
                            cur = next;
                            table_insert(keys, cnt, first, p.slots, p.slot_shift, next,
                                         (uint32_t)(walk * p.L + l));
                        }
                    }
                    if (tr && walk < p.W) tr[walk * p.L + l] = next;
                }
            }
        }
        __syncwarp();

        // sort key: count in the high half, (0xFFFF - first visit) in the low half; unique
        // per node because first-visit indices are unique.
        for (int i = lane; i < p.slots; i += 32) {
            const uint32_t c = cnt[i];
            cnt[i] = c ? ((c << 16) | (0xFFFFu - first[i])) : 0u;
        }
        // slot i is only ever touched by lane (i & 31) from here on: no sync needed
        int32_t* o_ids = p.out_ids + s * p.T;
        int32_t* o_cnt = p.out_counts + s * p.T;
        float* o_w = p.out_w + s * p.T;
        uint32_t total = 0;
        int nvalid = 0;
        for (int j = 0; j < p.T; ++j) {
            uint32_t best = 0;
            int bslot = -1;
            for (int i = lane; i < p.slots; i += 32) {
                const uint32_t v = cnt[i];
                if (v > best) { best = v; bslot = i; }
            }
            const uint32_t m = __reduce_max_sync(kFull, best);
            if (m == 0) break;  // fewer than T distinct nodes (warp-uniform)
            const int src = __ffs(__ballot_sync(kFull, best == m)) - 1;
            int node = 0;
            if (lane == src) { node = keys[bslot]; cnt[bslot] = 0u; }
            node = __shfl_sync(kFull, node, src);
            const uint32_t c = m >> 16;
            total += c;
            if (lane == (j & 31)) { o_ids[j] = node; o_cnt[j] = (int32_t)c; }
            ++nvalid;
        }
        // weights (float64 division like the reference, then the fp32 cast that
        // torch.tensor(list) applies in ImportancePooling, model/pinsage.py:140)
        for (int j = lane; j < p.T; j += 32) {
            if (j < nvalid) {
                o_w[j] = (float)((double)o_cnt[j] / (double)total);  // own earlier store
            } else {
                o_ids[j] = -1; o_cnt[j] = 0; o_w[j] = 0.0f;
            }
        }
        if (lane == 0) p.out_nvalid[s] = nvalid;
        __syncwarp();
    }
}

static int launch_walk(WalkParams& p, int cum_kind, bool count_only, cudaStream_t stream) {
    const int V = p.W * p.L;
    int slots = 64;
    while (slots < V + V / 4 + 1) slots <<= 1;
    p.slots = slots;
    int lg = 0;
    while ((1 << lg) < slots) ++lg;
    p.slot_shift = 32 - lg;
    const size_t per_warp = (size_t)slots * 3 * sizeof(int32_t);
    int warps = 8;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    if (per_warp * warps > 200 * 1024) {
        set_error("walk_topt: num_walks*walk_length=%d needs %zu B of shared memory per warp", V,
                  per_warp);
        return PB200_ERR_UNSUPPORTED;
    }
    const size_t smem = per_warp * warps;
    auto kern = count_only ? walk_topt_kernel<kCountTrace>
              : p.meta ? walk_topt_kernel<kIndexed>
              : cum_kind == 0 ? walk_topt_kernel<kFlatU32> : walk_topt_kernel<kFlatF64>;
    if (smem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // enough blocks to fill every SM several times; grid-stride over the rest
    int64_t blocks = ceil_div(p.n, warps);
    const int64_t cap = (int64_t)kSMs * 32;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(p);
    return check_launch("walk_topt_kernel");
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_walk_topt(const int64_t* row_ptr, const int32_t* col, const void* cum,
                               int cum_kind, int64_t num_nodes, const int32_t* starts, int64_t n,
                               int num_walks, int walk_length, int num_neighbors, uint64_t seed,
                               uint32_t epoch, int32_t* out_ids, int32_t* out_counts,
                               float* out_weights, int32_t* out_nvalid, int32_t* trace_out,
                               pb200_stream_t stream) {
    PB_REQUIRE(n >= 0 && num_walks > 0 && walk_length > 0 && num_neighbors > 0,
               "walk_topt: n=%lld W=%d L=%d T=%d must be positive", (long long)n, num_walks,
               walk_length, num_neighbors);
    PB_REQUIRE((int64_t)num_walks * walk_length <= 65535,
               "walk_topt: num_walks*walk_length must be <= 65535");
    PB_REQUIRE(cum_kind == 0 || cum_kind == 1, "walk_topt: cum_kind must be 0 (u32) or 1 (f64)");
    if (n == 0) return PB200_OK;
    PB_REQUIRE(row_ptr && col && cum && starts && out_ids && out_counts && out_weights &&
               out_nvalid, "walk_topt: null pointer");
    WalkParams p{};
    p.row_ptr = row_ptr; p.col = col; p.cum = cum; p.starts = starts; p.trace_in = nullptr;
    p.n = n; p.num_nodes = num_nodes; p.W = num_walks; p.L = walk_length; p.T = num_neighbors;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.epoch = epoch;
    p.out_ids = out_ids; p.out_counts = out_counts; p.out_w = out_weights;
    p.out_nvalid = out_nvalid; p.trace_out = trace_out;
    return launch_walk(p, cum_kind, false, (cudaStream_t)stream);
}

extern "C" int pb200_count_topt(const int32_t* trace, int64_t n, int visits_per_start,
                                int num_neighbors, int32_t* out_ids, int32_t* out_counts,
                                float* out_weights, int32_t* out_nvalid, pb200_stream_t stream) {
    PB_REQUIRE(n >= 0 && visits_per_start > 0 && visits_per_start <= 65535 && num_neighbors > 0,
               "count_topt: bad sizes");
    if (n == 0) return PB200_OK;
    PB_REQUIRE(trace && out_ids && out_counts && out_weights && out_nvalid,
               "count_topt: null pointer");
    WalkParams p{};
    p.trace_in = trace; p.n = n; p.W = visits_per_start; p.L = 1; p.T = num_neighbors;
    p.out_ids = out_ids; p.out_counts = out_counts; p.out_w = out_weights;
    p.out_nvalid = out_nvalid;
    return launch_walk(p, 0, true, (cudaStream_t)stream);
}

extern "C" int pb200_walk_topt_indexed(const uint32_t* meta, const uint32_t* idx,
                                       const uint32_t* leaf, int64_t num_nodes,
                                       const int32_t* starts, int64_t n, int num_walks,
                                       int walk_length, int num_neighbors, uint64_t seed,
                                       uint32_t epoch, int32_t* out_ids, int32_t* out_counts,
                                       float* out_weights, int32_t* out_nvalid, int32_t* trace_out,
                                       pb200_stream_t stream) {
    PB_REQUIRE(n >= 0 && num_walks > 0 && walk_length > 0 && num_neighbors > 0,
               "walk_topt_indexed: n=%lld W=%d L=%d T=%d must be positive", (long long)n,
               num_walks, walk_length, num_neighbors);
    PB_REQUIRE((int64_t)num_walks * walk_length <= 65535,
               "walk_topt_indexed: num_walks*walk_length must be <= 65535");
    if (n == 0) return PB200_OK;
    PB_REQUIRE(meta && leaf && starts && out_ids && out_counts && out_weights && out_nvalid,
               "walk_topt_indexed: null pointer");
    WalkParams p{};
    p.meta = reinterpret_cast<const uint4*>(meta); p.idx = idx; p.leaf = leaf; p.starts = starts;
    p.n = n; p.num_nodes = num_nodes; p.W = num_walks; p.L = walk_length; p.T = num_neighbors;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.epoch = epoch;
    p.out_ids = out_ids; p.out_counts = out_counts; p.out_w = out_weights;
    p.out_nvalid = out_nvalid; p.trace_out = trace_out;
    return launch_walk(p, 0, false, (cudaStream_t)stream);
}
