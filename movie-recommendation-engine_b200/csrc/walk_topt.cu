// walk_topt.cu -- K1: random-walk neighbour sampling, visit counting, top-T.
//
// Replaces RandomWalkSampler._single_walk / sample_neighbors / batch_sample_neighbors
// (reference utils/random_walk.py:52-142).  One warp per start node:
//   * lane l runs walks l, l+32, ...; every step draws one 53-bit uniform from
//     Philox4x32-10 (counter = (start, walk, step/2, epoch)) and picks the first edge whose
//     row-local cumulative weight exceeds u*total (integer-exact on quantised weights);
//   * visits go into a per-start shared-memory hash table (key, count | first-visit index);
//   * the warp selects the top-T by (count desc, first visit asc) -- the reference's
//     Counter + stable sorted() order (random_walk.py:101-107) -- and emits
//     weight = count / sum(kept counts) (random_walk.py:113-115).
// Bound by HBM latency/bandwidth: per step ~ (row_ptr pair + log2(deg) prefix probes + col).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "philox.cuh"

namespace pb200 {

constexpr int kEmpty = -1;
#ifndef PB200_WALK_DEFAULT_VARIANT
#define PB200_WALK_DEFAULT_VARIANT 102   /* binary search + register select + 6 blocks/SM */
#endif

struct WalkParams {
    const int64_t* __restrict__ row_ptr;
    const int32_t* __restrict__ col;
    const void* __restrict__ cum;
    const int32_t* __restrict__ starts;
    const int32_t* __restrict__ trace_in;  // count-only mode: [n, V]
    const uint4* __restrict__ meta;        // indexed mode (walk_index.cu)
    const uint32_t* __restrict__ idx;
    const uint32_t* __restrict__ leaf;
    int leaf_compact;                       // PB200_LEAF_COMPACT: 32-byte leaf blocks
    int leaf_format;
    int64_t n;              // work items = start nodes x epochs
    int64_t n_starts;       // start nodes per epoch (fast kernels: item s = (epoch s / n_starts, start s % n_starts))
    int64_t num_nodes;
    int W, L, T;
    int slots, slot_shift;  // hash table size (power of two) and 32 - log2(slots)
    uint32_t seed_lo, seed_hi, epoch;
    const uint32_t* __restrict__ epoch_dev;   // optional: added to epoch (CUDA-graph replays advance it on the device)
    int32_t* __restrict__ out_ids;
    int32_t* __restrict__ out_counts;
    float* __restrict__ out_w;
    int32_t* __restrict__ out_nvalid;
    int32_t* __restrict__ trace_out;
};

// Per-start hash table: keys[slots] (node id, -1 = empty), cnt[slots], first[slots].  Plain
// hardware atomics: many lanes of a warp hit the same popular node in the same instruction,
// which atomicAdd/atomicMin resolve in the LSU (a CAS loop on a packed word serialises).
__device__ __forceinline__ void table_insert(int32_t* keys, int slots, int shift, int node,
                                             uint32_t fs) {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(keys + slots);
    uint32_t* first = cnt + slots;
    uint32_t h = ((uint32_t)node * 2654435761u) >> shift;
    for (;;) {
        const int prev = atomicCAS(&keys[h], kEmpty, node);
        if (prev == kEmpty || prev == node) break;
        h = (h + 1) & (uint32_t)(slots - 1);
    }
    atomicAdd(&cnt[h], 1u);
    atomicMin(&first[h], fs);
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
// number of keys <= t in a sorted node.  The searched node always holds a key > t, so the
// answer is in [0, 7]: kBin uses three dependent compares, otherwise eight independent ones.
template <bool kBin>
__device__ __forceinline__ uint32_t count_le(const U8& k, uint32_t t) {
    if (kBin) {
        const bool m1 = k.v[3] <= t;
        const uint32_t p2 = m1 ? k.v[5] : k.v[1];
        const bool m2 = p2 <= t;
        const uint32_t lo = m2 ? k.v[2] : k.v[0], hi = m2 ? k.v[6] : k.v[4];
        const bool m3 = (m1 ? hi : lo) <= t;
        return (m1 ? 4u : 0u) + (m2 ? 2u : 0u) + (m3 ? 1u : 0u);
    }
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) c += (k.v[i] <= t);
    return c;
}


// Returns the next node (>= 0) or -1 at a dead end.  Same rule as pick_edge: first edge whose
// cumulative weight exceeds t = floor(k53 * total / 2^53).
template <bool kBin, bool kCompact>
__device__ __forceinline__ int indexed_step(const uint4* __restrict__ meta,
                                            const uint32_t* __restrict__ idx,
                                            const uint32_t* __restrict__ leaf, int cur,
                                            uint64_t k53) {
    const uint4 m = __ldg(meta + cur);   // {leaf block offset, degree, total, idx block offset}
    if (m.y == 0u) return -1;
    const uint64_t total = m.z;
    const uint64_t t64 = (__umul64hi(k53, total) << 11) | ((k53 * total) >> 53);
    const uint32_t t = (uint32_t)t64;
    // upper levels: L = ceil(log8(leaf blocks)); level l holds ceil(nb0 / 8^l) blocks, stored
    // top level first.  A real loop (the trip count is 1-2 for typical degrees).
    const uint32_t nb0 = (m.y + 7u) >> 3;
    uint32_t pos = 0;
    uint32_t sep = m.z;                  // compact leaf: last cumulative weight of the chosen block
    if (nb0 > 1u) {
        int l = (((31 - __clz(nb0 - 1u)) * 11) >> 5) + 1;      // floor(log2(nb0-1)) / 3 + 1
        uint32_t off = m.w;
#pragma unroll 1
        for (; l >= 1; --l) {
            const U8 k = ld256_keep(idx + ((size_t)(off + pos) << 3));
            const uint32_t c = count_le<kBin>(k, t);
            pos = pos * 8u + c;
            if (kCompact) {              // the key that selected the child = its separator
                sep = k.v[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) sep = (c == (uint32_t)i) ? k.v[i] : sep;
            }
            off += (nb0 + (1u << (3 * l)) - 1u) >> (3 * l);
        }
    }
    if (kCompact) {
        // 32-byte block: v[0..1] = 8 x u8 (sep - cum), v[2..5] = 8 x u16 id low, v[6..7] = 8 x u8 id high.
        // #(cum_i <= t) = #(sep - cum_i >= sep - t); sep > t, so the threshold is >= 1 and the zero
        // padding never counts; deltas are <= 255 by construction.
        const U8 w = ld256_stream(leaf + ((size_t)(m.x + pos) << 3));
        const uint32_t dt = sep - t;
        uint32_t c = 0u;
        if (dt <= 255u) {
            const uint32_t d4 = dt * 0x01010101u;
            c = (uint32_t)(__popc(__vcmpgeu4(w.v[0], d4)) + __popc(__vcmpgeu4(w.v[1], d4))) >> 3;
        }
        uint32_t lw = w.v[2];
        lw = (c >> 1) == 1u ? w.v[3] : lw; lw = (c >> 1) == 2u ? w.v[4] : lw; lw = (c >> 1) == 3u ? w.v[5] : lw;
        const uint32_t hw = (c >> 2) ? w.v[7] : w.v[6];
        return (int)(((lw >> ((c & 1u) * 16u)) & 0xFFFFu) | (((hw >> ((c & 3u) * 8u)) & 0xFFu) << 16));
    }
    const uint32_t* blk = leaf + ((size_t)(m.x + pos) << 4);
    const U8 keys = ld256_stream(blk);
    const U8 cols = ld256_stream(blk + 8);
    const uint32_t c = count_le<kBin>(keys, t);
    uint32_t next = cols.v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) next = (c == (uint32_t)i) ? cols.v[i] : next;
    return (int)next;
}


// ---- bucket step (walk_bucket.cu): meta -> ONE 32-byte bucket, two dependent loads ----
// (a, b) are the two Philox words of this step: k53 = (a >> 5) * 2^26 + (b >> 6) (uniform53),
// t = floor(k53 * S / 2^53) = (kh * S + mulhi(kl, S)) >> 21 with kh = k53 >> 32, kl = low word
// (floor(floor(x / 2^32) / 2^21) = floor(x / 2^53); kh * S * 2^32 is a multiple of 2^32).
// kLd: cache behaviour of the two divergent loads.  0: ld.global.nc (+ L2 evict_first on the bucket);
// 1: + L1::no_allocate; 2: ld.global.cg (L2 only); 3: L1::evict_first.  (L1 has no reuse to offer
// here -- 3 % hit rate -- but every outstanding divergent load pins one of its 128-byte lines.)
template <int kLd>
__device__ __forceinline__ U8 ld256_bucket(const uint32_t* p) {
    U8 r;
    if (kLd == 1)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                       "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    else if (kLd == 2)
        asm volatile("ld.global.cg.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                       "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    else if (kLd == 3)
        asm volatile("ld.global.nc.L1::evict_first.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                       "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    else
        r = ld256_stream(p);
    return r;
}
template <int kLd>
__device__ __forceinline__ uint4 ld_meta(const uint4* p) {
    uint4 r;
    if (kLd == 1)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (kLd == 2)
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (kLd == 3)
        asm volatile("ld.global.nc.L1::evict_first.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else
        r = __ldg(p);
    return r;
}

template <bool k32> __device__ __forceinline__ int bucket_pick(const U8& w, uint32_t tr);
template <int kLd = 0, bool k32 = false>
__device__ __forceinline__ int bucket_step(const uint4* __restrict__ meta,
                                           const uint32_t* __restrict__ leaf, int cur, uint32_t a,
                                           uint32_t b) {
    const uint4 m = ld_meta<kLd>(meta + cur);   // {first bucket, degree, row total S, shift s}
    if (m.y == 0u) return -1;
    const uint32_t kh = a >> 11;
    const uint32_t kl = ((a >> 5) << 26) | (b >> 6);
    const uint64_t prod = (uint64_t)kh * m.z + (uint64_t)__umulhi(kl, m.z);
    const uint32_t t = (uint32_t)(prod >> 21);
    const uint32_t j = t >> m.w;
    const uint32_t tr = t - (j << m.w);
    const U8 w = ld256_bucket<kLd>(leaf + ((size_t)(m.x + j) << 3));
    // slots with rel <= tr, bytewise: 0x80 + tr - rel keeps bit 7 iff rel <= tr (rel <= 128, tr <= 127:
    // no borrow crosses a byte).  rel is ascending and the edge holding t is in the bucket: c <= 7.
    return bucket_pick<k32>(w, tr);
}

// Predicated forms for the batched kernel: kB independent walks per lane are kept in flight, their
// loads issued back to back (no branch may separate them), so a lane that has nothing to do must
// not fetch.  mode: 0 = ld.global.nc, 2 = ld.global.cg (L2 only).
template <int kLd>
__device__ __forceinline__ uint4 ld_meta_if(const uint4* p, bool on) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (kLd == 2)
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4]; }"
                     : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"((uint32_t)on));
    else
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4]; }"
                     : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"((uint32_t)on));
    return r;
}
template <int kLd>
__device__ __forceinline__ U8 ld256_bucket_if(const uint32_t* p, bool on) {
    U8 r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0u;
    if (kLd == 2)
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %9, 0; @q ld.global.cg.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8]; }"
                     : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]),
                       "+r"(r.v[6]), "+r"(r.v[7]) : "l"(p), "r"((uint32_t)on));
    else
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %9, 0; @q ld.global.nc.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8]; }"
                     : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]),
                       "+r"(r.v[6]), "+r"(r.v[7]) : "l"(p), "r"((uint32_t)on));
    return r;
}
// bucket_step in two halves: address of the bucket + in-bucket threshold, then the pick
__device__ __forceinline__ uint32_t bucket_locate(const uint4& m, uint32_t a, uint32_t b, uint32_t& tr) {
    const uint32_t kh = a >> 11;
    const uint32_t kl = ((a >> 5) << 26) | (b >> 6);
    const uint64_t prod = (uint64_t)kh * m.z + (uint64_t)__umulhi(kl, m.z);
    const uint32_t t = (uint32_t)(prod >> 21);
    const uint32_t j = t >> m.w;
    tr = t - (j << m.w);
    return m.x + j;
}
// k32 = PB200_LEAF_BUCKET32: six slots, rel bytes 0..5 (bytes 6..7 hold 128 and are masked out), ids as words 2..7
template <bool k32>
__device__ __forceinline__ int bucket_pick(const U8& w, uint32_t tr) {
    const uint32_t t4 = tr * 0x01010101u + 0x80808080u;
    if (k32) {
        const uint32_t c = (uint32_t)(__popc((t4 - w.v[0]) & 0x80808080u) + __popc((t4 - w.v[1]) & 0x00008080u));
        const bool odd = c & 1u;                                   // c <= 5: the edge holding t is in the bucket
        const uint32_t lo = odd ? w.v[3] : w.v[2], mid = odd ? w.v[5] : w.v[4], hi = odd ? w.v[7] : w.v[6];
        return (int)(c >= 4u ? hi : (c >= 2u ? mid : lo));
    }
    const uint32_t c = (uint32_t)(__popc((t4 - w.v[0]) & 0x80808080u) + __popc((t4 - w.v[1]) & 0x80808080u));
    const uint32_t r0 = __byte_perm(w.v[2], w.v[3], c);
    const uint32_t r1 = __byte_perm(w.v[4], w.v[5], c);
    const uint32_t r2 = __byte_perm(w.v[6], w.v[7], c);
    return (int)(__byte_perm(__byte_perm(r0, r1, 0x0040u), r2, 0x0410u) & 0x00FFFFFFu);
}

enum WalkMode { kFlatU32 = 0, kFlatF64 = 1, kCountTrace = 2, kIndexed = 3, kIndexedCompact = 4, kIndexedBucket = 5, kIndexedBucket32 = 6 };

// 8 sorted values per lane (descending) -- Batcher odd-even merge sort network, 19 exchanges
__device__ __forceinline__ void cex(uint32_t& a, uint32_t& b) {   // a >= b afterwards
    const uint32_t hi = max(a, b), lo = min(a, b);
    a = hi; b = lo;
}
__device__ __forceinline__ void sort8_desc(uint32_t (&v)[8]) {
    cex(v[0], v[1]); cex(v[2], v[3]); cex(v[4], v[5]); cex(v[6], v[7]);
    cex(v[0], v[2]); cex(v[1], v[3]); cex(v[4], v[6]); cex(v[5], v[7]);
    cex(v[1], v[2]); cex(v[5], v[6]);
    cex(v[0], v[4]); cex(v[1], v[5]); cex(v[2], v[6]); cex(v[3], v[7]);
    cex(v[2], v[4]); cex(v[3], v[5]);
    cex(v[1], v[2]); cex(v[3], v[4]); cex(v[5], v[6]);
}

// Register top-T (slots <= 512, W*L <= 255, T <= 32): key = count<<17 | (255-first)<<9 | slot, so
// the winner's node id is one broadcast shared-memory load; keys are sorted once per lane in
// registers and every round pops the warp-wide maximum.  Lane j keeps the j-th winner: the
// results leave the warp as three coalesced stores.
template <bool kWide>
__device__ __forceinline__ void select_topt_regs(const WalkParams& p, const int32_t* keys, int64_t s,
                                                 int lane) {
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(keys + p.slots);
    const uint32_t* first = cnt + p.slots;
    uint32_t k[8], k2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int slot = i * 32 + lane;
        const uint32_t c = cnt[slot];
        k[i] = c ? ((c << 17) | ((255u - first[slot]) << 9) | (uint32_t)slot) : 0u;
        k2[i] = 0u;
        if (kWide) {
            const uint32_t c2 = cnt[slot + 256];
            k2[i] = c2 ? ((c2 << 17) | ((255u - first[slot + 256]) << 9) | (uint32_t)(slot + 256)) : 0u;
        }
    }
    sort8_desc(k);
    if (kWide) sort8_desc(k2);
    int my_id = -1; uint32_t my_cnt = 0, total = 0;
    int nvalid = 0;
    for (int j = 0; j < p.T; ++j) {
        const uint32_t head = kWide ? max(k[0], k2[0]) : k[0];
        const uint32_t m = __reduce_max_sync(kFull, head);
        if (m == 0) break;                     // fewer than T distinct nodes
        if (k[0] == m) {                       // unique winner pops its head
#pragma unroll
            for (int i = 0; i < 7; ++i) k[i] = k[i + 1];
            k[7] = 0u;
        } else if (kWide && k2[0] == m) {
#pragma unroll
            for (int i = 0; i < 7; ++i) k2[i] = k2[i + 1];
            k2[7] = 0u;
        }
        const uint32_t c = m >> 17;
        const int node = keys[m & 511u];       // broadcast
        total += c;
        my_id = lane == j ? node : my_id;
        my_cnt = lane == j ? c : my_cnt;
        ++nvalid;
    }
    if (lane < p.T) {
        const bool has = lane < nvalid;
        const int64_t o = s * p.T + lane;
        p.out_ids[o] = has ? my_id : -1;
        p.out_counts[o] = has ? (int32_t)my_cnt : 0;
        // float64 division like the reference, then the fp32 cast that
        // torch.tensor(list) applies in ImportancePooling (model/pinsage.py:140)
        p.out_w[o] = has ? (float)((double)my_cnt / (double)total) : 0.0f;
    }
    if (lane == 0) p.out_nvalid[s] = nvalid;
}

// Warp-level structure: one warp per start node (warps never wait for each other).  Lane l
// runs walks l, l+32, ...: the start node, its table and its row are warp-uniform, so the
// first step's loads are broadcasts.  (Measured alternatives, tools/tune_walk.py: two starts
// per warp and block-wide flattening of the (start, walk) pairs are both slower.)
template <int kMode, bool kBin, bool kRegSel, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) walk_topt_kernel(const WalkParams p) {
    constexpr bool kCountOnly = kMode == kCountTrace;
    using CumT = typename std::conditional<kMode == kFlatF64, double, uint32_t>::type;
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int kWarpsPerBlock = blockDim.x >> 5;
    const int V = p.W * p.L;
    int32_t* keys = smem + (size_t)warp * 3 * p.slots;   // [keys | cnt | first]
    const CumT* __restrict__ cum = static_cast<const CumT*>(p.cum);

    for (int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + warp; s < p.n;
         s += (int64_t)gridDim.x * kWarpsPerBlock) {
        for (int i = lane; i < p.slots; i += 32) {
            keys[i] = kEmpty; keys[p.slots + i] = 0; keys[2 * p.slots + i] = -1;
        }
        __syncwarp();

        if (kCountOnly) {
            const int32_t* tr = p.trace_in + s * V;
            for (int i = lane; i < V; i += 32) {
                const int v = tr[i];
                if (v >= 0) table_insert(keys, p.slots, p.slot_shift, v, (uint32_t)i);
            }
        } else {
            const int start = p.starts[s];
            const uint32_t epoch = p.epoch + (p.epoch_dev ? __ldg(p.epoch_dev) : 0u);
            for (int walk = lane; walk < p.W; walk += 32) {
                int32_t* tr = p.trace_out ? p.trace_out + (s * p.W + walk) * p.L : nullptr;
                int cur = start;
                Philox4 r;
                int l = 0;
                for (; l < p.L; ++l) {
                    if ((l & 1) == 0)
                        r = philox4x32_10((uint32_t)start, (uint32_t)walk, (uint32_t)(l >> 1),
                                          epoch, p.seed_lo, p.seed_hi);
                    const uint64_t k53 = (l & 1) ? uniform53(r.v[2], r.v[3])
                                                 : uniform53(r.v[0], r.v[1]);
                    int next = -1;
                    if (kMode == kIndexedBucket || kMode == kIndexedBucket32) {
                        next = (l & 1) ? bucket_step<0, kMode == kIndexedBucket32>(p.meta, p.leaf, cur, r.v[2], r.v[3])
                                       : bucket_step<0, kMode == kIndexedBucket32>(p.meta, p.leaf, cur, r.v[0], r.v[1]);
                    } else if (kMode == kIndexed || kMode == kIndexedCompact) {
                        next = indexed_step<kBin, kMode == kIndexedCompact>(p.meta, p.idx, p.leaf, cur, k53);
                    } else {
                        const int64_t r0 = __ldg(p.row_ptr + cur);
                        const int64_t r1 = __ldg(p.row_ptr + cur + 1);
                        if (r1 != r0) next = __ldg(p.col + pick_edge(cum, r0, r1, k53));
                    }
                    if (next < 0) break;   // dead end: random_walk.py:68-69
                    cur = next;
                    table_insert(keys, p.slots, p.slot_shift, next, (uint32_t)(walk * p.L + l));
                    if (tr) tr[l] = next;
                }
                if (tr) for (; l < p.L; ++l) tr[l] = -1;
            }
        }
        __syncwarp();

        // ---- top-T by (count desc, first visit asc) ----
        {
            uint32_t* cnt = reinterpret_cast<uint32_t*>(keys + p.slots);
            uint32_t* first = cnt + p.slots;
            int32_t* o_ids = p.out_ids + s * p.T;
            int32_t* o_cnt = p.out_counts + s * p.T;
            float* o_w = p.out_w + s * p.T;
            uint32_t total = 0;
            int nvalid = 0;
            if (kRegSel && p.slots <= 512 && V <= 255 && p.T <= 32) {
                if (p.slots == 512) select_topt_regs<true>(p, keys, s, lane);
                else select_topt_regs<false>(p, keys, s, lane);
                __syncwarp();
                continue;
            } else {
                // generic path: sort key = count<<16 | (0xFFFF - first), unique per node (first
                // visit indices are unique); T rounds of arg-max; a slot is only touched by lane
                // (slot & 31) from here on
                for (int i = lane; i < p.slots; i += 32) {
                    const uint32_t c = cnt[i];
                    cnt[i] = c ? ((c << 16) | (0xFFFFu - first[i])) : 0u;
                }
                for (int j = 0; j < p.T; ++j) {
                    uint32_t best = 0;
                    int bslot = -1;
                    for (int i = lane; i < p.slots; i += 32) {
                        const uint32_t v = cnt[i];
                        if (v > best) { best = v; bslot = i; }
                    }
                    const uint32_t m = __reduce_max_sync(kFull, best);
                    if (m == 0) break;
                    const int src = __ffs(__ballot_sync(kFull, best == m)) - 1;
                    int node = 0;
                    if (lane == src) { node = keys[bslot]; cnt[bslot] = 0u; }
                    node = __shfl_sync(kFull, node, src);
                    const uint32_t c = m >> 16;
                    total += c;
                    if (lane == (j & 31)) { o_ids[j] = node; o_cnt[j] = (int32_t)c; }
                    ++nvalid;
                }
            }
            // weights: float64 division like the reference, then the fp32 cast that
            // torch.tensor(list) applies in ImportancePooling (model/pinsage.py:140)
            for (int j = lane; j < p.T; j += 32) {
                if (j < nvalid) {
                    o_w[j] = (float)((double)o_cnt[j] / (double)total);  // own earlier store
                } else {
                    o_ids[j] = -1; o_cnt[j] = 0; o_w[j] = 0.0f;
                }
            }
            if (lane == 0) p.out_nvalid[s] = nvalid;
        }
        __syncwarp();
    }
}


// ---- lean kernel for the bucket index (the default path) --------------------------------
// Same algorithm as walk_topt_kernel, rebuilt around what ncu showed on the tree-index version
// (profiles/r1_walk_compact_leaf_ncu_details.txt: issue slots 64 %, 18 of 32 lanes active, 3560 warp
// instructions per start node): (i) bucket_step -- two dependent loads and ~35 instructions per step
// instead of 3-5 loads and ~160; (ii) warp-uniform control flow: every lane runs every round and
// every probe iteration under a predicate, so the warp never splits (the old kernel ran its top-T
// sort twice -- once for the 4 lanes of the last round, once for the other 28 -- and its hash
// atomics once per distinct probe-loop exit); (iii) Philox round keys precomputed on the host;
// (iv) double hashing instead of linear probing; (v) weights by IEEE fp32 division, which equals
// (float)((double)c / total) for c <= total <= 255 (tests/test_walk_host_logic.py checks all pairs).
struct PhiloxKeys { uint32_t k0[10], k1[10]; };

__device__ __forceinline__ Philox4 philox_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[r];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// ---- visit table: 256 slots x {key, count, first}, hardware shared-memory atomics -----------------
// All 32 lanes call; lanes with !alive only take part in the votes.  kIns selects the probe loop:
//   0  vote-terminated loop, bool flag            (first working version, 0.208 ms at C2)
//   3  vote-terminated loop, the slot register doubles as the "pending" flag
//   4  plain divergent loop closed by __syncwarp
// Measured and rejected (tools/tune_walk.py, gpurun_out/r2_tune2.log; tools/ubench_smem.cu): an
// atomics-free table -- lanes visiting the same node grouped by match.any, the group leader inserting
// with plain stores + write-then-verify -- 0.225 ms (256 slots) / 0.330 ms (512 slots: occupancy):
// match.any with 32 distinct values costs 64 cycles of issue per warp and 383 cycles of latency on
// B200, an ATOMS.CAS / ADD / MIN with spread addresses 7 / 3.5 / 2.7 cycles.
template <int kIns>
__device__ __forceinline__ void table_insert_uniform(int32_t* keys, int node, bool alive, uint32_t fs) {
    const uint32_t hh = (uint32_t)node * 2654435761u;
    const uint32_t step = ((hh >> 8) & 0xFEu) | 1u;      // odd: visits every slot of the 2^8 table
    uint32_t slot;
    if (kIns == 0) {
        uint32_t h = hh >> 24;
        bool pending = alive;
        while (__any_sync(kFull, pending)) {
            if (pending) {
                const int prev = atomicCAS(&keys[h], kEmpty, node);
                if (prev == kEmpty || prev == node) pending = false;
                else h = (h + step) & 255u;
            }
        }
        slot = h;
    } else if (kIns == 3) {
        int h = alive ? (int)(hh >> 24) : -1;             // h < 0: nothing (left) to do
        slot = 0;
        while (__any_sync(kFull, h >= 0)) {
            if (h >= 0) {
                const int prev = atomicCAS(&keys[h], kEmpty, node);
                const bool hit = prev == kEmpty || prev == node;
                slot = hit ? (uint32_t)h : slot;
                h = hit ? -1 : (int)(((uint32_t)h + step) & 255u);
            }
        }
    } else {
        uint32_t h = hh >> 24;
        if (alive) {
            for (;;) {
                const int prev = atomicCAS(&keys[h], kEmpty, node);
                if (prev == kEmpty || prev == node) break;
                h = (h + step) & 255u;
            }
        }
        __syncwarp();
        slot = h;
    }
    if (alive) {
        atomicAdd(reinterpret_cast<uint32_t*>(keys) + 256 + slot, 1u);
        atomicMin(reinterpret_cast<uint32_t*>(keys) + 512 + slot, fs);
    }
}

// kB visits per lane in ONE probe loop: the loop runs max (not sum) of the probe counts, and the vote /
// branch overhead is paid once per iteration for all kB items.  Order does not matter: counts are
// atomicAdd, first visits atomicMin.
template <int kB>
__device__ __forceinline__ void table_insert_multi(int32_t* keys, const int (&node)[kB], const bool (&alive)[kB],
                                                   const uint32_t (&fs)[kB]) {
    int h[kB]; uint32_t step[kB], slot[kB];
    int all = -1;
#pragma unroll
    for (int i = 0; i < kB; ++i) {
        const uint32_t hh = (uint32_t)node[i] * 2654435761u;
        step[i] = ((hh >> 8) & 0xFEu) | 1u;
        h[i] = alive[i] ? (int)(hh >> 24) : -1;           // h < 0: nothing (left) to do
        slot[i] = 0;
        all &= h[i];
    }
    while (__any_sync(kFull, all >= 0)) {                 // some h[i] >= 0
        all = -1;
#pragma unroll
        for (int i = 0; i < kB; ++i) {
            if (h[i] >= 0) {
                const int prev = atomicCAS(&keys[h[i]], kEmpty, node[i]);
                const bool hit = prev == kEmpty || prev == node[i];
                slot[i] = hit ? (uint32_t)h[i] : slot[i];
                h[i] = hit ? -1 : (int)(((uint32_t)h[i] + step[i]) & 255u);
            }
            all &= h[i];
        }
    }
#pragma unroll
    for (int i = 0; i < kB; ++i) {
        if (alive[i]) {
            atomicAdd(reinterpret_cast<uint32_t*>(keys) + 256 + slot[i], 1u);
            atomicMin(reinterpret_cast<uint32_t*>(keys) + 512 + slot[i], fs[i]);
        }
    }
}

// Register top-T for the lean kernel's table (see select_topt_regs); fp32 division for the weights.
// kSel 0: the winner of a round shifts its 8 sorted keys down in registers; 1: the sorted keys go back
// to the lane's own (now dead) count words and the lane keeps a head pointer.
template <int kSel>
__device__ __forceinline__ void select_topt_fast(const WalkParams& p, int32_t* keys, int64_t s, int lane) {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(keys) + 256;
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int slot = i * 32 + lane;
        const uint32_t c = cnt[slot];
        k[i] = c ? ((c << 17) | ((255u - cnt[256 + slot]) << 9) | (uint32_t)slot) : 0u;
    }
    sort8_desc(k);
    int my_id = -1; uint32_t my_cnt = 0, total = 0;
    int nvalid = 0;
    if (kSel == 1) {
#pragma unroll
        for (int i = 1; i < 8; ++i) cnt[i * 32 + lane] = k[i];      // only this lane reads these words
        cnt[256 + lane] = 0u;                                       // sentinel after the 8th key
        uint32_t head = k[0];
        uint32_t* nxt = cnt + 32 + lane;
        for (int j = 0; j < p.T; ++j) {
            const uint32_t m = __reduce_max_sync(kFull, head);
            if (m == 0) break;
            if (head == m) { head = *nxt; nxt += 32; }
            const uint32_t c = m >> 17;
            const int node = keys[m & 255u];   // broadcast
            total += c;
            my_id = lane == j ? node : my_id;
            my_cnt = lane == j ? c : my_cnt;
            ++nvalid;
        }
    } else {
        for (int j = 0; j < p.T; ++j) {
            const uint32_t m = __reduce_max_sync(kFull, k[0]);
            if (m == 0) break;                     // warp-uniform: fewer than T distinct nodes
            const bool win = k[0] == m;            // keys are unique (slot bits)
#pragma unroll
            for (int i = 0; i < 7; ++i) k[i] = win ? k[i + 1] : k[i];
            k[7] = win ? 0u : k[7];
            const uint32_t c = m >> 17;
            const int node = keys[m & 255u];       // broadcast
            total += c;
            my_id = lane == j ? node : my_id;
            my_cnt = lane == j ? c : my_cnt;
            ++nvalid;
        }
    }
    if (lane < p.T) {
        const bool has = lane < nvalid;
        const int64_t o = s * p.T + lane;
        p.out_ids[o] = has ? my_id : -1;
        p.out_counts[o] = has ? (int32_t)my_cnt : 0;
        p.out_w[o] = has ? __fdiv_rn((float)my_cnt, (float)total) : 0.0f;
    }
    if (lane == 0) p.out_nvalid[s] = nvalid;
}

// kVar = 100 * kLd (steps after the first) + 10 * kSel + kIns
template <int kL, bool kTrace, int kVar, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) walk_bucket_kernel(const WalkParams p, const PhiloxKeys pk) {
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int32_t* keys = smem + warp * 768;                  // [keys | cnt | first] x 256 slots
    const int L = kL ? kL : p.L;
    const uint32_t epoch0 = p.epoch + (p.epoch_dev ? __ldg(p.epoch_dev) : 0u);
    for (int64_t s = (int64_t)blockIdx.x * 8 + warp; s < p.n; s += (int64_t)gridDim.x * 8) {
        {   // clear: lane owns 8 consecutive slots of each array (two 128-bit stores each)
            int4* k4 = reinterpret_cast<int4*>(keys) + lane * 2;
            const int4 e = make_int4(-1, -1, -1, -1), z = make_int4(0, 0, 0, 0);
            k4[0] = e; k4[1] = e; k4[64] = z; k4[65] = z; k4[128] = e; k4[129] = e;
        }
        __syncwarp();
        const int64_t e_idx = s / p.n_starts;                   // several sampling epochs in one launch
        const int start = __ldg(p.starts + (s - e_idx * p.n_starts));
        const uint32_t epoch = epoch0 + (uint32_t)e_idx;
        for (int base = 0; base < p.W; base += 32) {
            const int walk = base + lane;
            bool alive = walk < p.W;
            int cur = start;
            Philox4 r;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                if ((l & 1) == 0)
                    r = philox_keys((uint32_t)start, (uint32_t)walk, (uint32_t)(l >> 1), epoch, pk);
                const uint32_t a = (l & 1) ? r.v[2] : r.v[0], b = (l & 1) ? r.v[3] : r.v[1];
                int next = -1;
                if (alive) next = l == 0 ? bucket_step<0>(p.meta, p.leaf, cur, a, b)
                                          : bucket_step<kVar / 100>(p.meta, p.leaf, cur, a, b);
                if (kTrace) { if (walk < p.W) p.trace_out[(s * p.W + walk) * L + l] = next; }
                alive = next >= 0;             // dead end: random_walk.py:68-69
                table_insert_uniform<kVar % 10>(keys, next, alive, (uint32_t)(walk * L + l));
                cur = alive ? next : cur;
            }
        }
        __syncwarp();
        select_topt_fast<(kVar / 10) % 10>(p, keys, s, lane);
        __syncwarp();
    }
}

// Batched form: kB walks per lane in flight (walks lane, lane + 32, ...).  The one-walk-per-lane kernel
// above is LATENCY bound once the step is this short (r2 measurements, tools/tune_walk.py: time =
// 23 us + 54 us per round of 32 walks, the 4-lane tail round of W = 100 alone 21 us; more resident
// warps make it slower, not faster), so the independent walks of a start node are overlapped inside
// the lane instead: all meta loads, then all bucket loads, then the picks and the table inserts.
template <int kL, bool kTrace, int kLd, int kB, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) walk_bucket_batched_kernel(const WalkParams p, const PhiloxKeys pk) {
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int32_t* keys = smem + warp * 768;                  // [keys | cnt | first] x 256 slots
    const int L = kL ? kL : p.L;
    const uint32_t epoch0 = p.epoch + (p.epoch_dev ? __ldg(p.epoch_dev) : 0u);
    for (int64_t s = (int64_t)blockIdx.x * 8 + warp; s < p.n; s += (int64_t)gridDim.x * 8) {
        {
            int4* k4 = reinterpret_cast<int4*>(keys) + lane * 2;
            const int4 e = make_int4(-1, -1, -1, -1), z = make_int4(0, 0, 0, 0);
            k4[0] = e; k4[1] = e; k4[64] = z; k4[65] = z; k4[128] = e; k4[129] = e;
        }
        __syncwarp();
        const int64_t e_idx = s / p.n_starts;                   // several sampling epochs in one launch
        const int start = __ldg(p.starts + (s - e_idx * p.n_starts));
        const uint32_t epoch = epoch0 + (uint32_t)e_idx;
        for (int base = 0; base < p.W; base += 32 * kB) {
            int cur[kB]; bool alive[kB]; Philox4 r[kB];
#pragma unroll
            for (int i = 0; i < kB; ++i) { alive[i] = base + 32 * i + lane < p.W; cur[i] = start; }
#pragma unroll
            for (int l = 0; l < L; ++l) {
                if ((l & 1) == 0) {
#pragma unroll
                    for (int i = 0; i < kB; ++i)
                        r[i] = philox_keys((uint32_t)start, (uint32_t)(base + 32 * i + lane), (uint32_t)(l >> 1), epoch, pk);
                }
                uint4 m[kB];
#pragma unroll
                for (int i = 0; i < kB; ++i) m[i] = ld_meta_if<(kLd & 3)>(p.meta + cur[i], alive[i]);
                U8 w[kB]; uint32_t tr[kB];
#pragma unroll
                for (int i = 0; i < kB; ++i) {
                    alive[i] = alive[i] && m[i].y != 0u;       // dead end: random_walk.py:68-69
                    const uint32_t blk = bucket_locate(m[i], (l & 1) ? r[i].v[2] : r[i].v[0], (l & 1) ? r[i].v[3] : r[i].v[1], tr[i]);
                    w[i] = ld256_bucket_if<(kLd & 3)>(p.leaf + ((size_t)blk << 3), alive[i]);
                }
                int next[kB]; uint32_t fs[kB];
#pragma unroll
                for (int i = 0; i < kB; ++i) {
                    const int walk = base + 32 * i + lane;
                    next[i] = alive[i] ? bucket_pick<(kLd & 4) != 0>(w[i], tr[i]) : -1;   // kLd & 4: 32-bit ids
                    fs[i] = (uint32_t)(walk * L + l);
                    if (kTrace) { if (walk < p.W) p.trace_out[(s * p.W + walk) * L + l] = next[i]; }
                    cur[i] = alive[i] ? next[i] : cur[i];
                }
                table_insert_multi<kB>(keys, next, alive, fs);
            }
        }
        __syncwarp();
        select_topt_fast<1>(p, keys, s, lane);
        __syncwarp();
    }
}

static void philox_round_keys(uint32_t seed_lo, uint32_t seed_hi, PhiloxKeys& k) {
    uint32_t k0 = seed_lo, k1 = seed_hi;
    for (int r = 0; r < 10; ++r) { k.k0[r] = k0; k.k1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

template <int kL, bool kTrace, int kVar, int kMinBlocks>
static int launch_bucket_variant(const WalkParams& p, const PhiloxKeys& pk, cudaStream_t stream) {
    int64_t blocks = ceil_div(p.n, (int64_t)8);
    const int64_t cap = (int64_t)kSMs * 32;
    if (blocks > cap) blocks = cap;
    static const int pad = [] { const char* e = getenv("PB200_WALK_PAD"); return e ? atoi(e) : 0; }();   // occupancy experiments
    const size_t smem = 8 * 768 * sizeof(int32_t) + (size_t)pad;
    if (smem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(walk_bucket_kernel<kL, kTrace, kVar, kMinBlocks>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const int carve = [] { const char* e = getenv("PB200_WALK_CARVE"); return e ? atoi(e) : -1; }();
    if (carve >= 0)
        PB_CUDA(cudaFuncSetAttribute(walk_bucket_kernel<kL, kTrace, kVar, kMinBlocks>,
                                     cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    walk_bucket_kernel<kL, kTrace, kVar, kMinBlocks><<<(unsigned)blocks, 256, smem, stream>>>(p, pk);
    return check_launch("walk_bucket_kernel");
}

// W*L <= 200 visits in a 256-slot table (load <= 0.78), first-visit index < 255, T <= 32
static bool bucket_fast_ok(const WalkParams& p) { return p.W * p.L <= 200 && p.T <= 32; }

#ifndef PB200_WALK_TABLE_DEFAULT
#define PB200_WALK_TABLE_DEFAULT 213
#endif
#ifndef PB200_WALK_BATCH_DEFAULT
#define PB200_WALK_BATCH_DEFAULT 2
#endif
static int launch_walk_bucket(const WalkParams& p, cudaStream_t stream) {
    // tuning knobs (tools/tune_walk.py): PB200_WALK_MINBLOCKS = 6 | 7 | 8 (register cap),
    // PB200_WALK_TABLE = 10 * select version + insert version
    static const int minb_env = [] { const char* e = getenv("PB200_WALK_MINBLOCKS"); return e ? atoi(e) : 6; }();
    static const int var = [] { const char* e = getenv("PB200_WALK_TABLE"); return e ? atoi(e) : PB200_WALK_TABLE_DEFAULT; }();
    PhiloxKeys pk;
    philox_round_keys(p.seed_lo, p.seed_hi, pk);
    // PB200_WALK_BATCH = walks per lane in flight (0 = the one-walk kernel); PB200_WALK_LD = 0 | 2
    static const int batch_env = [] { const char* e = getenv("PB200_WALK_BATCH"); return e ? atoi(e) : -1; }();
    static const bool minb_set = getenv("PB200_WALK_MINBLOCKS") != nullptr;
    // Default: 2 walks per lane in flight at 6 blocks / SM; launches that cannot fill the GPU several times
    // (a shard of a multi-GPU step) are latency bound and take 4 walks per lane at 4 blocks / SM instead
    // (r2_tune13: 7,803 starts x 2 samples 80 -> 59 us; 62,423 x 2: 334 vs 336 us).
    const bool small = p.n <= (int64_t)kSMs * 48 * 9;
    const int batch = batch_env >= 0 ? batch_env : (small ? 4 : PB200_WALK_BATCH_DEFAULT);
    const int minb = minb_set ? minb_env : (batch == 4 ? 4 : 6);
    static const int ld = [] { const char* e = getenv("PB200_WALK_LD"); return e ? atoi(e) : 2; }();
    if (batch > 0 || p.leaf_format == PB200_LEAF_BUCKET32) {      // 32-bit ids: the batched kernel only
        const int64_t cap = (int64_t)kSMs * 32;
        int64_t blocks = ceil_div(p.n, (int64_t)8);
        if (blocks > cap) blocks = cap;
        const size_t smem = 8 * 768 * sizeof(int32_t);
#define PB_BK(L_, T_, LD_, B_, M_) do { walk_bucket_batched_kernel<L_, T_, LD_, B_, M_><<<(unsigned)blocks, 256, smem, stream>>>(p, pk); \
                                        return check_launch("walk_bucket_batched_kernel"); } while (0)
#define PB_BB(L_, T_, LD_) do { if (batch == 2 && minb == 5) PB_BK(L_, T_, LD_, 2, 5); if (batch == 2) PB_BK(L_, T_, LD_, 2, 6); \
                                if (batch == 4 && minb == 5) PB_BK(L_, T_, LD_, 4, 5); if (batch == 4 && minb == 3) PB_BK(L_, T_, LD_, 4, 3); \
                                if (batch == 4) PB_BK(L_, T_, LD_, 4, 4); PB_BK(L_, T_, LD_, 1, 6); } while (0)
#define PB_BB32(L_, T_) do { if (batch == 2) PB_BK(L_, T_, 6, 2, 6); if (batch == 4) PB_BK(L_, T_, 6, 4, 4); PB_BK(L_, T_, 6, 1, 6); } while (0)
#define PB_BL(L_, T_) do { if (p.leaf_format == PB200_LEAF_BUCKET32) PB_BB32(L_, T_); if (ld == 0) PB_BB(L_, T_, 0); PB_BB(L_, T_, 2); } while (0)
        if (p.trace_out) { if (p.L == 2) PB_BL(2, true); PB_BL(0, true); }
        if (p.L == 2) PB_BL(2, false);
        PB_BL(0, false);
#undef PB_BL
#undef PB_BB32
#undef PB_BB
#undef PB_BK
    }
    // one walk per lane (PB200_WALK_BATCH=0): the stages measured on the way, kept selectable for tools/tune_walk.py --
    // 0 = first working version, 13 = lean probe loop + shared-memory select, 213 = 13 + ld.global.cg, at 6 or 8 blocks / SM
#define PB_B3(L_, T_, V_) do { if (minb == 8) return launch_bucket_variant<L_, T_, V_, 8>(p, pk, stream); \
                               return launch_bucket_variant<L_, T_, V_, 6>(p, pk, stream); } while (0)
#define PB_B(L_, T_) do { if (var == 13) PB_B3(L_, T_, 13); if (var == 213) PB_B3(L_, T_, 213); PB_B3(L_, T_, 0); } while (0)
    if (p.trace_out) { if (p.L == 2) PB_B(2, true); PB_B(0, true); }
    if (p.L == 2) PB_B(2, false);
    PB_B(0, false);
#undef PB_B
#undef PB_B3
}

// Measured and rejected (B200, config C2, tools/tune_walk.py; baseline 0.360 ms per launch):
//   * 2 or 4 walks per lane with the loads of every stage (meta / index level / leaf) issued
//     back to back: 0.361-0.524 ms.  More requests in flight do not help: ncu shows the LSU
//     data pipe (66 % of peak wavefronts -- every lane of a divergent 32 B load is its own
//     wavefront -- plus 16 % from the shared-memory atomics) and the issue slots (59 %) as the
//     loaded units, not memory latency or DRAM bandwidth (28 %).
//   * staging the start node's row in shared memory (coalesced 16 B copies) and running the
//     first step of all W walks as a binary search there: 0.41-0.50 ms (9 dependent LDS per
//     walk with bank conflicts cost more pipe cycles than the 2-3 index/leaf loads they
//     replace, and the extra 2-4 KB per warp lowers occupancy).
//   * (with the compact leaf, baseline 0.341 ms) an instance specialised for walk_length == 2
//     without trace output -- step loop unrolled, trace bookkeeping compiled out: 0.351 ms.

template <int kMode, bool kBin, bool kRegSel, int kMinBlocks>
static int launch_variant(const WalkParams& p, int warps, size_t smem, cudaStream_t stream) {
    auto kern = walk_topt_kernel<kMode, kBin, kRegSel, kMinBlocks>;
    if (smem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // enough blocks to fill every SM several times; grid-stride over the rest
    int64_t blocks = ceil_div(p.n, (int64_t)warps);
    const int64_t cap = (int64_t)kSMs * 32;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(p);
    return check_launch("walk_topt_kernel");
}

static int launch_walk(WalkParams& p, int cum_kind, bool count_only, cudaStream_t stream) {
    // tuning knob (tools/tune_walk.py): bit1 = binary in-node search, bit2 = register top-T
    // selection, bit3 = hash table twice as large, bits 4.. = min resident blocks per SM
    // (register cap); unset = tuned default
    static const int variant = [] { const char* e = getenv("PB200_WALK_VARIANT"); return e ? atoi(e) : -1; }();
    const int v = variant >= 0 ? variant : PB200_WALK_DEFAULT_VARIANT;
    const int V = p.W * p.L;
    int slots = 256;
    while (slots < V + V / 4 + 1) slots <<= 1;
    if (v & 8) slots <<= 1;
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
    if (count_only) return launch_variant<kCountTrace, false, true, 1>(p, warps, smem, stream);
    if (!p.meta)
        return cum_kind == 0 ? launch_variant<kFlatU32, false, true, 1>(p, warps, smem, stream)
                             : launch_variant<kFlatF64, false, true, 1>(p, warps, smem, stream);
    if (p.leaf_format == PB200_LEAF_BUCKET || p.leaf_format == PB200_LEAF_BUCKET32) {
        if (bucket_fast_ok(p) && !(v & 1)) return launch_walk_bucket(p, stream);
        return p.leaf_format == PB200_LEAF_BUCKET32
                   ? launch_variant<kIndexedBucket32, false, true, 1>(p, warps, smem, stream)
                   : launch_variant<kIndexedBucket, false, true, 1>(p, warps, smem, stream);   // generic sizes
    }
    const bool bin = v & 2, reg = v & 4;
    const int minb = v >> 4;
#define PB_V(B_, R_, M_) return p.leaf_compact ? launch_variant<kIndexedCompact, B_, R_, M_>(p, warps, smem, stream) \
                                              : launch_variant<kIndexed, B_, R_, M_>(p, warps, smem, stream)
#define PB_VM(B_, R_) do { if (minb == 5) PB_V(B_, R_, 5); if (minb == 6) PB_V(B_, R_, 6); \
                           if (minb == 7) PB_V(B_, R_, 7); PB_V(B_, R_, 1); } while (0)
    if (bin) { if (reg) PB_VM(true, true); PB_VM(true, false); }
    if (reg) PB_VM(false, true);
    PB_VM(false, false);
#undef PB_VM
#undef PB_V
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
    p.n = n; p.n_starts = n; p.num_nodes = num_nodes; p.W = num_walks; p.L = walk_length; p.T = num_neighbors;
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

extern "C" int pb200_walk_topt_indexed_ex(const uint32_t* meta, const uint32_t* idx,
                                       const uint32_t* leaf, int leaf_format, int64_t num_nodes,
                                       const int32_t* starts, int64_t n, int num_walks,
                                       int walk_length, int num_neighbors, uint64_t seed,
                                       uint32_t epoch, const uint32_t* epoch_dev, int32_t* out_ids,
                                       int32_t* out_counts, float* out_weights, int32_t* out_nvalid,
                                       int32_t* trace_out, pb200_stream_t stream) {
    PB_REQUIRE(n >= 0 && num_walks > 0 && walk_length > 0 && num_neighbors > 0,
               "walk_topt_indexed: n=%lld W=%d L=%d T=%d must be positive", (long long)n,
               num_walks, walk_length, num_neighbors);
    PB_REQUIRE((int64_t)num_walks * walk_length <= 65535,
               "walk_topt_indexed: num_walks*walk_length must be <= 65535");
    if (n == 0) return PB200_OK;
    PB_REQUIRE(meta && leaf && starts && out_ids && out_counts && out_weights && out_nvalid,
               "walk_topt_indexed: null pointer");
    WalkParams p{};
    PB_REQUIRE(leaf_format == PB200_LEAF_WIDE || leaf_format == PB200_LEAF_COMPACT || leaf_format == PB200_LEAF_BUCKET ||
                   leaf_format == PB200_LEAF_BUCKET32,
               "walk_topt_indexed: unknown leaf format %d", leaf_format);
    p.leaf_format = leaf_format;
    p.meta = reinterpret_cast<const uint4*>(meta); p.idx = idx; p.leaf = leaf; p.starts = starts;
    p.leaf_compact = leaf_format == PB200_LEAF_COMPACT;
    p.n = n; p.n_starts = n; p.num_nodes = num_nodes; p.W = num_walks; p.L = walk_length; p.T = num_neighbors;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.epoch = epoch;
    p.epoch_dev = epoch_dev;
    p.out_ids = out_ids; p.out_counts = out_counts; p.out_w = out_weights;
    p.out_nvalid = out_nvalid; p.trace_out = trace_out;
    return launch_walk(p, 0, false, (cudaStream_t)stream);
}

extern "C" int pb200_walk_topt_indexed_multi(const uint32_t* meta, const uint32_t* idx, const uint32_t* leaf,
                                             int leaf_format, int64_t num_nodes, const int32_t* starts, int64_t n,
                                             int num_walks, int walk_length, int num_neighbors, uint64_t seed,
                                             uint32_t epoch, const uint32_t* epoch_dev, int num_epochs,
                                             int32_t* out_ids, int32_t* out_counts, float* out_weights,
                                             int32_t* out_nvalid, int32_t* trace_out, pb200_stream_t stream) {
    PB_REQUIRE(num_epochs >= 1 && num_epochs <= 64, "walk_topt_indexed_multi: 1 <= num_epochs <= 64");
    WalkParams probe{};
    probe.W = num_walks; probe.L = walk_length; probe.T = num_neighbors;
    static const bool forced_generic = [] { const char* e = getenv("PB200_WALK_VARIANT"); return e && (atoi(e) & 1); }();
    if (num_epochs > 1 && n > 0 && (leaf_format == PB200_LEAF_BUCKET || leaf_format == PB200_LEAF_BUCKET32) && num_walks > 0 && walk_length > 0 &&
        num_neighbors > 0 && bucket_fast_ok(probe) && !forced_generic) {
        PB_REQUIRE(meta && leaf && starts && out_ids && out_counts && out_weights && out_nvalid,
                   "walk_topt_indexed_multi: null pointer");
        WalkParams p{};
        p.leaf_format = leaf_format;
        p.meta = reinterpret_cast<const uint4*>(meta); p.idx = idx; p.leaf = leaf; p.starts = starts;
        p.n = n * num_epochs; p.n_starts = n; p.num_nodes = num_nodes; p.W = num_walks; p.L = walk_length;
        p.T = num_neighbors;
        p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.epoch = epoch; p.epoch_dev = epoch_dev;
        p.out_ids = out_ids; p.out_counts = out_counts; p.out_w = out_weights;
        p.out_nvalid = out_nvalid; p.trace_out = trace_out;
        return launch_walk_bucket(p, (cudaStream_t)stream);       // ONE launch over (epoch, start) pairs
    }
    for (int e = 0; e < num_epochs; ++e) {                         // other formats / sizes: one launch per epoch
        const int64_t o = (int64_t)e * n;
        const int rc = pb200_walk_topt_indexed_ex(
            meta, idx, leaf, leaf_format, num_nodes, starts, n, num_walks, walk_length, num_neighbors, seed,
            epoch + (uint32_t)e, epoch_dev, out_ids ? out_ids + o * num_neighbors : nullptr,
            out_counts ? out_counts + o * num_neighbors : nullptr, out_weights ? out_weights + o * num_neighbors : nullptr,
            out_nvalid ? out_nvalid + o : nullptr,
            trace_out ? trace_out + o * num_walks * walk_length : nullptr, stream);
        if (rc) return rc;
    }
    return PB200_OK;
}

extern "C" int pb200_walk_topt_indexed(const uint32_t* meta, const uint32_t* idx,
                                       const uint32_t* leaf, int64_t num_nodes,
                                       const int32_t* starts, int64_t n, int num_walks,
                                       int walk_length, int num_neighbors, uint64_t seed,
                                       uint32_t epoch, int32_t* out_ids, int32_t* out_counts,
                                       float* out_weights, int32_t* out_nvalid, int32_t* trace_out,
                                       pb200_stream_t stream) {
    return pb200_walk_topt_indexed_ex(meta, idx, leaf, PB200_LEAF_WIDE, num_nodes, starts, n, num_walks, walk_length,
                                      num_neighbors, seed, epoch, nullptr, out_ids, out_counts,
                                      out_weights, out_nvalid, trace_out, stream);
}

namespace pb200 {
__global__ void u32_add_kernel(uint32_t* p, uint32_t delta) { *p += delta; }
}  // namespace pb200

extern "C" int pb200_u32_add(uint32_t* counter, uint32_t delta, pb200_stream_t stream) {
    PB_REQUIRE(counter, "u32_add: null pointer");
    pb200::u32_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
    return check_launch("u32_add_kernel");
}
