// topk_tc.cu -- K4 on the tensor cores: exact search as a tcgen05 (kind::tf32) scoring GEMM fused
// with a streaming per-query shortlist, followed by an exact fp32 re-rank and a certificate.
//
// Replaces the same reference code as topk.cu (E1: utils/evaluation.py:119-130,
// inference.py:112-118; E2: faiss IndexFlatL2.search at utils/nearest_neighbors.py:174-181) and
// returns BITWISE the same result as the fp32 kernel there:
//   1. search_tc_kernel  scores 128-query x 128-item tiles with UMMA 128x128x8 TF32 (operands
//      pre-rounded to TF32, fp32 accumulators in TMEM) and keeps, per query and item split, the
//      ks best items by TF32 score.  Persistent CTAs; per CTA two 128-query tiles stay resident
//      in shared memory (every streamed item tile is used by 256 queries: L2 -> SM traffic is
//      the limit of this kernel), item tiles stream through a TMA ring, two accumulator sets in
//      TMEM (4 x 128 columns) so the MMAs of tile i+1 overlap the scan of tile i.
//      Scan: one thread per query row (tcgen05.ld 32x32b) with the row's current ks-th best
//      score as a threshold in a register; 8 columns x 32 rows are rejected by one max,
//      compare and warp vote; a survivor is inserted into its row's sorted list in shared
//      memory by the whole warp (lane = rank: ballot for the position, shuffle for the
//      shift -- no loop).  Measured on the way (C3, all-item queries): candidates appended to
//      per-row scratch lists + cooperative rank-counting prunes 5.5 ms; per-thread insertion
//      sort in shared memory 11.5 ms (2 of 32 lanes active in the shift loop).
//   2. rerank_kernel     recomputes the shortlisted scores in fp32 with the fp32 kernel's
//      arithmetic (sequential fmaf over d, same norms), selects the top-k under the same
//      (score, id) total order, and CERTIFIES the result: with |tf32 score - fp32 score| <= eps
//      (eps = |q| max|x~ - x| + |q~ - q| max|x~| + accumulation slack, from the measured rounding
//      residuals of the operands: Cauchy-Schwarz, no assumption on the data), an item outside
//      the shortlist scores at most t + eps, t = the weakest shortlisted TF32 score; if the k-th
//      exact score beats that bound, no outside item can enter the top-k.
//   3. queries that fail the certificate (near-ties within eps) are re-run by the fp32 kernel
//      (topk.cu) through a device-side compacted list -- no host synchronisation.
#include <cstdlib>

#include "dense.cuh"
#include "ivf.cuh"
#include "tc_common.cuh"
#include "topk.cuh"

namespace pb200 {
namespace tcs {

using namespace tc;

constexpr int kTileN = 128;
constexpr int kNBytes = kTileN * 128;        // one K chunk (32 fp32) of a 128-item tile: 16 KB
constexpr int kProdWarp = 0, kMmaWarp = 1, kEpiWarp0 = 4;
constexpr int kThreads = (kEpiWarp0 + 8) * 32;   // 384
constexpr int kMaxStages = 6;

struct SearchParams {
    int64_t nq, nx;
    int d, nchunks, qt;              // qt = 128-query tiles per CTA (1 or 2)
    int ks;                          // shortlist length (16 or 32)
    int splits; int64_t split_len;   // item chunks per query group; chunk length (multiple of kTileN)
    int stages, align_slack;
    int scan_append;                 // 1: append + periodic warp-sort prune (ks = 16); 0: sorted insertion per candidate
    int kind;                        // 0: tf32 operands (32 per 128 B chunk row), 1: bf16 (64 per chunk row)
    const float* __restrict__ hx;    // 0.5 |x|^2, padded to a multiple of 128 items (L2 only)
    // IVF (pb200_ivf_search_tc): items are list ordered, every list padded to whole 128-item tiles;
    // a row scans a tile only if its query probes the tile's list
    const int32_t* __restrict__ tile_list;   // [nx / 128] list id of every item tile (NULL: no masking)
    const uint32_t* __restrict__ pmask;      // [nq][4] bit l = query probes list l (nlist <= 128)
    unsigned long long* short_keys;  // [splits][nq][ks]   sorted, 0 = empty; slot = first chunk of a segment
};

// sortable key: larger = better (higher score, then lower item index)
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t col) {
    return ((unsigned long long)f2ord(score) << 32) | (unsigned long long)(0xFFFFFFFFu - col);
}

// optional cycle accounting of one scan warp (make EXTRA=-DPB200_SEARCH_PROFILE)
#ifdef PB200_SEARCH_PROFILE
#define SP_NOW() clock64()
#define SP_ADD(var, t0) (var) += clock64() - (t0)
#else
#define SP_NOW() 0ll
#define SP_ADD(var, t0) do { } while (0)
#endif

enum { kBarFull = 0, kBarEmpty = kMaxStages, kBarAFull = 2 * kMaxStages, kBarAFree, kBarTFull,
       kBarTEmpty = kBarTFull + 2, kNumBars = kBarTEmpty + 2 };

// Candidates of one item column for the rows in `m` (lane = row of the warp): the WARP inserts
// each into its row's sorted list (lane = rank): ballot for the position, shuffle for the shift,
// no loop over ranks.  Two rows per iteration: independent chains hide the LDS / shuffle latency.
// Out of line on purpose: inlined at all 32 call sites of the scan it overflowed the
// instruction cache (ncu: stall_no_instruction dominant, 10.2 ms -> see DESIGN.md).
__device__ __noinline__ float insert_column(unsigned m, float fi, uint32_t klo, unsigned long long* wl,
                                            int ks, int lane, float thr) {
    const bool in = lane < ks;
    while (m) {
        const int s0 = __ffs(m) - 1;
        m &= m - 1;
        const bool two = m != 0u;                 // warp-uniform
        const int s1 = two ? __ffs(m) - 1 : s0;
        m &= m - 1;
        const unsigned long long key0 = ((unsigned long long)f2ord(__shfl_sync(kFull, fi, s0)) << 32) | klo;
        const unsigned long long key1 = ((unsigned long long)f2ord(__shfl_sync(kFull, fi, s1)) << 32) | klo;
        unsigned long long* rl0 = wl + s0 * ks;
        unsigned long long* rl1 = wl + s1 * ks;
        const unsigned long long mine0 = in ? rl0[lane] : 0ull;   // lanes >= ks hold 0
        const unsigned long long mine1 = in ? rl1[lane] : 0ull;
        const int pos0 = __popc(__ballot_sync(kFull, mine0 > key0));
        const int pos1 = __popc(__ballot_sync(kFull, mine1 > key1));
        const unsigned long long up0 = __shfl_up_sync(kFull, mine0, 1);
        const unsigned long long up1 = __shfl_up_sync(kFull, mine1, 1);
        const unsigned long long nv0 = lane < pos0 ? mine0 : (lane == pos0 ? key0 : up0);
        const unsigned long long nv1 = lane < pos1 ? mine1 : (lane == pos1 ? key1 : up1);
        if (lane >= pos0 && in) rl0[lane] = nv0;      // a lane only ever touches slot `lane`
        if (two && lane >= pos1 && in) rl1[lane] = nv1;
        const uint32_t last0 = __shfl_sync(kFull, (uint32_t)(nv0 >> 32), ks - 1);
        const uint32_t last1 = __shfl_sync(kFull, (uint32_t)(nv1 >> 32), ks - 1);
        if (lane == s0 && last0) thr = ord2f(last0);
        if (two && lane == s1 && last1) thr = ord2f(last1);
    }
    return thr;
}

// ---- scan mode 1: append + prune (ks = 16) ----
// A row (= one lane) APPENDS every candidate that beats its threshold to its own buffer -- a predicated
// store and an increment, no cross-lane traffic, eight independent appends per 8-column group instead of a
// ~45-instruction dependent chain per insertion -- and the threshold is only refreshed when a buffer is
// about to overflow: the warp then sorts that row's <= 30 keys (bitonic network over the lanes), keeps the
// best 16 and reads the new threshold off rank 15.  The shortlist is still exactly the ks best TF32 scores of
// the scanned columns: a key that belongs to the final top-ks beat every earlier (laxer) threshold, so it was
// appended, and a prune only drops keys that 16 better ones dominate.
constexpr int kScanCap = 30;          // keys per row buffer (lanes 30, 31 of the sort hold "empty")
constexpr int kScanStride = 31;       // odd row stride (in keys): the per-lane appends spread over the banks
constexpr int kScanTrig = kScanCap - 8;   // a group appends <= 8 keys per row: prune above this fill

// Bitonic sort (descending) of one key per lane.
__device__ __forceinline__ unsigned long long warp_sort_desc(unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long other = __shfl_xor_sync(kFull, key, j);
            const bool take_max = ((lane & k) == 0) == ((lane & j) == 0);     // k == 32: descending everywhere
            const bool gt = key > other;
            key = (gt == take_max) ? key : other;
        }
    }
    return key;
}

// Prunes every row of the warp whose fill exceeds `trig` (trig < 0: every non-empty row): the warp sorts the
// row's keys, keeps the best ks and reads the new threshold off rank ks - 1.  Returns the lane's updated
// threshold; `cnt` is the lane's own fill.
// Measured alternatives (C3, cycles of one scan warp, make EXTRA=-DPB200_SEARCH_PROFILE): this form spends 2.0 M
// of 7.2 M cycles in prunes (~1,100 per row: a dependent chain of 30 shuffles); four rows per pass with
// interleaved chains + pruning every row above a lower fill: ~700 per row but 1.6 x the rows, same 2.0 M, and the
// longer stalls of one warp made the other scan warps wait for the accumulator (exact search 4.53 -> 5.14 ms);
// every lane folding its own pending keys into its sorted prefix by insertion (all 32 rows per event): 13 k
// cycles per event, 1.7 M in total (5.05 ms).  The pending buffers cost the third ring stage (shared memory),
// which is what bounds this mode now: the scan warps wait 19 % of their time for the next accumulator.
__device__ __noinline__ float scan_prune(unsigned long long* wl, int& cnt, float thr, int ks, int lane, int trig) {
    unsigned need = __ballot_sync(kFull, cnt > trig && cnt > 0);
    while (need) {
        const int r = __ffs(need) - 1;
        need &= need - 1;
        const int c_r = __shfl_sync(kFull, cnt, r);
        unsigned long long key = lane < c_r ? wl[r * kScanStride + lane] : 0ull;
        key = warp_sort_desc(key, lane);
        if (lane < ks) wl[r * kScanStride + lane] = key;
        const uint32_t last = __shfl_sync(kFull, (uint32_t)(key >> 32), ks - 1);    // 0: fewer than ks keys so far
        if (lane == r) {
            cnt = c_r < ks ? c_r : ks;
            if (last) thr = ord2f(last);
        }
    }
    __syncwarp();
    return thr;
}

template <int kMetric>
__global__ void __launch_bounds__(kThreads, 1) search_tc_kernel(const SearchParams p,
                                                                const __grid_constant__ CUtensorMap tm_q,
                                                                const __grid_constant__ CUtensorMap tm_x) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    if ((int)(smem - smem_raw) > p.align_slack) {   // cannot happen while the dynamic window is 1 KB aligned
        if (threadIdx.x == 0) printf("search_tc_kernel: shared-memory window misaligned by %d B\n", (int)(smem - smem_raw));
        return;
    }
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qt = p.qt, nchunks = p.nchunks, S = p.stages;
    const uint32_t a_base = sbase;                                       // [qt][nchunks] 16 KB tiles
    const uint32_t b_base = sbase + (uint32_t)(qt * nchunks) * kABytes;  // [S] 16 KB tiles
    uint8_t* tail = smem + (size_t)(qt * nchunks) * kABytes + (size_t)S * kNBytes;
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(tail);   // [qt * 128][ks] sorted, 0 = empty
    const size_t list_bytes = (size_t)(p.scan_append ? kScanStride : p.ks) * qt * kTileM * 8;
    const uint32_t bars = smem_u32(tail + list_bytes);
    auto bar = [&](int slot) { return bars + 8u * (uint32_t)slot; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + list_bytes + kNumBars * 8);
    const int tmem_cols = qt == 2 ? 512 : 256;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(bar(kBarFull + s), 1); mbar_init(bar(kBarEmpty + s), 1); }
        mbar_init(bar(kBarAFull), 1);
        mbar_init(bar(kBarAFree), 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar(kBarTFull + a), 1);
            mbar_init(bar(kBarTEmpty + a), (uint32_t)(qt * 128));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // Work = (query group of qt*128 rows) x (item chunk) sub-units in row-major order, cut into
    // gridDim.x contiguous ranges.  A CTA walks its range as SEGMENTS: maximal runs of chunks of
    // one query group.  Per segment the query tiles are loaded once and one sorted list per row
    // is kept; it is stored at short_keys[first chunk of the segment][q].  Balanced to one chunk,
    // and a query group is split over as few CTAs (= lists to merge) as possible.
    const int64_t rows_per_cta = (int64_t)qt * kTileM;
    const int64_t qgroups = (p.nq + rows_per_cta - 1) / rows_per_cta;
    const int64_t total_sub = qgroups * p.splits;
    const int64_t t_begin = total_sub * blockIdx.x / gridDim.x, t_end = total_sub * (blockIdx.x + 1) / gridDim.x;
    struct Seg { int64_t qg, n_begin, n_end, t_next; int slot, ntiles; };
    auto segment = [&](int64_t t) {
        Seg g;
        g.qg = t / p.splits;
        g.slot = (int)(t % p.splits);
        g.t_next = min(t_end, (g.qg + 1) * p.splits);
        g.n_begin = (int64_t)g.slot * p.split_len;
        g.n_end = min(p.nx, (int64_t)(g.slot + (g.t_next - t)) * p.split_len);
        g.ntiles = g.n_end > g.n_begin ? (int)((g.n_end - g.n_begin + kTileN - 1) / kTileN) : 0;
        return g;
    };

    if (warp == kProdWarp) {
        // ===================== TMA producer =====================
        int gc = 0, it = 0;
        const int chunk_elems = p.kind == 1 ? 64 : kChunkK;
        for (int64_t t = t_begin; t < t_end; ++it) {
            const Seg sg = segment(t);
            t = sg.t_next;
            const int64_t qg = sg.qg, n_begin = sg.n_begin;
            const int ntiles = sg.ntiles;
            if (it > 0) mbar_wait(bar(kBarAFree), (uint32_t)((it - 1) & 1));   // MMAs of the previous segment done
            if (lane == 0) {
                mbar_expect_tx(bar(kBarAFull), (uint32_t)(qt * nchunks) * kABytes);
                for (int qi = 0; qi < qt; ++qi)
                    for (int c = 0; c < nchunks; ++c)
                        tma_load_2d(a_base + (uint32_t)(qi * nchunks + c) * kABytes, &tm_q, c * chunk_elems,
                                    (int)(qg * rows_per_cta + qi * kTileM), bar(kBarAFull));
            }
            __syncwarp();
            for (int nt = 0; nt < ntiles; ++nt) {
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int s = gc % S, uu = gc / S;
                    if (uu > 0) mbar_wait(bar(kBarEmpty + s), (uint32_t)((uu - 1) & 1));
                    if (lane == 0) {
                        mbar_expect_tx(bar(kBarFull + s), (uint32_t)kNBytes);
                        tma_load_2d(b_base + (uint32_t)s * kNBytes, &tm_x, c * chunk_elems,
                                    (int)(n_begin + (int64_t)nt * kTileN), bar(kBarFull + s));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        // D=F32, A=B=TF32, K-major, N=128 (>>3 at bit 17), M=128 (>>4 at bit 24)
        // (kind::f16 with bf16 operands: format code 1 instead of 2)
        const uint32_t fmt = p.kind == 1 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                               ((uint32_t)(kTileM >> 4) << 24);
        int gc = 0, it = 0, tc_ = 0;
        for (int64_t t = t_begin; t < t_end; ++it) {
            const Seg sg = segment(t);
            t = sg.t_next;
            const int ntiles = sg.ntiles;
            mbar_wait(bar(kBarAFull), (uint32_t)(it & 1));
            for (int nt = 0; nt < ntiles; ++nt, ++tc_) {
                const int b = tc_ & 1, ub = tc_ >> 1;
                if (ub > 0) mbar_wait(bar(kBarTEmpty + b), (uint32_t)((ub - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int s = gc % S, uu = gc / S;
                    mbar_wait(bar(kBarFull + s), (uint32_t)(uu & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint32_t b_s = b_base + (uint32_t)s * kNBytes;
                        for (int qi = 0; qi < qt; ++qi) {
                            const uint32_t a_s = a_base + (uint32_t)(qi * nchunks + c) * kABytes;
                            const uint32_t tmem_d = tmem_base + (uint32_t)((b * qt + qi) * kTileN);
                            if (p.kind == 1) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)   // 4 x K=16 bf16 = one 128-byte chunk row
                                    mma_bf16(tmem_d, make_desc(a_s + k * 32), make_desc(b_s + k * 32), idesc,
                                             (uint32_t)((c | k) != 0));
                            } else {
#pragma unroll
                                for (int k = 0; k < kChunkK / 8; ++k)
                                    mma_tf32(tmem_d, make_desc(a_s + k * 32), make_desc(b_s + k * 32), idesc,
                                             (uint32_t)((c | k) != 0));
                            }
                        }
                        mma_commit(bar(kBarEmpty + s));
                        if (c == nchunks - 1) mma_commit(bar(kBarTFull + b));
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) mma_commit(bar(kBarAFree));
            __syncwarp();
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4 * qt) {
        // ===================== scan warps: one thread per query row =====================
        const int e = warp - kEpiWarp0;
        const int qi = e >> 2, quarter = warp & 3;           // TMEM lane quarter = warp % 4
        const int row0 = qi * kTileM + quarter * 32;         // this warp's 32 rows of the CTA
        const int ks = p.ks;
        const bool append = p.scan_append != 0;
        // row r of the warp: wl[r * ks + rank] (sorted), or wl[r * kScanStride + i] (append buffer)
        unsigned long long* wl = lists + (size_t)row0 * (append ? kScanStride : ks);
        int tc_ = 0;
        long long sp_wait = 0, sp_ld = 0, sp_fast = 0, sp_slow = 0, sp_groups = 0, sp_slowg = 0, sp_t0 = SP_NOW();
        long long sp_late_slow = 0, sp_late_slowg = 0, sp_late_cand = 0, sp_late_votes = 0, sp_cand = 0;
        for (int64_t t = t_begin; t < t_end;) {
            const Seg sg = segment(t);
            t = sg.t_next;
            const int64_t qg = sg.qg;
            const int sp = sg.slot, ntiles = sg.ntiles;
            const int n_begin = (int)sg.n_begin, n_end = (int)sg.n_end;
            const int64_t q = qg * rows_per_cta + row0 + lane;
            const bool row_ok = q < p.nq;
            if (!append) for (int i = lane; i < 32 * ks; i += 32) wl[i] = 0ull;
            __syncwarp();
            int cnt = 0;                                    // append mode: keys in this row's buffer
            float thr = row_ok ? -INFINITY : INFINITY;      // score of rank ks-1 once the list is full
            uint4 pm = make_uint4(0u, 0u, 0u, 0u);
            if (p.pmask && row_ok) pm = __ldg(reinterpret_cast<const uint4*>(p.pmask) + q);
            for (int nt = 0; nt < ntiles; ++nt, ++tc_) {
                const int b = tc_ & 1, ub = tc_ >> 1;
                bool elig = true;                           // IVF: does this row's query probe the tile's list?
                if (p.tile_list) {
                    const int l = __ldg(p.tile_list + (n_begin / kTileN + nt));
                    const uint32_t w = l < 32 ? pm.x : (l < 64 ? pm.y : (l < 96 ? pm.z : pm.w));
                    elig = l >= 0 && ((w >> (l & 31)) & 1u);
                }
                float thr_t = elig ? thr : INFINITY;
                const long long w0 = SP_NOW();
                mbar_wait(bar(kBarTFull + b), (uint32_t)(ub & 1));
                SP_ADD(sp_wait, w0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                       (uint32_t)((b * qt + qi) * kTileN);
#pragma unroll 1
                for (int cb = 0; cb < kTileN / 32; ++cb) {
                    uint32_t v[32];
                    const long long l0 = SP_NOW();
                    tmem_ld32(taddr + (uint32_t)(cb * 32), v);
                    SP_ADD(sp_ld, l0);
                    if (cb == kTileN / 32 - 1) {   // accumulator fully read: hand it back
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(bar(kBarTEmpty + b));
                    }
                    const int col0 = n_begin + nt * kTileN + cb * 32;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float f[8];
                        if (kMetric == PB200_METRIC_L2) {
                            const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.hx + col0 + g * 8));
                            const float4 h1 = __ldg(reinterpret_cast<const float4*>(p.hx + col0 + g * 8 + 4));
                            const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]) - hh[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]);
                        }
                        const long long g0 = SP_NOW();
                        const float gm = fmaxf(fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3])),
                                               fmaxf(fmaxf(f[4], f[5]), fmaxf(f[6], f[7])));
                        const bool slowp = __any_sync(kFull, gm >= thr_t);
                        SP_ADD(sp_fast, g0);
                        sp_groups += 1;
                        if (slowp && append) {
                            const long long s0 = SP_NOW();
                            sp_slowg += 1;
                            unsigned long long* mine = wl + lane * kScanStride;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (f[i] >= thr_t && col0 + g * 8 + i < n_end) {
                                    mine[cnt] = make_key(f[i], (uint32_t)(col0 + g * 8 + i));
                                    ++cnt;
                                }
                            }
                            if (__any_sync(kFull, cnt > kScanTrig)) {
                                const long long p0 = SP_NOW();
                                __syncwarp();
                                thr = scan_prune(wl, cnt, thr, ks, lane, kScanTrig);
                                thr_t = elig ? thr : INFINITY;
                                SP_ADD(sp_late_votes, p0);                                   // cycles in prunes
                                sp_late_slowg += 1;                                          // prune events
                            }
                            SP_ADD(sp_slow, s0);
                        } else if (slowp) {
                            const long long s0 = SP_NOW();
                            sp_slowg += 1;
                            // all 8 votes first, against the threshold at the start of the group: a vote
                            // that waits for the previous column's insertion would serialise eight
                            // ~30-cycle round trips; a candidate that a later insertion has outdated
                            // is simply rejected by insert_column (position = ks).  (Merging the
                            // candidates of all 8 columns into 4-row batches was slower: 5.46 vs 4.88 ms.)
                            unsigned m8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                m8[i] = __ballot_sync(kFull, f[i] >= thr_t && col0 + g * 8 + i < n_end);
#ifdef PB200_SEARCH_PROFILE
                            {
                                int nc = 0;
                                for (int i = 0; i < 8; ++i) nc += __popc(m8[i]);
                                sp_cand += nc;
                                if (nt >= 40) { sp_late_cand += nc; sp_late_slowg += 1; SP_ADD(sp_late_votes, s0); }
                            }
#endif
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (m8[i]) {
                                    thr = insert_column(m8[i], f[i], 0xFFFFFFFFu - (uint32_t)(col0 + g * 8 + i), wl, ks,
                                                        lane, thr);
                                    thr_t = elig ? thr : INFINITY;
                                }
                            }
                            SP_ADD(sp_slow, s0);
#ifdef PB200_SEARCH_PROFILE
                            if (nt >= 40) SP_ADD(sp_late_slow, s0);
#endif
                        }
                    }
                }
            }
#ifdef PB200_SEARCH_PROFILE
            if (t >= t_end && lane == 0 && blockIdx.x == 3)
                printf("scan warp %d: total %lld  wait_tfull %lld  tmem_ld %lld  fast %lld  slow %lld  groups %lld  slow_groups %lld  "
                       "candidates %lld | tiles >= 40 of a segment: slow %lld  slow_groups %lld  candidates %lld  votes %lld\n",
                       e, (long long)(clock64() - sp_t0), sp_wait, sp_ld, sp_fast, sp_slow, sp_groups, sp_slowg, sp_cand,
                       sp_late_slow, sp_late_slowg, sp_late_cand, sp_late_votes);
#endif
            // segment done: every row's sorted list -> short_keys[sp][q][0..ks)
            __syncwarp();
            if (append) {
                thr = scan_prune(wl, cnt, thr, ks, lane, -1);       // sort what is left in every buffer
                for (int r = 0; r < 32; ++r) {
                    const int64_t qr = qg * rows_per_cta + row0 + r;
                    if (qr >= p.nq) break;
                    const int c_r = __shfl_sync(kFull, cnt, r);
                    if (lane < ks)
                        p.short_keys[((size_t)sp * p.nq + qr) * ks + lane] = lane < c_r ? wl[r * kScanStride + lane] : 0ull;
                }
            } else {
                for (int r = 0; r < 32; ++r) {
                    const int64_t qr = qg * rows_per_cta + row0 + r;
                    if (qr >= p.nq) break;
                    if (lane < ks) p.short_keys[((size_t)sp * p.nq + qr) * ks + lane] = wl[r * ks + lane];
                }
            }
            __syncwarp();
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

// ---- operand preparation: TF32-rounded copies with their rounding residuals, 0.5|x|^2 ----
// One warp per row: out = rna_tf32(in); resid[row] = |out - in|_2 (optional); the maxima of the
// residual and of |out|_2 over all rows go to max_bits[0..1] (non-negative floats: bit patterns
// are ordered).  The residual norms make the error bound of the certificate data dependent:
// |<q~,x~> - <q,x>| <= |q| |x~ - x| + |q~ - q| |x~|  (Cauchy-Schwarz), ~3x tighter than 2^-10 |q||x|.
__global__ void __launch_bounds__(256) round_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         int64_t n, int d, float* __restrict__ resid,
                                                         unsigned int* max_bits) {
    __shared__ float s_r[8], s_n[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    float r2 = 0.f, n2 = 0.f;
    for (int c = lane * 4; row < n && c < d; c += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + row * d + c));
        float4 o;
        o.x = __uint_as_float(to_tf32(v.x)); o.y = __uint_as_float(to_tf32(v.y));
        o.z = __uint_as_float(to_tf32(v.z)); o.w = __uint_as_float(to_tf32(v.w));
        *reinterpret_cast<float4*>(out + row * d + c) = o;
        const float e0 = o.x - v.x, e1 = o.y - v.y, e2 = o.z - v.z, e3 = o.w - v.w;   // exact (Sterbenz)
        r2 = fmaf(e0, e0, r2); r2 = fmaf(e1, e1, r2); r2 = fmaf(e2, e2, r2); r2 = fmaf(e3, e3, r2);
        n2 = fmaf(o.x, o.x, n2); n2 = fmaf(o.y, o.y, n2); n2 = fmaf(o.z, o.z, n2); n2 = fmaf(o.w, o.w, n2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        r2 += __shfl_xor_sync(kFull, r2, o);
        n2 += __shfl_xor_sync(kFull, n2, o);
    }
    const float r = sqrtf(r2) * 1.0001f, nn = sqrtf(n2) * 1.0001f;   // cover the fp32 rounding of the sums
    if (lane == 0) {
        if (resid && row < n) resid[row] = r;
        s_r[warp] = (row < n && r == r) ? r : 0.f;
        s_n[warp] = (row < n && nn == nn) ? nn : 0.f;
    }
    __syncthreads();
    if (max_bits && threadIdx.x == 0) {          // one atomic pair per block, not per row
        float mr = 0.f, mn = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { mr = fmaxf(mr, s_r[w]); mn = fmaxf(mn, s_n[w]); }
        if (mr > 0.f) atomicMax(max_bits, __float_as_uint(mr));
        if (mn > 0.f) atomicMax(max_bits + 1, __float_as_uint(mn));
    }
}

__global__ void half_norm_max_kernel(const float* __restrict__ xn, int64_t nx, int64_t nx_pad,
                                     float* __restrict__ hx, unsigned int* xmax_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx_pad;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float v = i < nx ? xn[i] : 0.f;
        hx[i] = 0.5f * v;
        if (v == v) m = fmaxf(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(xmax_bits, __float_as_uint(m));   // m >= 0: bits are ordered
}

struct RerankParams {
    const float* __restrict__ q; const float* __restrict__ x;
    int64_t nq, nx;
    int d, k, ks, splits, metric;
    const float* __restrict__ qn; const float* __restrict__ xn;
    const unsigned int* __restrict__ xmax_bits;   // [0] max |x|^2, [1] max |x~ - x|, [2] max |x~|
    const float* __restrict__ q_resid;            // |q~ - q| per query
    const int32_t* __restrict__ exclude; int32_t id_offset;
    const unsigned long long* __restrict__ short_keys;
    float* __restrict__ out_scores; int32_t* __restrict__ out_ids;
    int32_t* qsel; int32_t* qsel_count;
    // IVF mode: shortlisted positions are rows of the padded list-ordered copy; src_pos maps them
    // to list order (-1 = padding), x = list_vecs (fp32, list order), ids_map = list_ids, and the
    // distance is the direct form sum (q - v)^2 of ivf_search_kernel (same arithmetic, same order)
    const int32_t* __restrict__ src_pos; const int32_t* __restrict__ ids_map;
};

__global__ void __launch_bounds__(256) rerank_kernel(const RerankParams p) {
    extern __shared__ float s_qrows[];   // [8][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 8 + warp;
    if (q >= p.nq) return;
    float* sq = s_qrows + (size_t)warp * p.d;
    for (int c = lane; c < p.d; c += 32) sq[c] = __ldg(p.q + q * p.d + c);
    // t = the best TF32 score an item OUTSIDE the shortlists can have
    float tt = -INFINITY;
    for (int sp = lane; sp < p.splits; sp += 32) {
        const unsigned long long last = p.short_keys[((size_t)sp * p.nq + q) * p.ks + p.ks - 1];
        if (last) tt = fmaxf(tt, ord2f((uint32_t)(last >> 32)));
    }
    for (int o = 16; o > 0; o >>= 1) tt = fmaxf(tt, __shfl_xor_sync(kFull, tt, o));
    __syncwarp();
    const float qn = p.qn[q];
    const int excl = p.exclude ? p.exclude[q] : -1;
    TopkLane best; best.bad = INFINITY; best.id = INT_MAX;
    const int nc = p.splits * p.ks;
    for (int base = 0; base < nc; base += 32) {
        const int j = base + lane;
        unsigned long long key = 0ull;
        if (j < nc) key = p.short_keys[((size_t)(j / p.ks) * p.nq + q) * p.ks + (j % p.ks)];
        bool valid = key != 0ull;
        const uint32_t col = 0xFFFFFFFFu - (uint32_t)key;
        float bad = INFINITY;
        int gid = -1;
        if (valid && p.src_pos) {
            const int pos = p.src_pos[col];
            valid = pos >= 0;
            if (valid) {
                const float4* xr = reinterpret_cast<const float4*>(p.x + (int64_t)pos * p.d);
                float dist = 0.f;
                for (int c4 = 0; c4 < p.d / 4; ++c4) {
                    const float4 xv = __ldg(xr + c4);
                    const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * c4);
                    float t = qv.x - xv.x; dist = fmaf(t, t, dist);
                    t = qv.y - xv.y; dist = fmaf(t, t, dist);
                    t = qv.z - xv.z; dist = fmaf(t, t, dist);
                    t = qv.w - xv.w; dist = fmaf(t, t, dist);
                }
                bad = dist;
                gid = p.ids_map[pos];
                valid = bad == bad;
            }
        } else if (valid) {
            // the fp32 kernel's arithmetic: acc = fmaf(q[c], x[c], acc), c ascending
            const float4* xr = reinterpret_cast<const float4*>(p.x + (int64_t)col * p.d);
            float acc = 0.f;
            for (int c4 = 0; c4 < p.d / 4; ++c4) {
                const float4 xv = __ldg(xr + c4);
                const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * c4);
                acc = fmaf(qv.x, xv.x, acc); acc = fmaf(qv.y, xv.y, acc);
                acc = fmaf(qv.z, xv.z, acc); acc = fmaf(qv.w, xv.w, acc);
            }
            bad = p.metric == PB200_METRIC_IP ? -acc : fmaxf(qn + __ldg(p.xn + col) - 2.f * acc, 0.f);
            gid = (int)col + p.id_offset;
            valid = gid != excl && bad == bad;
        }
        topk_offer(best, bad, gid, valid, p.k, lane);
    }
    const int nsel = __popc(__ballot_sync(kFull, lane < p.k && best.id != INT_MAX));
    const float bad_k = __shfl_sync(kFull, best.bad, p.k - 1);
    bool certified = tt == -INFINITY;   // every list short: the shortlists hold all items
    if (!certified && nsel == p.k) {
        const float xmax = __uint_as_float(p.xmax_bits[0]);
        // |<q~,x~> - <q,x>| <= |q| |x~ - x| + |q~ - q| |x~|, plus the accumulation error of the
        // tensor core (d products summed in fp32, budgeted at 2^-22 |q||x| each), 2 % margin
        const float qnorm = sqrtf(qn), xnorm = sqrtf(xmax);
        const float eps = 1.02f * (qnorm * __uint_as_float(p.xmax_bits[1]) +
                                   p.q_resid[q] * __uint_as_float(p.xmax_bits[2]) +
                                   (float)p.d * 2.384185791015625e-07f * qnorm * xnorm);
        // an outside item's `bad` is at least this:
        // (IVF: the direct-form distance and |q|^2 carry ~d 2^-24 relative rounding each)
        const float slack = p.src_pos ? (float)p.d * 2.4e-7f * (qn + xmax) : 4e-6f * (qn + xmax);
        const float bound = p.metric == PB200_METRIC_IP ? -(tt + eps) : qn - 2.f * (tt + eps) - slack;
        certified = bad_k < bound;
    }
    if (certified) {
        if (lane < p.k) {
            const bool has = best.id != INT_MAX;
            const bool largest = p.metric == PB200_METRIC_IP;
            p.out_ids[q * p.k + lane] = has ? best.id : -1;
            p.out_scores[q * p.k + lane] = has ? (largest ? -best.bad : best.bad) : (largest ? -INFINITY : INFINITY);
        }
    } else if (lane == 0) {
        p.qsel[atomicAdd(p.qsel_count, 1)] = (int32_t)q;
    }
}

// ---- host side ----
struct Plan {
    int qt = 0, nchunks = 0, ks = 0, splits = 0, stages = 0, grid = 0, fsplits = 0, fsplits0 = 0, align_slack = 1024;
    int scan_append = 0;
    int64_t split_len, nx_pad, cap_f, cap_f0;
    size_t smem_bytes;
    // workspace offsets
    size_t off_qr, off_xr, off_qn, off_xn, off_hx, off_qres, off_misc, off_short, off_qsel, off_pbad,
        off_pids, total;
};

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// kernel geometry for nq x nx scores with `nchunks` 128-byte K chunks per row and lists of p.ks
static bool make_geometry(int64_t nq, int64_t nx, Plan& p) {
    // Two resident query tiles per CTA halve the L2 -> SM item traffic; needs room for the operands,
    // the sorted lists and >= 3 ring stages.  (2 CTAs/SM with one tile each -- twice the scan warps --
    // was measured slower, 4.99 vs 4.28 ms at C3, and removed.)
    p.align_slack = 1024;
    auto fixed_bytes = [&](int qt) {
        return (size_t)qt * p.nchunks * kABytes + (size_t)(p.scan_append ? kScanStride : p.ks) * qt * kTileM * 8 +
               kNumBars * 8 + 16;
    };
    const size_t budget = 225 * 1024;
    // append-mode buffers are twice the sorted lists: two ring stages are enough there (the scan, not the
    // item stream, sets the pace: tensor pipe 20-25 % busy), the sorted-insertion mode keeps >= 3
    const int min_stages2 = p.scan_append ? 2 : 3;
    p.qt = (nq > 128 && fixed_bytes(2) + 1024 + min_stages2 * (size_t)kNBytes <= budget) ? 2 : 1;
    const size_t fixed = fixed_bytes(p.qt) + p.align_slack;
    if (fixed + 2 * (size_t)kNBytes > budget) return false;
    int st = (int)((budget - fixed) / kNBytes);
    if (st > kMaxStages) st = kMaxStages;
    p.stages = st;
    p.smem_bytes = fixed + (size_t)st * kNBytes;
    const int64_t rows = (int64_t)p.qt * kTileM;
    const int64_t qgroups = ceil_div(nq, rows);
    // item chunks per query group: enough sub-units that the contiguous partition over the CTAs is
    // balanced to ~1/6 of a CTA's share (many query groups -> few chunks -> few lists per query)
    int chunks = 1;
    const int forced = env_int("PB200_TOPK_TC_SPLITS", 0);
    while (chunks < 64 && qgroups * chunks < 6 * kSMs && ceil_div(nx, chunks * 2) >= 512) chunks *= 2;
    if (forced > 0 && forced <= 64) chunks = forced;
    p.split_len = ceil_div(ceil_div(nx, chunks), kTileN) * kTileN;
    p.splits = (int)ceil_div(nx, p.split_len);          // no empty chunk
    const int64_t total_sub = qgroups * p.splits;
    p.grid = (int)(total_sub < kSMs ? total_sub : kSMs);
    return true;
}

static bool make_plan(int64_t nq, int64_t nx, int dim, int k, bool has_exclude, Plan* pl) {
    if (dim % 4 || dim > 256 || dim <= 0 || nq <= 0 || nx <= 0) return false;
    const int need = k + (has_exclude ? 1 : 0);
    if (need > 24) return false;
    Plan p{};
    p.ks = need <= 12 ? 16 : 32;
    const int ks_env = env_int("PB200_TOPK_TC_KS", 0);
    if (ks_env == 16 || ks_env == 32) { if (ks_env >= need) p.ks = ks_env; }
    p.nchunks = (dim + kChunkK - 1) / kChunkK;
    p.scan_append = (p.ks == 16 && env_int("PB200_TOPK_TC_SCAN", 1) != 0) ? 1 : 0;
    if (!make_geometry(nq, nx, p)) return false;
    p.nx_pad = ceil_div(nx, kTileN) * kTileN + kTileN;
    // fp32 re-runs: a first round for up to 1,024 uncertified queries with many item splits (the
    // usual case is a handful of queries: latency of one block chain), then nq/4 per round
    p.cap_f0 = nq < 1024 ? (int64_t)align_up((size_t)nq, 64) : 1024;
    p.fsplits0 = (int)(nx / 512 > 0 ? (nx / 512 < 64 ? nx / 512 : 64) : 1);
    p.cap_f = (int64_t)align_up((size_t)ceil_div(nq, 4), 64);
    p.fsplits = (int)(nx / 512 > 0 ? (nx / 512 < 16 ? nx / 512 : 16) : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t r = off; off += align_up(bytes, 256); return r; };
    p.off_qr = take((size_t)nq * dim * 4);
    p.off_xr = take((size_t)nx * dim * 4);
    p.off_qn = take((size_t)nq * 4);
    p.off_xn = take((size_t)nx * 4);
    p.off_hx = take((size_t)p.nx_pad * 4);
    p.off_qres = take((size_t)nq * 4);
    p.off_misc = take(256);
    p.off_short = take((size_t)p.splits * nq * p.ks * 8);
    p.off_qsel = take((size_t)nq * 4);
    const size_t part = (size_t)(p.cap_f * p.fsplits > p.cap_f0 * p.fsplits0 ? p.cap_f * p.fsplits : p.cap_f0 * p.fsplits0) * 32 * 4;
    p.off_pbad = take(part);
    p.off_pids = take(part);
    p.total = off;
    *pl = p;
    return true;
}

}  // namespace tcs
}  // namespace pb200

using namespace pb200;

extern "C" int pb200_topk_tc_supported(int64_t nq, int64_t nx, int dim, int k, int has_exclude) {
    tcs::Plan pl;
    return tcs::make_plan(nq, nx, dim, k, has_exclude != 0, &pl) ? 1 : 0;
}

extern "C" size_t pb200_topk_tc_workspace_bytes(int64_t nq, int64_t nx, int dim, int k) {
    tcs::Plan a{}, b{};
    const bool oka = tcs::make_plan(nq, nx, dim, k, false, &a);
    const bool okb = tcs::make_plan(nq, nx, dim, k, true, &b);
    const size_t ta = oka ? a.total : 0, tb = okb ? b.total : 0;
    return ta > tb ? ta : tb;
}

extern "C" int pb200_topk_tc(const float* queries, int64_t nq, const float* items, int64_t nx,
                             int dim, int k, int metric, const int32_t* exclude_ids,
                             int32_t id_offset, float* out_scores, int32_t* out_ids,
                             void* workspace, size_t workspace_bytes, int32_t* stats_out,
                             pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(nq >= 0 && nx >= 0 && dim > 0 && k > 0, "topk_tc: bad sizes");
    PB_REQUIRE(metric == PB200_METRIC_IP || metric == PB200_METRIC_L2, "topk_tc: unknown metric");
    PB_REQUIRE(nx + (int64_t)id_offset < 2147483647ll, "topk_tc: ids overflow int32");
    if (nq == 0) return PB200_OK;
    tcs::Plan pl;
    if (!tcs::make_plan(nq, nx, dim, k, exclude_ids != nullptr, &pl)) {
        set_error("topk_tc: needs dim %% 4 == 0, dim <= 256, k (+1 with exclude_ids) <= 24, nx > 0 "
                  "(got dim=%d k=%d nx=%lld)", dim, k, (long long)nx);
        return PB200_ERR_UNSUPPORTED;
    }
    PB_REQUIRE(queries && items && out_scores && out_ids && workspace, "topk_tc: null pointer");
    PB_REQUIRE(((uintptr_t)queries | (uintptr_t)items | (uintptr_t)workspace) % 16 == 0,
               "topk_tc: queries / items / workspace must be 16-byte aligned");
    if (workspace_bytes < pl.total) {
        set_error("topk_tc: workspace %zu B < required %zu B", workspace_bytes, pl.total);
        return PB200_ERR_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    float* qr = reinterpret_cast<float*>(ws + pl.off_qr);
    float* xr = reinterpret_cast<float*>(ws + pl.off_xr);
    float* qn = reinterpret_cast<float*>(ws + pl.off_qn);
    float* xn = reinterpret_cast<float*>(ws + pl.off_xn);
    float* hx = reinterpret_cast<float*>(ws + pl.off_hx);
    unsigned int* xmax_bits = reinterpret_cast<unsigned int*>(ws + pl.off_misc);
    int32_t* qsel_count = reinterpret_cast<int32_t*>(ws + pl.off_misc + 64);
    int32_t* qsel = reinterpret_cast<int32_t*>(ws + pl.off_qsel);

    PB_CUDA(cudaMemsetAsync(ws + pl.off_misc, 0, 256, stream));
    float* q_resid = reinterpret_cast<float*>(ws + pl.off_qres);
    tcs::round_rows_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, stream>>>(queries, qr, nq, dim, q_resid, nullptr);
    int rc = check_launch("round_rows_kernel");
    if (rc) return rc;
    tcs::round_rows_kernel<<<(unsigned)ceil_div(nx, 8), 256, 0, stream>>>(items, xr, nx, dim, nullptr, xmax_bits + 1);
    rc = check_launch("round_rows_kernel");
    if (rc) return rc;
    rc = row_sqnorm_run(queries, nq, dim, qn, stream);
    if (rc) return rc;
    rc = row_sqnorm_run(items, nx, dim, xn, stream);
    if (rc) return rc;
    {
        const int64_t blocks = ceil_div(pl.nx_pad, 256) < kSMs * 4 ? ceil_div(pl.nx_pad, 256) : kSMs * 4;
        tcs::half_norm_max_kernel<<<(unsigned)blocks, 256, 0, stream>>>(xn, nx, pl.nx_pad, hx, xmax_bits);
        rc = check_launch("half_norm_max_kernel");
        if (rc) return rc;
    }

    alignas(64) CUtensorMap tm_q, tm_x;
    if (!make_map(&tm_q, qr, nq, dim, tc::kTileM) || !make_map(&tm_x, xr, nx, dim, tcs::kTileN)) {
        set_error("topk_tc: cuTensorMapEncodeTiled failed (nq=%lld nx=%lld dim=%d)", (long long)nq,
                  (long long)nx, dim);
        return PB200_ERR_CUDA;
    }
    tcs::SearchParams sp{};
    sp.nq = nq; sp.nx = nx; sp.d = dim; sp.nchunks = pl.nchunks; sp.qt = pl.qt; sp.ks = pl.ks;
    sp.splits = pl.splits; sp.split_len = pl.split_len; sp.stages = pl.stages; sp.align_slack = pl.align_slack; sp.kind = 0; sp.scan_append = pl.scan_append;
    sp.hx = hx;
    sp.short_keys = reinterpret_cast<unsigned long long*>(ws + pl.off_short);
    // slots no segment starts at stay empty (0)
    PB_CUDA(cudaMemsetAsync(sp.short_keys, 0, (size_t)pl.splits * nq * pl.ks * 8, stream));
    if (metric == PB200_METRIC_IP) {
        PB_CUDA(cudaFuncSetAttribute(tcs::search_tc_kernel<PB200_METRIC_IP>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
        tcs::search_tc_kernel<PB200_METRIC_IP><<<pl.grid, tcs::kThreads, pl.smem_bytes, stream>>>(sp, tm_q, tm_x);
    } else {
        PB_CUDA(cudaFuncSetAttribute(tcs::search_tc_kernel<PB200_METRIC_L2>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
        tcs::search_tc_kernel<PB200_METRIC_L2><<<pl.grid, tcs::kThreads, pl.smem_bytes, stream>>>(sp, tm_q, tm_x);
    }
    rc = check_launch("search_tc_kernel");
    if (rc) return rc;

    tcs::RerankParams rp{};
    rp.q = queries; rp.x = items; rp.nq = nq; rp.nx = nx; rp.d = dim; rp.k = k; rp.ks = pl.ks;
    rp.splits = pl.splits; rp.metric = metric; rp.qn = qn; rp.xn = xn; rp.xmax_bits = xmax_bits; rp.q_resid = q_resid;
    rp.exclude = exclude_ids; rp.id_offset = id_offset; rp.short_keys = sp.short_keys;
    rp.out_scores = out_scores; rp.out_ids = out_ids; rp.qsel = qsel; rp.qsel_count = qsel_count;
    tcs::rerank_kernel<<<(unsigned)ceil_div(nq, 8), 256, (size_t)8 * dim * 4, stream>>>(rp);
    rc = check_launch("rerank_kernel");
    if (rc) return rc;

    // uncertified queries: exact fp32 kernel on the compacted list, cap_f slots per launch
    float* part_bad = reinterpret_cast<float*>(ws + pl.off_pbad);
    int32_t* part_ids = reinterpret_cast<int32_t*>(ws + pl.off_pids);
    rc = topk_fp32_run(queries, pl.cap_f0 < nq ? pl.cap_f0 : nq, items, nx, dim, k, metric, qn, xn, exclude_ids,
                       id_offset, out_scores, out_ids, part_bad, part_ids, pl.fsplits0, qsel, qsel_count, 0, stream);
    if (rc) return rc;
    for (int64_t base = pl.cap_f0; base < nq; base += pl.cap_f) {
        const int64_t slots = nq - base < pl.cap_f ? nq - base : pl.cap_f;
        rc = topk_fp32_run(queries, slots, items, nx, dim, k, metric, qn, xn, exclude_ids, id_offset,
                           out_scores, out_ids, part_bad, part_ids, pl.fsplits, qsel, qsel_count, base,
                           stream);
        if (rc) return rc;
    }
    if (stats_out)   // [0] = number of queries re-run in fp32 (device-side counter, copied in stream order)
        PB_CUDA(cudaMemcpyAsync(stats_out, qsel_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    return PB200_OK;
}


// =====================================================================================
// Exhaustive Hamming top-k (LSHIndex.search, utils/nearest_neighbors.py:47-68; what
// faiss.IndexLSH computes) on the tensor cores: codes expanded to +-1 bf16 vectors,
// <a, b> = nbits - 2 hamming(a, b) -- an exact small integer in fp32 -- so the same fused
// GEMM + shortlist kernel (kind::f16, bf16 operands) yields the exact top-k: no re-rank, no
// certificate.  Order: (distance asc, id asc) = pb200_hamming_topk.
// =====================================================================================
namespace pb200 {
namespace tcs {

__global__ void expand_codes_kernel(const uint8_t* __restrict__ codes, int64_t nbytes, uint4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = codes[i];   // bit j (LSB first) -> element 8 i + j: 1 -> +1.0, 0 -> -1.0 (bf16)
        uint4 o;
        o.x = ((b & 1u) ? 0x3F80u : 0xBF80u) | (((b & 2u) ? 0x3F80u : 0xBF80u) << 16);
        o.y = ((b & 4u) ? 0x3F80u : 0xBF80u) | (((b & 8u) ? 0x3F80u : 0xBF80u) << 16);
        o.z = ((b & 16u) ? 0x3F80u : 0xBF80u) | (((b & 32u) ? 0x3F80u : 0xBF80u) << 16);
        o.w = ((b & 64u) ? 0x3F80u : 0xBF80u) | (((b & 128u) ? 0x3F80u : 0xBF80u) << 16);
        out[i] = o;
    }
}

// warp per query: merge the per-segment lists (already exact) -> distances + ids
__global__ void __launch_bounds__(256) hamming_finish_kernel(const unsigned long long* __restrict__ short_keys,
                                                             int64_t nq, int splits, int ks, int k, int nbits,
                                                             int32_t id_offset, float* __restrict__ out_dist,
                                                             int32_t* __restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    TopkLane best; best.bad = INFINITY; best.id = INT_MAX;
    const int nc = splits * ks;
    for (int base = 0; base < nc; base += 32) {
        const int j = base + lane;
        unsigned long long key = 0ull;
        if (j < nc) key = short_keys[((size_t)(j / ks) * nq + q) * ks + (j % ks)];
        const bool valid = key != 0ull;
        const float dot = ord2f((uint32_t)(key >> 32));
        const float dist = 0.5f * ((float)nbits - dot);          // exact: both are small integers
        const int gid = (int)(0xFFFFFFFFu - (uint32_t)key) + id_offset;
        topk_offer(best, dist, gid, valid, k, lane);
    }
    if (lane < k) {
        const bool has = best.id != INT_MAX;
        out_ids[q * k + lane] = has ? best.id : -1;
        out_dist[q * k + lane] = has ? best.bad : INFINITY;
    }
}

struct HPlan { Plan g; size_t off_q, off_x, off_short, total; bool shared; };

static bool make_hplan(int64_t nq, int64_t nx, int code_bytes, int k, bool shared, HPlan* hp) {
    const int nbits = code_bytes * 8;
    if (code_bytes <= 0 || nbits > 512 || k <= 0 || k > 32 || nq <= 0 || nx <= 0) return false;
    HPlan h{};
    h.g.ks = k <= 16 ? 16 : 32;
    h.g.scan_append = (h.g.ks == 16 && env_int("PB200_TOPK_TC_SCAN", 1) != 0) ? 1 : 0;
    h.g.nchunks = (nbits + 63) / 64;
    if (!make_geometry(nq, nx, h.g)) return false;
    h.shared = shared;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t r = off; off += align_up(bytes, 256); return r; };
    h.off_x = take((size_t)nx * nbits * 2);
    h.off_q = shared ? h.off_x : take((size_t)nq * nbits * 2);
    h.off_short = take((size_t)h.g.splits * nq * h.g.ks * 8);
    h.total = off;
    *hp = h;
    return true;
}

}  // namespace tcs
}  // namespace pb200

extern "C" int pb200_hamming_topk_tc_supported(int64_t nq, int64_t nx, int code_bytes, int k) {
    tcs::HPlan h;
    return tcs::make_hplan(nq, nx, code_bytes, k, false, &h) ? 1 : 0;
}

extern "C" size_t pb200_hamming_topk_tc_workspace_bytes(int64_t nq, int64_t nx, int code_bytes, int k) {
    tcs::HPlan h;
    return tcs::make_hplan(nq, nx, code_bytes, k, false, &h) ? h.total : 0;
}

extern "C" int pb200_hamming_topk_tc(const uint8_t* codes_q, int64_t nq, const uint8_t* codes_x, int64_t nx,
                                     int code_bytes, int k, int32_t id_offset, float* out_dist,
                                     int32_t* out_ids, void* workspace, size_t workspace_bytes,
                                     pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(nq >= 0 && nx >= 0 && code_bytes > 0 && k > 0, "hamming_topk_tc: bad sizes");
    PB_REQUIRE(nx + (int64_t)id_offset < 2147483647ll, "hamming_topk_tc: ids overflow int32");
    if (nq == 0) return PB200_OK;
    const bool shared = codes_q == codes_x && nq == nx;
    tcs::HPlan h;
    if (!tcs::make_hplan(nq, nx, code_bytes, k, shared, &h)) {
        set_error("hamming_topk_tc: needs code_bytes <= 64, k <= 32, nx > 0 (got code_bytes=%d k=%d nx=%lld)",
                  code_bytes, k, (long long)nx);
        return PB200_ERR_UNSUPPORTED;
    }
    PB_REQUIRE(codes_q && codes_x && out_dist && out_ids && workspace, "hamming_topk_tc: null pointer");
    PB_REQUIRE((uintptr_t)workspace % 16 == 0, "hamming_topk_tc: workspace must be 16-byte aligned");
    if (workspace_bytes < h.total) {
        set_error("hamming_topk_tc: workspace %zu B < required %zu B", workspace_bytes, h.total);
        return PB200_ERR_WORKSPACE;
    }
    const int nbits = code_bytes * 8;
    char* ws = static_cast<char*>(workspace);
    void* xb = ws + h.off_x;
    void* qb = ws + h.off_q;
    {
        const int64_t nb = nx * code_bytes;
        const int64_t blocks = ceil_div(nb, 256) < kSMs * 8 ? ceil_div(nb, 256) : kSMs * 8;
        tcs::expand_codes_kernel<<<(unsigned)blocks, 256, 0, stream>>>(codes_x, nb, static_cast<uint4*>(xb));
        int rc = check_launch("expand_codes_kernel");
        if (rc) return rc;
        if (!shared) {
            const int64_t nbq = nq * code_bytes;
            const int64_t bq = ceil_div(nbq, 256) < kSMs * 8 ? ceil_div(nbq, 256) : kSMs * 8;
            tcs::expand_codes_kernel<<<(unsigned)bq, 256, 0, stream>>>(codes_q, nbq, static_cast<uint4*>(qb));
            rc = check_launch("expand_codes_kernel");
            if (rc) return rc;
        }
    }
    alignas(64) CUtensorMap tm_q, tm_x;
    if (!make_map_bf16(&tm_q, qb, nq, nbits, tc::kTileM) || !make_map_bf16(&tm_x, xb, nx, nbits, tcs::kTileN)) {
        set_error("hamming_topk_tc: cuTensorMapEncodeTiled failed (nq=%lld nx=%lld bits=%d)", (long long)nq,
                  (long long)nx, nbits);
        return PB200_ERR_CUDA;
    }
    const tcs::Plan& pl = h.g;
    tcs::SearchParams sp{};
    sp.nq = nq; sp.nx = nx; sp.d = nbits; sp.nchunks = pl.nchunks; sp.qt = pl.qt; sp.ks = pl.ks;
    sp.splits = pl.splits; sp.split_len = pl.split_len; sp.stages = pl.stages; sp.align_slack = pl.align_slack;
    sp.scan_append = pl.scan_append;
    sp.kind = 1; sp.hx = nullptr;
    sp.short_keys = reinterpret_cast<unsigned long long*>(ws + h.off_short);
    PB_CUDA(cudaMemsetAsync(sp.short_keys, 0, (size_t)pl.splits * nq * pl.ks * 8, stream));
    PB_CUDA(cudaFuncSetAttribute(tcs::search_tc_kernel<PB200_METRIC_IP>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    tcs::search_tc_kernel<PB200_METRIC_IP><<<pl.grid, tcs::kThreads, pl.smem_bytes, stream>>>(sp, tm_q, tm_x);
    int rc = check_launch("search_tc_kernel");
    if (rc) return rc;
    tcs::hamming_finish_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, stream>>>(
        sp.short_keys, nq, pl.splits, pl.ks, k, nbits, id_offset, out_dist, out_ids);
    return check_launch("hamming_finish_kernel");
}


// =====================================================================================
// IVF "Weak AND" search (WeakANDIndex.search, utils/nearest_neighbors.py:115-139) on the
// tensor cores: the items are the list-ordered vectors with every list padded to whole
// 128-item tiles (TF32-rounded copy `xp`), the scoring GEMM covers all of them, and a row
// scans a tile only if its query probes the tile's list (128-bit probe mask per query) -- the
// lists a query does not probe never reach its shortlist.  Exact fp32 re-rank with
// ivf_search_kernel's direct-form distance, certificate as in pb200_topk_tc, uncertified
// queries re-run by ivf_search_kernel: results equal pb200_ivf_search bit for bit.
// =====================================================================================
namespace pb200 {
namespace tcs {

__global__ void probe_mask_kernel(const int32_t* __restrict__ probes, int64_t nq, int nprobe, uint4* __restrict__ pmask) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int j = 0; j < nprobe; ++j) {
            const int l = probes[q * nprobe + j];
            if (l >= 0 && l < 128) w[l >> 5] |= 1u << (l & 31);
        }
        pmask[q] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

struct IPlan { Plan g; int64_t cap_pp; size_t off_qr, off_qn, off_qres, off_misc, off_pmask, off_short, off_qsel, off_pbad, off_pids, total; };

static bool make_iplan(int64_t nq, int64_t np, int dim, int k, int nlist, int nprobe, IPlan* ip) {
    if (dim % 4 || dim > 256 || dim <= 0 || nq <= 0 || np <= 0 || np % kTileN || nlist > 128 || k <= 0 || k > 24)
        return false;
    IPlan h{};
    h.g.ks = k <= 12 ? 16 : 32;
    h.g.scan_append = (h.g.ks == 16 && env_int("PB200_TOPK_TC_SCAN", 1) != 0) ? 1 : 0;
    h.g.nchunks = (dim + kChunkK - 1) / kChunkK;
    if (!make_geometry(nq, np, h.g)) return false;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t r = off; off += align_up(bytes, 256); return r; };
    h.off_qr = take((size_t)nq * dim * 4);
    h.off_qn = take((size_t)nq * 4);
    h.off_qres = take((size_t)nq * 4);
    h.off_misc = take(256);
    h.off_pmask = take((size_t)nq * 16);
    h.off_short = take((size_t)h.g.splits * nq * h.g.ks * 8);
    h.off_qsel = take((size_t)nq * 4);
    h.cap_pp = nq < 4096 ? nq : 4096;     // uncertified queries re-run with one warp per (query, probe)
    h.off_pbad = take((size_t)h.cap_pp * nprobe * 32 * 4);
    h.off_pids = take((size_t)h.cap_pp * nprobe * 32 * 4);
    h.total = off;
    *ip = h;
    return true;
}

}  // namespace tcs
}  // namespace pb200

extern "C" int pb200_ivf_search_tc_supported(int64_t nq, int64_t np, int dim, int k, int nlist) {
    tcs::IPlan h;
    return tcs::make_iplan(nq, np, dim, k, nlist, 1, &h) ? 1 : 0;
}

extern "C" size_t pb200_ivf_search_tc_workspace_bytes(int64_t nq, int64_t np, int dim, int k, int nlist,
                                                      int nprobe) {
    tcs::IPlan h;
    return tcs::make_iplan(nq, np, dim, k, nlist, nprobe, &h) ? h.total : 0;
}

extern "C" int pb200_ivf_search_tc(const float* queries, int64_t nq, int dim, const int32_t* probes, int nprobe,
                                   int nlist, const int32_t* list_offsets, const int32_t* list_ids,
                                   const float* list_vecs, const float* xp, const float* hxp,
                                   const int32_t* src_pos, const int32_t* tile_list, int64_t np,
                                   const uint32_t* xstats, int k, float* out_dist, int32_t* out_ids,
                                   void* workspace, size_t workspace_bytes, int32_t* stats_out,
                                   pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(nq >= 0 && dim > 0 && nprobe > 0 && k > 0, "ivf_search_tc: bad sizes");
    if (nq == 0) return PB200_OK;
    tcs::IPlan h;
    if (!tcs::make_iplan(nq, np, dim, k, nlist, nprobe, &h)) {
        set_error("ivf_search_tc: needs dim %% 4 == 0, dim <= 256, k <= 24, nlist <= 128, padded items %% 128 == 0 "
                  "(got dim=%d k=%d nlist=%d np=%lld)", dim, k, nlist, (long long)np);
        return PB200_ERR_UNSUPPORTED;
    }
    PB_REQUIRE(queries && probes && list_offsets && list_ids && list_vecs && xp && hxp && src_pos && tile_list &&
               xstats && out_dist && out_ids && workspace, "ivf_search_tc: null pointer");
    PB_REQUIRE(((uintptr_t)queries | (uintptr_t)xp | (uintptr_t)list_vecs | (uintptr_t)workspace) % 16 == 0,
               "ivf_search_tc: queries / xp / list_vecs / workspace must be 16-byte aligned");
    if (workspace_bytes < h.total) {
        set_error("ivf_search_tc: workspace %zu B < required %zu B", workspace_bytes, h.total);
        return PB200_ERR_WORKSPACE;
    }
    const tcs::Plan& pl = h.g;
    char* ws = static_cast<char*>(workspace);
    float* qr = reinterpret_cast<float*>(ws + h.off_qr);
    float* qn = reinterpret_cast<float*>(ws + h.off_qn);
    float* q_resid = reinterpret_cast<float*>(ws + h.off_qres);
    int32_t* qsel_count = reinterpret_cast<int32_t*>(ws + h.off_misc + 64);
    uint4* pmask = reinterpret_cast<uint4*>(ws + h.off_pmask);
    int32_t* qsel = reinterpret_cast<int32_t*>(ws + h.off_qsel);
    PB_CUDA(cudaMemsetAsync(ws + h.off_misc, 0, 256, stream));
    tcs::round_rows_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, stream>>>(queries, qr, nq, dim, q_resid, nullptr);
    int rc = check_launch("round_rows_kernel");
    if (rc) return rc;
    rc = row_sqnorm_run(queries, nq, dim, qn, stream);
    if (rc) return rc;
    {
        const int64_t blocks = ceil_div(nq, 256) < kSMs * 8 ? ceil_div(nq, 256) : kSMs * 8;
        tcs::probe_mask_kernel<<<(unsigned)blocks, 256, 0, stream>>>(probes, nq, nprobe, pmask);
        rc = check_launch("probe_mask_kernel");
        if (rc) return rc;
    }
    alignas(64) CUtensorMap tm_q, tm_x;
    if (!make_map(&tm_q, qr, nq, dim, tc::kTileM) || !make_map(&tm_x, xp, np, dim, tcs::kTileN)) {
        set_error("ivf_search_tc: cuTensorMapEncodeTiled failed (nq=%lld np=%lld dim=%d)", (long long)nq,
                  (long long)np, dim);
        return PB200_ERR_CUDA;
    }
    tcs::SearchParams sp{};
    sp.nq = nq; sp.nx = np; sp.d = dim; sp.nchunks = pl.nchunks; sp.qt = pl.qt; sp.ks = pl.ks;
    sp.splits = pl.splits; sp.split_len = pl.split_len; sp.stages = pl.stages; sp.align_slack = pl.align_slack;
    sp.scan_append = pl.scan_append;
    sp.kind = 0; sp.hx = hxp; sp.tile_list = tile_list; sp.pmask = reinterpret_cast<const uint32_t*>(pmask);
    sp.short_keys = reinterpret_cast<unsigned long long*>(ws + h.off_short);
    PB_CUDA(cudaMemsetAsync(sp.short_keys, 0, (size_t)pl.splits * nq * pl.ks * 8, stream));
    PB_CUDA(cudaFuncSetAttribute(tcs::search_tc_kernel<PB200_METRIC_L2>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    tcs::search_tc_kernel<PB200_METRIC_L2><<<pl.grid, tcs::kThreads, pl.smem_bytes, stream>>>(sp, tm_q, tm_x);
    rc = check_launch("search_tc_kernel");
    if (rc) return rc;

    tcs::RerankParams rp{};
    rp.q = queries; rp.x = list_vecs; rp.nq = nq; rp.nx = np; rp.d = dim; rp.k = k; rp.ks = pl.ks;
    rp.splits = pl.splits; rp.metric = PB200_METRIC_L2; rp.qn = qn; rp.xn = nullptr; rp.xmax_bits = xstats;
    rp.q_resid = q_resid; rp.exclude = nullptr; rp.id_offset = 0; rp.short_keys = sp.short_keys;
    rp.out_scores = out_dist; rp.out_ids = out_ids; rp.qsel = qsel; rp.qsel_count = qsel_count;
    rp.src_pos = src_pos; rp.ids_map = list_ids;
    tcs::rerank_kernel<<<(unsigned)ceil_div(nq, 8), 256, (size_t)8 * dim * 4, stream>>>(rp);
    rc = check_launch("rerank_kernel");
    if (rc) return rc;
    // uncertified queries: the list-scan kernel on the compacted list (device-side count); the first
    // cap_pp of them with one warp per (query, probe) + merge -- a handful of queries would otherwise
    // wait for one warp to walk all nprobe lists (6 ms at C4) -- the rest one warp per query
    float* part_bad = reinterpret_cast<float*>(ws + h.off_pbad);
    int32_t* part_ids = reinterpret_cast<int32_t*>(ws + h.off_pids);
    rc = ivf_search_run(queries, h.cap_pp, dim, probes, nprobe, list_offsets, list_ids, list_vecs, k, out_dist,
                        out_ids, qsel, qsel_count, 0, part_bad, part_ids, stream);
    if (rc) return rc;
    rc = topk_merge_run(part_bad, part_ids, h.cap_pp, nprobe * 32, 1, 0, k, out_dist, out_ids, qsel, qsel_count, 0,
                        stream);
    if (rc) return rc;
    if (nq > h.cap_pp) {
        rc = ivf_search_run(queries, nq - h.cap_pp, dim, probes, nprobe, list_offsets, list_ids, list_vecs, k,
                            out_dist, out_ids, qsel, qsel_count, h.cap_pp, nullptr, nullptr, stream);
        if (rc) return rc;
    }
    if (stats_out)
        PB_CUDA(cudaMemcpyAsync(stats_out, qsel_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    return PB200_OK;
}


// =====================================================================================
// LSH hash (LSHIndex.build / the query side of .search, utils/nearest_neighbors.py:35-43,
// 59-66; faiss IndexLSH: y = A x, bit j = y_j >= 0) on the tensor cores: Y = X A^T by the
// tcgen05 TF32 kernel of pb200_gather_dense, then one warp per vector packs the sign bits.
// Every |y| below the measured TF32 error bound is recomputed in fp32 with
// lsh_encode_kernel's own arithmetic (sequential fmaf over d), so the codes are bit-identical
// to pb200_lsh_encode (about 1 % of the projections at C3).
// =====================================================================================
namespace pb200 {
namespace tcs {

__global__ void __launch_bounds__(256) lsh_pack_kernel(const float* __restrict__ x, int64_t n, int d,
                                                       const float* __restrict__ proj, int nbits, int bit0,
                                                       int nb, const float* __restrict__ y,
                                                       const unsigned int* __restrict__ astats,
                                                       uint8_t* __restrict__ codes) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* xr = x + row * d;
    float n2 = 0.f, r2 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float v = __ldg(xr + c);
        const float e = __uint_as_float(to_tf32(v)) - v;
        n2 = fmaf(v, v, n2); r2 = fmaf(e, e, r2);
    }
    for (int o = 16; o > 0; o >>= 1) { n2 += __shfl_xor_sync(kFull, n2, o); r2 += __shfl_xor_sync(kFull, r2, o); }
    const float xnorm = sqrtf(n2) * 1.0001f, xres = sqrtf(r2) * 1.0001f;
    // |<x~,a~> - <x,a>| <= |x| max|a~ - a| + |x~ - x| max|a~|  + accumulation slack
    const float amax = __uint_as_float(astats[1]);
    const float eps = 1.02f * (xnorm * __uint_as_float(astats[0]) + xres * amax +
                               (float)d * 2.384185791015625e-07f * xnorm * amax);
    // lane owns byte `lane` of this 256-bit chunk: bits 8 lane .. 8 lane + 7
    if (lane * 8 >= nb) return;
    const float4 y0 = *reinterpret_cast<const float4*>(y + row * nb + lane * 8);
    const float4 y1 = *reinterpret_cast<const float4*>(y + row * nb + lane * 8 + 4);
    const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    uint32_t byte = 0u;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float v = yy[b];
        if (!(fabsf(v) >= eps)) {                      // uncertain (or NaN): exact fp32, reference order
            const float* a = proj + (int64_t)(bit0 + lane * 8 + b) * d;
            v = 0.f;
            for (int k = 0; k < d; ++k) v = fmaf(__ldg(a + k), __ldg(xr + k), v);
        }
        byte |= (v >= 0.f ? 1u : 0u) << b;
    }
    codes[row * (nbits / 8) + bit0 / 8 + lane] = (uint8_t)byte;
}

}  // namespace tcs
}  // namespace pb200

extern "C" size_t pb200_lsh_encode_tc_workspace_bytes(int64_t n, int dim, int nbits) {
    const int nb = nbits < 256 ? nbits : 256;
    return align_up((size_t)(nbits > 0 ? nbits : 1) * dim * 4, 256) + align_up((size_t)(n > 0 ? n : 1) * nb * 4, 256) + 256;
}

extern "C" int pb200_lsh_encode_tc(const float* x, int64_t n, int dim, const float* proj, int nbits,
                                   uint8_t* codes, void* workspace, size_t workspace_bytes,
                                   pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(n >= 0 && dim > 0 && nbits > 0 && nbits % 32 == 0,
               "lsh_encode_tc: nbits must be a positive multiple of 32");
    if (n == 0) return PB200_OK;
    PB_REQUIRE(x && proj && codes && workspace, "lsh_encode_tc: null pointer");
    if (dim % 4 || ((uintptr_t)x | (uintptr_t)proj | (uintptr_t)workspace) % 16) {
        set_error("lsh_encode_tc: needs dim %% 4 == 0 and 16-byte aligned x / proj / workspace (dim=%d)", dim);
        return PB200_ERR_UNSUPPORTED;
    }
    const size_t need = pb200_lsh_encode_tc_workspace_bytes(n, dim, nbits);
    if (workspace_bytes < need) {
        set_error("lsh_encode_tc: workspace %zu B < required %zu B", workspace_bytes, need);
        return PB200_ERR_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    float* proj_r = reinterpret_cast<float*>(ws);
    float* y = reinterpret_cast<float*>(ws + align_up((size_t)nbits * dim * 4, 256));
    unsigned int* astats = reinterpret_cast<unsigned int*>(ws + need - 256);   // [0] max |a~ - a|, [1] max |a~|
    PB_CUDA(cudaMemsetAsync(astats, 0, 256, stream));
    tcs::round_rows_kernel<<<(unsigned)ceil_div(nbits, 8), 256, 0, stream>>>(proj, proj_r, nbits, dim, nullptr, astats);
    int rc = check_launch("round_rows_kernel");
    if (rc) return rc;
    for (int bit0 = 0; bit0 < nbits; bit0 += 256) {
        const int nb = nbits - bit0 < 256 ? nbits - bit0 : 256;
        DenseParams p{};
        p.a1 = x; p.k1 = dim; p.a2 = nullptr; p.k2 = 0; p.pool_x = nullptr;
        p.lists = ListArgs{nullptr, nullptr, nullptr, nullptr, 1, 0, 0};
        p.w = proj_r + (size_t)bit0 * dim; p.bias = nullptr; p.ln_gamma = nullptr; p.ln_beta = nullptr;
        p.n = n; p.n_out = nb; p.flags = 0; p.out = y;
        if (!gather_dense_tf32_supported(p)) {
            set_error("lsh_encode_tc: projection shape not covered by the tensor-core kernel (dim=%d)", dim);
            return PB200_ERR_UNSUPPORTED;
        }
        rc = gather_dense_tf32(p, stream);
        if (rc) return rc;
        tcs::lsh_pack_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(x, n, dim, proj, nbits, bit0, nb, y, astats,
                                                                         codes);
        rc = check_launch("lsh_pack_kernel");
        if (rc) return rc;
    }
    return PB200_OK;
}
