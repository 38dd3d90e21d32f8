// dense_tc.cu -- tcgen05 (5th-gen tensor core, kind::tf32) path of pb200_gather_dense.
//
//   out[m,:] = epi( [A1[m,:K1] | A2row(m)] . W^T + bias ),  A2row dense or pooled on the fly
//
// Persistent kernel: one CTA per SM loops over 128-row tiles (UMMA M=128, N = n_out rounded up
// to 16 <= 256, K=8 per MMA).  16 warps, S-stage shared-memory ring of K-major 128B-swizzled
// operand tiles, TWO accumulators in TMEM so the epilogue of tile i overlaps the mainloop of
// tile i+1:
//   warps 0..7    "pool" warps: the K chunks that need arithmetic -- the pooled neighbourhood
//                 sum_j w_j h[id_j] (16-byte vector loads of the neighbour rows, two neighbours
//                 per element in flight, next chunk issued before the current one is reduced)
//                 and any operand that still has to be rounded to TF32 -- written with the
//                 swizzle applied by hand; they also compact the tile's neighbour lists with
//                 the reference class's filter / align / renormalise rule
//   warp 8        TMA producer: one elected thread issues cp.async.bulk.tensor (SWIZZLE_128B
//                 tensor maps) for the W tile and for TF32-ready A1 columns of every K chunk,
//                 completion via mbarrier expect_tx / complete_tx, up to S chunks ahead
//   warp 9        prefetches the next tile's raw neighbour lists (cp.async)
//   warp 10       one elected thread issues tcgen05.mma and tcgen05.commit's stages back to
//                 the producers and finished accumulators to the epilogue
//   warps 12..15  epilogue: tcgen05.ld (one output row per thread), bias, ReLU, row L2-norm
//                 from registers, optional rounding to TF32 (so the next layer can cp.async
//                 it), staged through a small shared tile for coalesced 16 B stores
// No TMA descriptor is needed: the smem matrix descriptors follow the canonical K-major
// SWIZZLE_128B layout (8 rows x 128 B atoms, SBO = 1024 B).  W should be pre-rounded to TF32
// (pb200_round_tf32): the tensor core ignores the 13 low mantissa bits of its operands.
#include <cstdlib>

#include "dense.cuh"
#include "tc_common.cuh"

namespace pb200 {

namespace tc {

// ---- optional cycle accounting (make EXTRA=-DPB200_TC_PROFILE; tools/prof_dense.py) ----
#ifdef PB200_TC_PROFILE
__device__ unsigned long long g_prof[148][12];
#define PROF_ADD(slot, t_begin) do { if (lane == 0) atomicAdd(&g_prof[blockIdx.x % 148][slot], (unsigned long long)(clock64() - (t_begin))); } while (0)
#define PROF_NOW() clock64()
#else
#define PROF_ADD(slot, t_begin) do { } while (0)
#define PROF_NOW() 0ll
#endif


constexpr int kPoolWarps = 8, kCopyWarps = 2;       // warp 8: TMA producer, warp 9: list prefetch
constexpr int kPoolThreads = kPoolWarps * 32;         // 256
constexpr int kCopyThreads = kCopyWarps * 32;         // 64
constexpr int kMmaWarp = kPoolWarps + kCopyWarps;     // 10 (warp 11 idle: keeps warp & 3 == TMEM lane quarter)
constexpr int kEpiWarp0 = 12;                         // 12..15
constexpr int kEpiThreads = 128;
constexpr int kThreads = (kEpiWarp0 + 4) * 32;        // 512
constexpr int kEpiStageFloats = 32 * 36;              // per epilogue warp: 32 rows x (32+4) floats

struct TcGeom {
    int umma_n;       // n_out rounded up to 16
    int tmem_cols;    // power of two >= 2 * umma_n (two accumulators)
    int stages;
    int stage_bytes;  // A tile (16 KB) + W tile (umma_n * 128 B)
    int list_bytes;   // one compacted list buffer: nv[128] + id[128][T] + w[128][T]
    size_t smem_bytes;
    size_t off_epi, off_lists, off_raw, off_bias, off_bars;
    int debug;        // PB200_TC_VARIANT (tuning experiments)
};

__host__ __device__ inline TcGeom geometry(int n_out, int T, bool pooled) {
    TcGeom g{};
    g.umma_n = (n_out + 15) / 16 * 16;
    g.tmem_cols = 32;
    while (g.tmem_cols < 2 * g.umma_n) g.tmem_cols <<= 1;
    g.stage_bytes = kABytes + g.umma_n * 128;                   // multiples of 1024
    g.list_bytes = pooled ? kTileM * (4 + T * 8) : 0;
    const size_t raw = pooled ? (size_t)kTileM * (8 + (size_t)T * 8) : 0;
    const size_t fixed = 1024 /*alignment slack*/ + 4 * kEpiStageFloats * 4 + 2 * (size_t)g.list_bytes +
                         raw + 256 * 4 /*bias*/ + 256 /*barriers*/;
    int s = (int)((225 * 1024 - fixed) / g.stage_bytes);
    g.stages = s > 4 ? 4 : s;
    g.off_epi = (size_t)g.stages * g.stage_bytes;
    g.off_lists = g.off_epi + 4 * kEpiStageFloats * 4;
    g.off_raw = g.off_lists + 2 * (size_t)g.list_bytes;
    g.off_bias = g.off_raw + raw;
    g.off_bars = g.off_bias + 256 * 4;
    g.smem_bytes = g.off_bars + 256 + 1024;
    return g;
}

// barrier slots (8 B each) inside the barrier block
enum { kBarFullC = 0, kBarFullP = 4, kBarEmpty = 8, kBarTFull = 12, kBarTEmpty = 14, kBarRawReady = 16,
       kBarRawFree = 17, kTmemSlot = 20 * 8 };

struct PoolRegs { float4 t[4][2]; };   // two neighbour rows in flight per (row, 16 B chunk) slot

__global__ void __launch_bounds__(kThreads, 1) dense_tc_kernel(const DenseParams p, const TcGeom g,
                                                               const __grid_constant__ CUtensorMap tm_w,
                                                               const __grid_constant__ CUtensorMap tm_a,
                                                               const __grid_constant__ CUtensorMap tm_a2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.lists.T;
    const bool pooled = p.pool_x != nullptr && p.k2 > 0;
    const int K = p.k1 + p.k2;
    const int nchunks = (K + kChunkK - 1) / kChunkK;
    const int64_t ntiles = (p.n + kTileM - 1) / kTileM;
    // K chunks [0, first_reg) hold TF32-ready A1 columns only: the copy warps cp.async them.
    // Chunks [first_reg, nchunks) need registers (pooling, or rounding to TF32).
    const bool a1_tma = (p.flags & PB200_IN_A1_TF32) && p.k1 >= kChunkK && (p.k1 % kChunkK) == 0;
    const bool a2_tma = a1_tma && (p.flags & PB200_IN_A2_TF32) && !pooled && p.a2 && (p.k2 % kChunkK) == 0;
    const int first_reg = a2_tma ? nchunks : (a1_tma ? p.k1 / kChunkK : 0);

    float* s_bias = reinterpret_cast<float*>(smem + g.off_bias);        // [256]
    const uint32_t bars = sbase + (uint32_t)g.off_bars;
    auto bar = [&](int slot) { return bars + 8u * (uint32_t)slot; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.off_bars + kTmemSlot);
    auto list_nv = [&](int b) { return reinterpret_cast<int*>(smem + g.off_lists + (size_t)b * g.list_bytes); };
    int* r_len = reinterpret_cast<int*>(smem + g.off_raw);              // raw copy of one tile's lists
    int* r_wlen = r_len + kTileM;
    int* r_id = r_wlen + kTileM;
    float* r_w = reinterpret_cast<float*>(r_id + (size_t)kTileM * T);

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(bar(kBarFullC + s), 1);              // TMA producer: expect_tx + complete_tx
            mbar_init(bar(kBarFullP + s), kPoolThreads);   // plain arrives
            mbar_init(bar(kBarEmpty + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar(kBarTFull + a), 1);
            mbar_init(bar(kBarTEmpty + a), kEpiThreads);
        }
        mbar_init(bar(kBarRawReady), 32);              // list-prefetch warp: cp.async arrives
        mbar_init(bar(kBarRawFree), kPoolThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {   // MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 256) s_bias[tid] = (p.bias && tid < p.n_out) ? p.bias[tid] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kPoolWarps) {
        // ===================== pool warps =====================
        const int j = tid & 7;    // 16-byte chunk inside the 128 B row; lives at chunk j ^ (r & 7)
        const int r0 = tid >> 3;  // this thread's rows are r0, r0+32, r0+64, r0+96
        uint32_t a_off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + 32 * i;
            a_off[i] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
        }
        int gc = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it, gc += nchunks) {
            const int64_t m0 = tile * kTileM;
            const int rows_left = (int)min((int64_t)kTileM, p.n - m0);
            int* s_nv = list_nv(it & 1);
            int* s_id = s_nv + kTileM;
            float* s_w = reinterpret_cast<float*>(s_id + (size_t)kTileM * T);
            const long long t_c = PROF_NOW();
            if (pooled) {
                // raw lists were prefetched by the copy warps: compact them (reference rule)
                mbar_wait(bar(kBarRawReady), (uint32_t)(it & 1));
                ListArgs la = p.lists;
                la.ids = r_id; la.weights = p.lists.weights ? r_w : nullptr; la.list_len = r_len; la.weight_len = r_wlen;
                for (int r = warp; r < kTileM; r += kPoolWarps) {
                    int nv = 0;
                    if (r < rows_left) nv = prepare_list(la, r, s_id + r * T, s_w + r * T, lane);
                    if (lane == 0) s_nv[r] = nv;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kPoolThreads) : "memory");
                mbar_arrive(bar(kBarRawFree));       // the raw buffer may be refilled for the next tile
            }
            if (warp == 0) PROF_ADD(0, t_c);         // pool: wait raw lists + compaction
            const float* a1b = p.a1 + (m0 + r0) * p.k1 + j * 4;
            const float* a2b = (pooled || !p.a2) ? nullptr : p.a2 + (m0 + r0) * p.k2 + j * 4;

            // issue the loads of chunk c: up to two source rows per slot
            auto issue = [&](int c, PoolRegs& E) {
                const long long t_i = PROF_NOW();
                const int k = c * kChunkK + j * 4;          // first column of this thread's 16 B
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    E.t[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    E.t[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r0 + 32 * i < rows_left && k < K) {
                        if (k < p.k1) {
                            E.t[i][0] = __ldg(reinterpret_cast<const float4*>(a1b + (int64_t)(32 * i) * p.k1 + c * kChunkK));
                        } else if (!pooled) {
                            E.t[i][0] = __ldg(reinterpret_cast<const float4*>(a2b + (int64_t)(32 * i) * p.k2 + (c * kChunkK - p.k1)));
                        } else {
                            const int r = r0 + 32 * i, nv = s_nv[r], kk = k - p.k1;
                            if (nv > 0) E.t[i][0] = __ldg(reinterpret_cast<const float4*>(p.pool_x + (int64_t)s_id[r * T] * p.k2 + kk));
                            if (nv > 1) E.t[i][1] = __ldg(reinterpret_cast<const float4*>(p.pool_x + (int64_t)s_id[r * T + 1] * p.k2 + kk));
                        }
                    }
                }
                if (warp == 0) PROF_ADD(1, t_i);     // pool: issuing loads
            };
            // reduce chunk c and hand it to the tensor core
            auto finish = [&](int c, const PoolRegs& E) {
                const long long t_f = PROF_NOW();
                const int k = c * kChunkK + j * 4;
                const bool pool_chunk = pooled && k >= p.k1;
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!pool_chunk) { v[i] = E.t[i][0]; continue; }
                    const int r = r0 + 32 * i;
                    const int nv = r < rows_left ? s_nv[r] : 0;
                    const float w0 = nv > 0 ? s_w[r * T] : 0.f, w1 = nv > 1 ? s_w[r * T + 1] : 0.f;
                    float4 a;
                    a.x = fmaf(w1, E.t[i][1].x, w0 * E.t[i][0].x); a.y = fmaf(w1, E.t[i][1].y, w0 * E.t[i][0].y);
                    a.z = fmaf(w1, E.t[i][1].z, w0 * E.t[i][0].z); a.w = fmaf(w1, E.t[i][1].w, w0 * E.t[i][0].w);
                    for (int q = 2; q < nv; ++q) {    // rare: more than two valid neighbours
                        const float4 t = __ldg(reinterpret_cast<const float4*>(
                            p.pool_x + (int64_t)s_id[r * T + q] * p.k2 + (k - p.k1)));
                        const float wq = s_w[r * T + q];
                        a.x = fmaf(wq, t.x, a.x); a.y = fmaf(wq, t.y, a.y);
                        a.z = fmaf(wq, t.z, a.z); a.w = fmaf(wq, t.w, a.w);
                    }
                    v[i] = a;
                }
                const int g_c = gc + c;
                const int s = g_c % g.stages, u = g_c / g.stages;
                if (warp == 0) PROF_ADD(2, t_f);     // pool: waiting for data + reduce
                const long long t_w = PROF_NOW();
                // the MMAs that read this stage last time must have completed
                if (u > 0) mbar_wait(bar(kBarEmpty + s), (uint32_t)((u - 1) & 1));
                if (warp == 0) PROF_ADD(3, t_w);     // pool: waiting for a free stage
                const uint32_t st = sbase + (uint32_t)s * g.stage_bytes;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(st + a_off[i], to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
                mbar_arrive(bar(kBarFullP + s));
            };
            // Chunks the copy warps fill alone: step through their ring slots in order (mbarrier
            // parity waits are only meaningful one phase apart, so no slot may be skipped).
            auto pass = [&](int c) {
                const long long t_p = PROF_NOW();
                const int g_c = gc + c;
                const int s = g_c % g.stages, u = g_c / g.stages;
                if (u > 0) mbar_wait(bar(kBarEmpty + s), (uint32_t)((u - 1) & 1));
                mbar_arrive(bar(kBarFullP + s));
                if (warp == 0) PROF_ADD(4, t_p);     // pool: stepping through copy-only chunks
            };
            PoolRegs e0, e1;
            if (first_reg < nchunks) issue(first_reg, e0);   // gathers start while the MMAs run on A1
            for (int c = 0; c < first_reg && c < nchunks; ++c) pass(c);
            for (int c = first_reg; c < nchunks; c += 2) {
                if (c + 1 < nchunks) issue(c + 1, e1);
                finish(c, e0);
                if (c + 1 < nchunks) {
                    if (c + 2 < nchunks) issue(c + 2, e0);
                    finish(c + 1, e1);
                }
            }
        }
    } else if (warp == kPoolWarps) {
        // ===================== TMA producer (one elected lane) =====================
        // W rows n_out..umma_n-1 and K-tail columns are outside the tensor: TMA zero-fills them
        const uint32_t w_bytes = (uint32_t)g.umma_n * 128u, a_bytes = (uint32_t)kABytes;
        int gc = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row0 = (int)(tile * kTileM);
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int s = gc % g.stages, u = gc / g.stages;
                const long long t_e = PROF_NOW();
                if (u > 0) mbar_wait(bar(kBarEmpty + s), (uint32_t)((u - 1) & 1));
                PROF_ADD(5, t_e);                    // copy: waiting for a free stage
                const long long t_q = PROF_NOW();
                if (lane == 0) {
                    const uint32_t sa = sbase + (uint32_t)s * g.stage_bytes;
                    const bool with_a = c < first_reg;   // TF32-ready A1 columns ride along
                    mbar_expect_tx(bar(kBarFullC + s), w_bytes + (with_a ? a_bytes : 0u));
                    tma_load_2d(sa + kABytes, &tm_w, c * kChunkK, 0, bar(kBarFullC + s));
                    if (with_a) {
                        if (c * kChunkK < p.k1) tma_load_2d(sa, &tm_a, c * kChunkK, row0, bar(kBarFullC + s));
                        else tma_load_2d(sa, &tm_a2, c * kChunkK - p.k1, row0, bar(kBarFullC + s));
                    }
                }
                __syncwarp();
                PROF_ADD(6, t_q);                    // copy: issuing TMA
            }
        }
    } else if (warp == kPoolWarps + 1) {
        // ===================== list prefetch warp =====================
        if (pooled) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                // the pool warps must have compacted the previous contents
                if (it >= 1) mbar_wait(bar(kBarRawFree), (uint32_t)((it - 1) & 1));
                const int64_t m0 = tile * kTileM;
                const int rows = (int)min((int64_t)kTileM, p.n - m0);
                // the tile's padded lists are one contiguous block of global memory
                for (int i = lane; i < rows * T; i += 32) {
                    cp_async4(smem_u32(r_id + i), p.lists.ids + m0 * T + i);
                    if (p.lists.weights) cp_async4(smem_u32(r_w + i), p.lists.weights + m0 * T + i);
                }
                for (int i = lane; i < rows; i += 32) {
                    if (p.lists.list_len) cp_async4(smem_u32(r_len + i), p.lists.list_len + m0 + i);
                    else r_len[i] = T;
                    if (p.lists.weight_len) cp_async4(smem_u32(r_wlen + i), p.lists.weight_len + m0 + i);
                    else if (p.lists.list_len) cp_async4(smem_u32(r_wlen + i), p.lists.list_len + m0 + i);
                    else r_wlen[i] = T;
                }
                cp_async_arrive(bar(kBarRawReady));
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one elected lane) =====================
        // instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.umma_n >> 3) << 17) |
                               ((uint32_t)(kTileM >> 4) << 24);
        int gc = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, ua = it >> 1;
            // the epilogue must have drained this accumulator (two tiles ago)
            if (ua > 0) mbar_wait(bar(kBarTEmpty + acc), (uint32_t)((ua - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * g.umma_n);
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int s = gc % g.stages, u = gc / g.stages;
                const long long t_a = PROF_NOW();
                mbar_wait(bar(kBarFullC + s), (uint32_t)(u & 1));   // TMA data landed
                PROF_ADD(7, t_a);                                   // mma: waiting for cp.async data
                const long long t_b = PROF_NOW();
                mbar_wait(bar(kBarFullP + s), (uint32_t)(u & 1));   // pool warps
                PROF_ADD(8, t_b);                                   // mma: waiting for pool data
                const long long t_m = PROF_NOW();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t a_s = sbase + (uint32_t)s * g.stage_bytes;
                    const uint32_t b_s = a_s + kABytes;
#pragma unroll
                    for (int k = 0; k < kChunkK / 8; ++k)   // 4 MMAs of K = 8 (32 bytes) per chunk
                        mma_tf32(tmem_d, make_desc(a_s + k * 32), make_desc(b_s + k * 32), idesc,
                                 (uint32_t)((c | k) != 0));
                    mma_commit(bar(kBarEmpty + s));          // frees the stage when the MMAs finish
                    if (c == nchunks - 1) mma_commit(bar(kBarTFull + acc));
                }
                __syncwarp();
                PROF_ADD(9, t_m);                                   // mma: fence + issue + commit
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue warps (TMEM lane quarter = warp & 3) =====================
        const int quarter = warp & 3;
        float* stg = reinterpret_cast<float*>(smem + g.off_epi) + (warp - kEpiWarp0) * kEpiStageFloats;
        const bool relu = p.flags & PB200_EPI_RELU;
        const bool l2 = p.flags & PB200_EPI_L2NORM;
        const bool round_out = p.flags & PB200_EPI_ROUND_TF32;
        const bool vec_out = (p.n_out & 3) == 0 && ((uintptr_t)p.out & 15) == 0;
        const int ncc = (g.umma_n + 31) / 32;            // 32-column pieces
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, ua = it >> 1;
            const int64_t m0 = tile * kTileM;
            const long long t_x = PROF_NOW();
            mbar_wait(bar(kBarTFull + acc), (uint32_t)(ua & 1));
            if (warp == kEpiWarp0) PROF_ADD(10, t_x);              // epilogue: waiting for an accumulator
            const long long t_y = PROF_NOW();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * g.umma_n);
            uint32_t v[32];
            float scale = 1.f;
            if (l2) {
                float ss = 0.f;
                for (int cc = 0; cc < ncc; ++cc) {
                    tmem_ld32(taddr + cc * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + ((cc * 32 + i) & 255));
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float f = __uint_as_float(v[i + e]) + bb[e];
                            if (relu) f = fmaxf(f, 0.f);
                            if (cc * 32 + i + e < p.n_out) ss = fmaf(f, f, ss);
                        }
                    }
                }
                scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);      // F.normalize eps
            }
            for (int cc = 0; cc < ncc; ++cc) {
                tmem_ld32(taddr + cc * 32, v);
                if (cc == ncc - 1) {   // accumulator fully read: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(bar(kBarTEmpty + acc));
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + ((cc * 32 + i) & 255));
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                    float f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = __uint_as_float(v[i + e]) + bb[e];
                        if (relu) x = fmaxf(x, 0.f);
                        x *= scale;
                        f[e] = round_out ? __uint_as_float(to_tf32(x)) : x;
                    }
                    *reinterpret_cast<float4*>(stg + lane * 36 + i) = make_float4(f[0], f[1], f[2], f[3]);
                }
                __syncwarp();
                // 8 lanes cover one row's 128 B; 4 rows per instruction
#pragma unroll
                for (int rr0 = 0; rr0 < 32; rr0 += 4) {
                    const int rr = rr0 + (lane >> 3), c4 = (lane & 7) * 4;
                    const int64_t m = m0 + quarter * 32 + rr;
                    const int col = cc * 32 + c4;
                    if (m < p.n && col < p.n_out) {
                        const float4 x = *reinterpret_cast<const float4*>(stg + rr * 36 + c4);
                        float* o = p.out + m * p.n_out + col;
                        if (vec_out) {
                            *reinterpret_cast<float4*>(o) = x;
                        } else {
                            o[0] = x.x;
                            if (col + 1 < p.n_out) o[1] = x.y;
                            if (col + 2 < p.n_out) o[2] = x.z;
                            if (col + 3 < p.n_out) o[3] = x.w;
                        }
                    }
                }
                __syncwarp();
            }
            if (warp == kEpiWarp0) PROF_ADD(11, t_y);              // epilogue: both passes
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                     ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
}

}  // namespace tc

bool gather_dense_tf32_supported(const DenseParams& p) {
    const bool pooled = p.pool_x != nullptr && p.k2 > 0;
    if (p.n_out > 256 || (p.flags & PB200_EPI_LAYERNORM)) return false;
    if (p.k1 % 4 || p.k2 % 4) return false;
    if (((uintptr_t)p.a1 | (uintptr_t)p.a2 | (uintptr_t)p.pool_x | (uintptr_t)p.w) % 16) return false;
    return tc::geometry(p.n_out, p.lists.T, pooled).stages >= 2;
}

int gather_dense_tf32(const DenseParams& p, cudaStream_t stream) {
    if (!gather_dense_tf32_supported(p)) {
        set_error("gather_dense: PB200_PREC_TF32 needs n_out <= 256, k1 %% 4 == k2 %% 4 == 0, 16 B "
                  "aligned operands, no LayerNorm epilogue (got n_out=%d k1=%d k2=%d)", p.n_out,
                  p.k1, p.k2);
        return PB200_ERR_UNSUPPORTED;
    }
    const bool pooled = p.pool_x != nullptr && p.k2 > 0;
    tc::TcGeom g = tc::geometry(p.n_out, p.lists.T, pooled);
    static const int dbg = [] { const char* e = getenv("PB200_TC_VARIANT"); return e ? atoi(e) : 0; }();
    g.debug = dbg;
    alignas(64) CUtensorMap tm_w, tm_a, tm_a2;
    const int K = p.k1 + p.k2;
    if (!make_map(&tm_w, p.w, p.n_out, K, g.umma_n)) {
        set_error("gather_dense: cuTensorMapEncodeTiled failed for W [%d, %d]", p.n_out, K);
        return PB200_ERR_CUDA;
    }
    // A1 is streamed by TMA only when it is TF32-ready; otherwise the map is a (valid) dummy
    const bool a_tma = (p.flags & PB200_IN_A1_TF32) && p.k1 >= tc::kChunkK && p.k1 % tc::kChunkK == 0;
    if (!(a_tma ? make_map(&tm_a, p.a1, p.n, p.k1, tc::kTileM) : make_map(&tm_a, p.w, p.n_out, K, g.umma_n))) {
        set_error("gather_dense: cuTensorMapEncodeTiled failed for A1 [%lld, %d]", (long long)p.n, p.k1);
        return PB200_ERR_CUDA;
    }
    const bool a2_tma = a_tma && (p.flags & PB200_IN_A2_TF32) && !pooled && p.a2 && p.k2 % tc::kChunkK == 0;
    if (!(a2_tma ? make_map(&tm_a2, p.a2, p.n, p.k2, tc::kTileM) : make_map(&tm_a2, p.w, p.n_out, K, g.umma_n))) {
        set_error("gather_dense: cuTensorMapEncodeTiled failed for A2 [%lld, %d]", (long long)p.n, p.k2);
        return PB200_ERR_CUDA;
    }
    PB_CUDA(cudaFuncSetAttribute(tc::dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)g.smem_bytes));
    const int64_t ntiles = ceil_div(p.n, tc::kTileM);
    const unsigned grid = (unsigned)(ntiles < kSMs ? ntiles : kSMs);   // persistent: one CTA per SM
    tc::dense_tc_kernel<<<grid, tc::kThreads, g.smem_bytes, stream>>>(p, g, tm_w, tm_a, tm_a2);
    return check_launch("dense_tc_kernel");
}

}  // namespace pb200

namespace pb200 {
__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __uint_as_float(tc::to_tf32(in[i]));
}
}  // namespace pb200

extern "C" int pb200_round_tf32(const float* in, float* out, int64_t n, pb200_stream_t stream) {
    using namespace pb200;
    PB_REQUIRE(n >= 0 && (n == 0 || (in && out)), "round_tf32: bad arguments");
    if (n == 0) return PB200_OK;
    const int64_t blocks = ceil_div(n, 256) < kSMs * 8 ? ceil_div(n, 256) : kSMs * 8;
    round_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
    return check_launch("round_tf32_kernel");
}

#ifdef PB200_TC_PROFILE
extern "C" int pb200_debug_tc_profile(unsigned long long* out_host, int reset) {
    cudaDeviceSynchronize();
    if (out_host) cudaMemcpyFromSymbol(out_host, pb200::tc::g_prof, sizeof(unsigned long long) * 148 * 12);
    if (reset) { static unsigned long long z[148 * 12]; cudaMemcpyToSymbol(pb200::tc::g_prof, z, sizeof(z)); }
    return 0;
}
#endif
