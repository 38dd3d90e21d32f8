// dense_tc.cu -- tcgen05 (5th-gen tensor core, kind::tf32) path of pb200_gather_dense.
//
//   out[m,:] = epi( [A1[m,:K1] | A2row(m)] . W^T + bias ),  A2row dense or pooled on the fly
//
// One CTA per 128-row tile (UMMA M=128, N = n_out rounded up to 16 <= 256, K=8 per MMA),
// 13 warps, S-stage shared-memory ring of K-major 128B-swizzled operand tiles:
//   warps 0..7   A producers: [h | sum_j w_j h[id_j]] built from 16-byte vector loads (the
//                neighbour rows), rounded to TF32 (cvt.rna), written with the swizzle applied
//                by hand; a 3-deep register ring keeps two K chunks of loads in flight
//   warps 8..11  W producers: cp.async (16 B, L2-only) straight into the swizzled stage,
//                completion reported with cp.async.mbarrier.arrive.noinc; up to S chunks ahead
//   warp 12      one elected thread issues tcgen05.mma (accumulator in TMEM) and
//                tcgen05.commit's each stage back to the producers
//   epilogue     warps 0..3 read the accumulator once (tcgen05.ld, one output row per thread):
//                bias, ReLU, sum of squares -> staging tile in the idle stage buffers; all 12
//                producer warps then scale (row L2-norm) and store coalesced 16-byte vectors
// No TMA descriptor is needed: the smem matrix descriptors follow the canonical K-major
// SWIZZLE_128B layout (8 rows x 128 B atoms, SBO = 1024 B).  W should be pre-rounded to TF32
// (pb200_round_tf32): the tensor core ignores the 13 low mantissa bits of its operands.
#include "dense.cuh"

namespace pb200 {

namespace tc {

constexpr int kAWarps = 8, kBWarps = 4;
constexpr int kAThreads = kAWarps * 32;               // 256
constexpr int kBThreads = kBWarps * 32;               // 128
constexpr int kProducerWarps = kAWarps + kBWarps;     // 12
constexpr int kProducers = kProducerWarps * 32;       // 384
constexpr int kThreads = kProducers + 32;             // + MMA warp
constexpr int kTileM = 128;
constexpr int kChunkK = 32;                           // fp32 per 128-byte swizzle row
constexpr int kABytes = kTileM * 128;                 // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float f) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {   // arrives when this thread's copies land
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (Blackwell: version = 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address (16 B units)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8-row group pitch
    d |= (uint64_t)1 << 46;                          // descriptor version
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcGeom {
    int umma_n;      // n_out rounded up to 16
    int tmem_cols;   // power of two >= 32
    int stages;
    int b_bytes;     // umma_n * 128
    int stage_bytes;
    size_t smem_bytes;
    size_t off_lists, off_bias, off_bars;
};

__host__ __device__ inline TcGeom geometry(int n_out, int T, bool pooled) {
    TcGeom g{};
    g.umma_n = (n_out + 15) / 16 * 16;
    g.tmem_cols = 32;
    while (g.tmem_cols < g.umma_n) g.tmem_cols <<= 1;
    g.b_bytes = g.umma_n * 128;
    g.stage_bytes = kABytes + g.b_bytes;                        // multiples of 1024
    // raw copy of the tile's padded lists (ids, weights, 2 lengths) + compacted (id, weight, count)
    const size_t lists = pooled ? (size_t)kTileM * (12 + (size_t)T * 16) : 0;
    const size_t fixed = 1024 /*alignment slack*/ + lists + 256 * 4 /*bias*/ + 1024 /*barriers + row scales*/;
    int s = (int)((220 * 1024 - fixed) / g.stage_bytes);
    g.stages = s > 4 ? 4 : s;
    g.off_lists = (size_t)g.stages * g.stage_bytes;
    g.off_bias = g.off_lists + lists;
    g.off_bars = g.off_bias + 256 * 4;
    g.smem_bytes = g.off_bars + 1024 + 1024;
    return g;
}

__global__ void __launch_bounds__(kThreads, 1) dense_tc_kernel(const DenseParams p, const TcGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.lists.T;
    const bool pooled = p.pool_x != nullptr && p.k2 > 0;
    const int K = p.k1 + p.k2;
    const int nchunks = (K + kChunkK - 1) / kChunkK;
    const int64_t m0 = (int64_t)blockIdx.x * kTileM;

    int* s_nv = reinterpret_cast<int*>(smem + g.off_lists);             // [128]
    int* s_id = s_nv + kTileM;                                          // [128][T]
    float* s_w = reinterpret_cast<float*>(s_id + (size_t)kTileM * T);   // [128][T]
    int* r_len = reinterpret_cast<int*>(s_w + (size_t)kTileM * T);      // raw: [128] list_len
    int* r_wlen = r_len + kTileM;                                       //      [128] weight_len
    int* r_id = r_wlen + kTileM;                                        //      [128][T]
    float* r_w = reinterpret_cast<float*>(r_id + (size_t)kTileM * T);   //      [128][T]
    float* s_bias = reinterpret_cast<float*>(smem + g.off_bias);        // [256]
    const uint32_t bars = sbase + (uint32_t)g.off_bars;
    // barrier i: full[s] = bars + 8*s, empty[s] = bars + 8*(4+s), done = bars + 64, tmem ptr at +72
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.off_bars + 72);

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(bars + 8 * s, kAThreads + kBThreads);   // A: plain arrives, W: cp.async arrives
            mbar_init(bars + 8 * (4 + s), 1);
        }
        mbar_init(bars + 64, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarps) {   // MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 256) s_bias[tid] = (p.bias && tid < p.n_out) ? p.bias[tid] : 0.f;
    if (pooled && warp < kProducerWarps) {
        // the tile's padded lists are one contiguous block of global memory: coalesced copy
        const int64_t rows = min((int64_t)kTileM, p.n - m0);
        const int nent = (int)rows * T;
        for (int i = tid; i < nent; i += kProducers) {
            r_id[i] = p.lists.ids[m0 * T + i];
            r_w[i] = p.lists.weights ? p.lists.weights[m0 * T + i] : 0.f;
        }
        for (int i = tid; i < (int)rows; i += kProducers) {
            r_len[i] = p.lists.list_len ? p.lists.list_len[m0 + i] : T;
            r_wlen[i] = p.lists.weight_len ? p.lists.weight_len[m0 + i] : r_len[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (pooled && warp < kProducerWarps) {
        // filter / align / renormalise each row's list with the reference class's rule
        ListArgs la = p.lists;
        la.ids = r_id; la.weights = p.lists.weights ? r_w : nullptr; la.list_len = r_len; la.weight_len = r_wlen;
        for (int r = warp; r < kTileM; r += kProducerWarps) {
            int nv = 0;
            if (m0 + r < p.n) nv = prepare_list(la, r, s_id + r * T, s_w + r * T, lane);
            if (lane == 0) s_nv[r] = nv;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");   // producers only
    }

    if (warp < kAWarps) {
        // ===================== A producers =====================
        // thread constants: 4 (row, 16 B chunk) pairs; chunk j of row r lives at chunk j ^ (r & 7)
        const int j = tid & 7;
        int a_row[4]; uint32_t a_off[4]; bool a_ok[4];
        const float* a1p[4]; const float* a2p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (tid + kAThreads * i) >> 3;
            const int64_t m = m0 + r;
            a_row[i] = r;
            a_ok[i] = m < p.n;
            a_off[i] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
            a1p[i] = p.a1 + (a_ok[i] ? m : 0) * p.k1 + j * 4;
            a2p[i] = (pooled || !p.a2) ? nullptr : p.a2 + (a_ok[i] ? m : 0) * p.k2 + j * 4;
        }
        auto load_chunk = [&](int c, float4 (&av)[4]) {
            const int k = c * kChunkK + j * 4;          // first column of this thread's 16 B
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_ok[i] && k < K) {
                    if (k < p.k1) {
                        v = __ldg(reinterpret_cast<const float4*>(a1p[i] + c * kChunkK));
                    } else if (!pooled) {
                        v = __ldg(reinterpret_cast<const float4*>(a2p[i] + (c * kChunkK - p.k1)));
                    } else {
                        const int r = a_row[i], nv = s_nv[r], kk = k - p.k1;
                        const int* ids = s_id + r * T;
                        const float* ws = s_w + r * T;
#pragma unroll 2
                        for (int q = 0; q < nv; ++q) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(
                                p.pool_x + (int64_t)ids[q] * p.k2 + kk));
                            const float wgt = ws[q];
                            v.x = fmaf(wgt, t.x, v.x); v.y = fmaf(wgt, t.y, v.y);
                            v.z = fmaf(wgt, t.z, v.z); v.w = fmaf(wgt, t.w, v.w);
                        }
                    }
                }
                av[i] = v;
            }
        };
        auto store_chunk = [&](int c, const float4 (&av)[4]) {
            const int s = c % g.stages, u = c / g.stages;
            // the MMAs that read this stage last time must have completed
            if (u > 0) mbar_wait(bars + 8 * (4 + s), (uint32_t)((u - 1) & 1));
            const uint32_t st = sbase + (uint32_t)s * g.stage_bytes;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                st_shared_v4(st + a_off[i], to_tf32(av[i].x), to_tf32(av[i].y), to_tf32(av[i].z), to_tf32(av[i].w));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
            mbar_arrive(bars + 8 * s);
        };
        // 3-deep register ring: two chunks of loads are in flight while one is being stored
        float4 b0[4], b1[4], b2[4];
        load_chunk(0, b0);
        if (nchunks > 1) load_chunk(1, b1);
        for (int c = 0; c < nchunks; c += 3) {
            if (c + 2 < nchunks) load_chunk(c + 2, b2);
            store_chunk(c, b0);
            if (c + 1 < nchunks) {
                if (c + 3 < nchunks) load_chunk(c + 3, b0);
                store_chunk(c + 1, b1);
            }
            if (c + 2 < nchunks) {
                if (c + 4 < nchunks) load_chunk(c + 4, b1);
                store_chunk(c + 2, b2);
            }
        }
    } else if (warp < kProducerWarps) {
        // ===================== W producers (cp.async) =====================
        const int t = tid - kAThreads;                  // 0..127
        const int j = t & 7;
        const int b_iters = (g.umma_n * 8 + kBThreads - 1) / kBThreads;   // <= 16
        // rows n_out..umma_n-1 of the W tile are zero in every stage: written once
        for (int s = 0; s < g.stages; ++s)
            for (int i = 0; i < b_iters; ++i) {
                const int r = (t + kBThreads * i) >> 3;
                if (r >= p.n_out && r < g.umma_n)
                    st_shared_v4(sbase + (uint32_t)s * g.stage_bytes + kABytes +
                                 (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)), 0u, 0u, 0u, 0u);
            }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % g.stages, u = c / g.stages;
            if (u > 0) mbar_wait(bars + 8 * (4 + s), (uint32_t)((u - 1) & 1));
            const uint32_t st = sbase + (uint32_t)s * g.stage_bytes + kABytes;
            const int k = c * kChunkK + j * 4;
            const int rem = (K - k) * 4;
            const uint32_t nbytes = rem >= 16 ? 16u : (rem > 0 ? (uint32_t)rem : 0u);   // K tail: zero fill
            const float* src = p.w + (k < K ? k : 0);
#pragma unroll 4
            for (int i = 0; i < b_iters; ++i) {
                const int r = (t + kBThreads * i) >> 3;
                if (r < p.n_out)
                    cp_async16(st + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)),
                               src + (int64_t)r * K, nbytes);
            }
            cp_async_arrive(bars + 8 * s);
        }
    } else {
        // ===================== MMA issuer (one elected lane) =====================
        // instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.umma_n >> 3) << 17) |
                               ((uint32_t)(kTileM >> 4) << 24);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % g.stages, u = c / g.stages;
            mbar_wait(bars + 8 * s, (uint32_t)(u & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async data -> async proxy
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_s = sbase + (uint32_t)s * g.stage_bytes;
                const uint32_t b_s = a_s + kABytes;
#pragma unroll
                for (int k = 0; k < kChunkK / 8; ++k)   // 4 MMAs of K = 8 (32 bytes) per chunk
                    mma_tf32(tmem_base, make_desc(a_s + k * 32), make_desc(b_s + k * 32), idesc,
                             (uint32_t)((c | k) != 0));
                mma_commit(bars + 8 * (4 + s));          // frees the stage when the MMAs finish
                if (c == nchunks - 1) mma_commit(bars + 64);
            }
            __syncwarp();
        }
    }

    // ===================== epilogue =====================
    // warps 0..3 (TMEM lane = output row): one pass over the accumulator -> bias, ReLU, sum of
    // squares -> un-normalised row into the staging tile (the idle stage buffers, row stride
    // N+4 floats: conflict-free 16 B stores); then all 16 producer warps scale and write the
    // tile with coalesced 16 B stores.
    float* stg = reinterpret_cast<float*>(smem);
    float* s_scale = reinterpret_cast<float*>(smem + g.off_bars + 128);   // [128] after the barriers
    const int ldst = g.umma_n + 4;
    if (warp < 4) {
        mbar_wait(bars + 64, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool relu = p.flags & PB200_EPI_RELU;
        const int ncc = g.umma_n / 16;                   // 16-column pieces
        float ss = 0.f;
        uint32_t v[32];
        for (int cc = 0; cc < ncc; cc += 2) {
            tmem_ld32(taddr + cc * 16, v);               // 32 columns (reads past umma_n stay inside the allocation)
            const int ncol = min(32, g.umma_n - cc * 16);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                if (i < ncol) {
                    float f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = cc * 16 + i + e;
                        float x = __uint_as_float(v[i + e]) + s_bias[col & 255];
                        if (relu) x = fmaxf(x, 0.f);
                        if (col >= p.n_out) x = 0.f;
                        ss = fmaf(x, x, ss);
                        f[e] = x;
                    }
                    *reinterpret_cast<float4*>(stg + (size_t)row * ldst + cc * 16 + i) =
                        make_float4(f[0], f[1], f[2], f[3]);
                }
            }
        }
        s_scale[row] = (p.flags & PB200_EPI_L2NORM) ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;   // F.normalize eps
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp < kProducerWarps) {
        const bool vec_out = (p.n_out & 3) == 0 && ((uintptr_t)p.out & 15) == 0;
        const int n4 = (p.n_out + 3) >> 2;               // float4 pieces per row
        for (int idx = tid; idx < kTileM * n4; idx += kProducers) {
            const int r = idx / n4, c4 = (idx - r * n4) * 4;
            const int64_t m = m0 + r;
            if (m >= p.n) continue;
            float4 x = *reinterpret_cast<const float4*>(stg + (size_t)r * ldst + c4);
            const float sc = s_scale[r];
            x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
            float* o = p.out + m * p.n_out + c4;
            if (vec_out) {
                *reinterpret_cast<float4*>(o) = x;
            } else {
                o[0] = x.x;
                if (c4 + 1 < p.n_out) o[1] = x.y;
                if (c4 + 2 < p.n_out) o[2] = x.z;
                if (c4 + 3 < p.n_out) o[3] = x.w;
            }
        }
    }
    if (warp == kProducerWarps) {
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
    const tc::TcGeom g = tc::geometry(p.n_out, p.lists.T, pooled);
    // the epilogue stages the whole 128 x (N+4) fp32 tile in the pipeline buffers
    return g.stages >= 2 && (size_t)g.stages * g.stage_bytes >= (size_t)tc::kTileM * (g.umma_n + 4) * 4;
}

int gather_dense_tf32(const DenseParams& p, cudaStream_t stream) {
    if (!gather_dense_tf32_supported(p)) {
        set_error("gather_dense: PB200_PREC_TF32 needs n_out <= 256, k1 %% 4 == k2 %% 4 == 0, 16 B "
                  "aligned operands, no LayerNorm epilogue (got n_out=%d k1=%d k2=%d)", p.n_out,
                  p.k1, p.k2);
        return PB200_ERR_UNSUPPORTED;
    }
    const bool pooled = p.pool_x != nullptr && p.k2 > 0;
    const tc::TcGeom g = tc::geometry(p.n_out, p.lists.T, pooled);
    PB_CUDA(cudaFuncSetAttribute(tc::dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)g.smem_bytes));
    tc::dense_tc_kernel<<<(unsigned)ceil_div(p.n, tc::kTileM), tc::kThreads, g.smem_bytes, stream>>>(p, g);
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
