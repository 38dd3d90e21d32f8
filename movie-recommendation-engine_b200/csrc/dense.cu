// dense.cu -- G1-G3: fused [gather -> importance-weighted sum -> concat -> dense -> epilogue],
// CUDA-core fp32 path (PB200_PREC_FP32).  The tcgen05 tensor-core path lives in
// dense_tc.cu; this one is the exact-fp32 reference kernel kept for bisecting and for shapes
// the tensor-core path does not cover.
//
// Replaces, per conv layer of PinSage.forward (reference model/pinsage.py:232-240):
//   h_neigh = importance_pooling(h, nbrs, wts); h_self = lin_self(h)   [folded by the host]
//   h = normalize(relu(lin_update(cat[h_self, h_neigh])))
// and the input/output projections (:202, :248-249), GraphConvLayer (model/layers.py:44-77)
// and ImportanceAggregator's Linear+LayerNorm (model/aggregators.py:258-283).
//
// Tile: 32 rows x BN columns per 256-thread block, K in chunks of 32.  The A tile is built in
// shared memory from A1 columns and, for the second half of K, either a dense A2 row or the
// pooled neighbourhood row computed on the fly from the per-block (id, weight) lists.
#include "dense.cuh"

namespace pb200 {

constexpr int BM = 32, BK = 32;

template <int BN, bool kVec>
__global__ void __launch_bounds__(256) dense_kernel(const DenseParams p) {
    constexpr int CN = BN / 32;        // columns per thread
    constexpr int AP = BM + 4;         // padded strides (keep 16 B alignment)
    constexpr int WP = BN + 4;
    extern __shared__ __align__(16) float smem_f[];
    float* a_s = smem_f;                       // [BK][AP]
    float* w_s = a_s + BK * AP;                // [BK][WP]
    const int T = p.lists.T;
    int* s_nv = reinterpret_cast<int*>(w_s + BK * WP);        // [BM]
    int* s_id = s_nv + BM;                                    // [BM][T]
    float* s_w = reinterpret_cast<float*>(s_id + BM * T);     // [BM][T]

    const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = p.k1 + p.k2;
    const bool pooled = p.pool_x != nullptr && p.k2 > 0;

    if (pooled) {  // each warp prepares 4 rows' lists
        for (int r = ty; r < BM; r += 8) {
            const int64_t m = m0 + r;
            int nv = 0;
            if (m < p.n) nv = prepare_list(p.lists, m, s_id + r * T, s_w + r * T, lane);
            if (lane == 0) s_nv[r] = nv;
        }
    }
    __syncthreads();

    float acc[4][CN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) acc[i][c] = 0.f;

    // thread's column c -> tile column
    auto tile_col = [&](int c) { return CN == 8 ? (c >> 2) * 128 + lane * 4 + (c & 3) : lane * CN + c; };

    for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- A tile: thread loads row ar, 4 consecutive k ----
        {
            const int ar = tid >> 3, kk = (tid & 7) * 4;
            const int64_t m = m0 + ar;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < p.n) {
                if (kVec) {
                    const int k = k0 + kk;
                    if (k < p.k1) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(p.a1 + m * p.k1 + k));
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    } else if (k < K) {
                        const int c = k - p.k1;
                        if (pooled) {
                            const int nv = s_nv[ar];
                            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                            for (int r = 0; r < nv; ++r) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(
                                    p.pool_x + (int64_t)s_id[ar * T + r] * p.k2 + c));
                                const float wgt = s_w[ar * T + r];
                                s.x = fmaf(wgt, t.x, s.x); s.y = fmaf(wgt, t.y, s.y);
                                s.z = fmaf(wgt, t.z, s.z); s.w = fmaf(wgt, t.w, s.w);
                            }
                            v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w;
                        } else {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(p.a2 + m * p.k2 + c));
                            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = k0 + kk + i;
                        if (k < p.k1) v[i] = __ldg(p.a1 + m * p.k1 + k);
                        else if (k < K) {
                            const int c = k - p.k1;
                            if (pooled) {
                                const int nv = s_nv[ar];
                                float s = 0.f;
                                for (int r = 0; r < nv; ++r)
                                    s = fmaf(s_w[ar * T + r],
                                             __ldg(p.pool_x + (int64_t)s_id[ar * T + r] * p.k2 + c), s);
                                v[i] = s;
                            } else v[i] = __ldg(p.a2 + m * p.k2 + c);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) a_s[(kk + i) * AP + ar] = v[i];
        }
        // ---- W tile: w_s[k][nn] = W[n0+nn][k0+k] ----
        for (int idx = tid; idx < BN * (BK / 4); idx += 256) {
            const int nn = idx >> 3, kk = (idx & 7) * 4;
            const int nrow = n0 + nn;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (nrow < p.n_out) {
                if (kVec && k0 + kk + 3 < K) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p.w + (int64_t)nrow * K + k0 + kk));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k0 + kk + i < K) v[i] = __ldg(p.w + (int64_t)nrow * K + k0 + kk + i);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) w_s[(kk + i) * WP + nn] = v[i];
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(a_s + k * AP + ty * 4);
            float wv[CN];
            if (CN == 8) {
                const float4 w0 = *reinterpret_cast<const float4*>(w_s + k * WP + lane * 4);
                const float4 w1 = *reinterpret_cast<const float4*>(w_s + k * WP + 128 + lane * 4);
                wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
                wv[4 % CN] = w1.x; wv[5 % CN] = w1.y; wv[6 % CN] = w1.z; wv[7 % CN] = w1.w;
            } else if (CN == 4) {
                const float4 w0 = *reinterpret_cast<const float4*>(w_s + k * WP + lane * 4);
                wv[0] = w0.x; wv[1] = w0.y; wv[2 % CN] = w0.z; wv[3 % CN] = w0.w;
            } else {
                const float2 w0 = *reinterpret_cast<const float2*>(w_s + k * WP + lane * 2);
                wv[0] = w0.x; wv[1 % CN] = w0.y;
            }
            const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CN; ++c) acc[i][c] = fmaf(a[i], wv[c], acc[i][c]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, ReLU, row norm (a warp owns all BN columns of its 4 rows) ----
    const bool relu = p.flags & PB200_EPI_RELU;
    const bool l2 = p.flags & PB200_EPI_L2NORM;
    const bool ln = p.flags & PB200_EPI_LAYERNORM;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        float ss = 0.f, sum = 0.f;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            const int col = n0 + tile_col(c);
            float v = acc[i][c];
            if (col < p.n_out) {
                if (p.bias) v += __ldg(p.bias + col);
                if (relu) v = fmaxf(v, 0.f);
            } else v = 0.f;
            acc[i][c] = v;
            ss = fmaf(v, v, ss);
            sum += v;
        }
        if (l2 || ln) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ss += __shfl_xor_sync(kFull, ss, o);
                sum += __shfl_xor_sync(kFull, sum, o);
            }
        }
        float scale = 1.f, shift = 0.f;
        if (l2) scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);          // F.normalize eps
        if (ln) {
            const float mean = sum / (float)p.n_out;
            float var = 0.f;
#pragma unroll
            for (int c = 0; c < CN; ++c)
                if (n0 + tile_col(c) < p.n_out) { const float d = acc[i][c] - mean; var = fmaf(d, d, var); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(kFull, var, o);
            scale = rsqrtf(var / (float)p.n_out + 1e-5f);
            shift = -mean * scale;
        }
        if (m < p.n) {
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const int col = n0 + tile_col(c);
                if (col < p.n_out) {
                    float v = fmaf(acc[i][c], scale, shift);
                    if (ln) v = fmaf(v, __ldg(p.ln_gamma + col), __ldg(p.ln_beta + col));
                    if (p.flags & PB200_EPI_ROUND_TF32) {
                        uint32_t r;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
                        v = __uint_as_float(r);
                    }
                    p.out[m * p.n_out + col] = v;
                }
            }
        }
    }
}

// in-place row normalisation for n_out > 256 (the fused epilogue needs the whole row)
__global__ void row_norm_kernel(float* out, int64_t n, int n_out, int flags,
                                const float* __restrict__ g, const float* __restrict__ b) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    float* o = out + row * n_out;
    float ss = 0.f, sum = 0.f;
    for (int c = lane; c < n_out; c += 32) { const float v = o[c]; ss = fmaf(v, v, ss); sum += v; }
    for (int k = 16; k > 0; k >>= 1) { ss += __shfl_xor_sync(kFull, ss, k); sum += __shfl_xor_sync(kFull, sum, k); }
    if (flags & PB200_EPI_L2NORM) {
        const float scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        for (int c = lane; c < n_out; c += 32) o[c] *= scale;
    } else {
        const float mean = sum / (float)n_out;
        float var = 0.f;
        for (int c = lane; c < n_out; c += 32) { const float d = o[c] - mean; var = fmaf(d, d, var); }
        for (int k = 16; k > 0; k >>= 1) var += __shfl_xor_sync(kFull, var, k);
        const float scale = rsqrtf(var / (float)n_out + 1e-5f);
        for (int c = lane; c < n_out; c += 32) o[c] = fmaf((o[c] - mean) * scale, g[c], b[c]);
    }
}

template <int BN>
static int launch_dense(const DenseParams& p, bool vec, int tiles_n, cudaStream_t stream) {
    const int T = p.lists.T;
    const size_t smem = sizeof(float) * (BK * (BM + 4) + BK * (BN + 4)) + sizeof(int) * BM +
                        (size_t)BM * T * 8;
    auto kern = vec ? dense_kernel<BN, true> : dense_kernel<BN, false>;
    if (smem > 200 * 1024) {
        set_error("gather_dense: max_neighbors=%d too large for shared memory", T);
        return PB200_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(p.n, BM), tiles_n);
    kern<<<grid, 256, smem, stream>>>(p);
    return check_launch("dense_kernel");
}

int gather_dense_fp32(const DenseParams& p0, cudaStream_t stream) {
    DenseParams p = p0;
    const int N = p.n_out;
    const bool norm = p.flags & (PB200_EPI_L2NORM | PB200_EPI_LAYERNORM);
    const bool vec = p.k1 % 4 == 0 && p.k2 % 4 == 0 && ((uintptr_t)p.a1 % 16 == 0) &&
                     ((uintptr_t)p.a2 % 16 == 0) && ((uintptr_t)p.pool_x % 16 == 0) &&
                     ((uintptr_t)p.w % 16 == 0);
    int rc;
    if (N <= 64) rc = launch_dense<64>(p, vec, 1, stream);
    else if (N <= 128) rc = launch_dense<128>(p, vec, 1, stream);
    else if (N <= 256) rc = launch_dense<256>(p, vec, 1, stream);
    else {
        p.flags &= ~(PB200_EPI_L2NORM | PB200_EPI_LAYERNORM);
        rc = launch_dense<256>(p, vec, (int)ceil_div(N, 256), stream);
        if (rc == PB200_OK && norm) {
            row_norm_kernel<<<(unsigned)ceil_div(p.n, 8), 256, 0, stream>>>(
                p.out, p.n, N, p0.flags, p.ln_gamma, p.ln_beta);
            rc = check_launch("row_norm_kernel");
        }
    }
    return rc;
}

}  // namespace pb200

using namespace pb200;


extern "C" int pb200_gather_dense(const float* a1, int k1, const float* a2, int k2,
                                  const float* pool_x, int64_t pool_rows, const int32_t* ids,
                                  const float* weights, const int32_t* list_len,
                                  const int32_t* weight_len, int max_neighbors, int pool_mode,
                                  const float* w, const float* bias, const float* ln_gamma,
                                  const float* ln_beta, int64_t n, int n_out, int flags,
                                  int precision, float* out, pb200_stream_t stream) {
    PB_REQUIRE(n >= 0 && n_out > 0 && k1 >= 0 && k2 >= 0 && k1 + k2 > 0, "gather_dense: bad sizes");
    PB_REQUIRE(!(flags & PB200_EPI_L2NORM) || !(flags & PB200_EPI_LAYERNORM),
               "gather_dense: L2NORM and LAYERNORM are exclusive");
    PB_REQUIRE(!(flags & PB200_EPI_LAYERNORM) || (ln_gamma && ln_beta),
               "gather_dense: LAYERNORM needs gamma and beta");
    if (n == 0) return PB200_OK;
    PB_REQUIRE(w && out && (k1 == 0 || a1), "gather_dense: null pointer");
    PB_REQUIRE(k2 == 0 || a2 || pool_x, "gather_dense: k2 > 0 needs a2 or pool_x");
    PB_REQUIRE(!pool_x || (ids && max_neighbors > 0 && pool_mode >= PB200_POOL_PINSAGE &&
                           pool_mode <= PB200_POOL_MEAN),
               "gather_dense: pooled input needs ids, max_neighbors > 0 and a weighted mode");
    DenseParams p{};
    p.a1 = a1; p.k1 = k1; p.a2 = pool_x ? nullptr : a2; p.k2 = k2; p.pool_x = k2 ? pool_x : nullptr;
    p.lists = ListArgs{ids, weights, list_len, weight_len, pool_x ? max_neighbors : 1, pool_mode,
                       pool_rows};
    p.w = w; p.bias = bias; p.ln_gamma = ln_gamma; p.ln_beta = ln_beta;
    p.n = n; p.n_out = n_out; p.flags = flags; p.out = out;
    if (precision == PB200_PREC_TF32 ||
        (precision == PB200_PREC_AUTO && n >= 64 && gather_dense_tf32_supported(p)))
        return gather_dense_tf32(p, (cudaStream_t)stream);
    PB_REQUIRE(precision == PB200_PREC_FP32 || precision == PB200_PREC_AUTO,
               "gather_dense: unknown precision %d", precision);
    return gather_dense_fp32(p, (cudaStream_t)stream);
}
