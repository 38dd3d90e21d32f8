// topk.cu -- K4: exact search = scoring GEMM fused with a streaming per-query top-k, and the
// (score, id) total-order merge used after the multi-GPU all-gather.
//
// Replaces utils/evaluation.py:119-130 / inference.py:112-118 (E1: q.E^T, self = -inf, topk)
// and faiss IndexFlatL2.search at utils/nearest_neighbors.py:174-181 (E2).
// CUDA-core fp32 scoring (exact-fp32 mode); score tiles never touch global memory: each
// 64x64 tile goes registers -> shared memory -> per-warp register-resident top-k lists
// (lane r holds rank r), so HBM/L2 traffic is the operands only.
// Items can be split across blockIdx.y; the per-split lists are merged by the same warp
// top-k under the (score, id) total order, which makes results tiling- and shard-invariant.
#include "topk.cuh"

namespace pb200 {

constexpr int TQ = 64, TX = 64, TK = 16;

__global__ void row_sqnorm_kernel(const float* __restrict__ x, int64_t n, int d, float* out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = x[row * d + c]; s = fmaf(v, v, s); }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) out[row] = s;
}

struct TopkParams {
    const float* __restrict__ q; int64_t nq;
    const float* __restrict__ x; int64_t nx;
    int d, k_pass, metric;
    const float* __restrict__ qn; const float* __restrict__ xn;   // squared norms (L2)
    const int32_t* __restrict__ exclude;
    int32_t id_offset;
    // floor = rank (col_off - 1) of the final output written by the previous pass
    const float* __restrict__ final_scores; const int32_t* __restrict__ final_ids;
    int k_total, col_off;
    int splits; int64_t split_len;
    float* __restrict__ part_bad; int32_t* __restrict__ part_ids;   // [nq, splits, 32]
    // optional query selection (pb200_topk_tc re-runs only its uncertified queries): slot i of
    // this launch is query qsel[qsel_base + i], for i < min(nq, *qsel_count - qsel_base)
    const int32_t* __restrict__ qsel; const int32_t* __restrict__ qsel_count; int64_t qsel_base;
};

__device__ __forceinline__ int64_t sel_count(const int32_t* qsel, const int32_t* cnt, int64_t base, int64_t nq) {
    if (!qsel) return nq;
    const int64_t c = (int64_t)*cnt - base;
    return c < nq ? (c < 0 ? 0 : c) : nq;
}

__global__ void __launch_bounds__(256) topk_tile_kernel(const TopkParams p) {
    __shared__ __align__(16) float q_s[TK][TQ + 4];
    __shared__ __align__(16) float x_s[TK][TX + 4];
    __shared__ float sc[TQ][TX + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t q0 = (int64_t)blockIdx.x * TQ;
    const int64_t xs = (int64_t)blockIdx.y * p.split_len;
    const int64_t xe = min(p.nx, xs + p.split_len);
    const int k = p.k_pass;
    const int64_t nq_eff = sel_count(p.qsel, p.qsel_count, p.qsel_base, p.nq);
    if (q0 >= nq_eff) return;

    TopkLane best[8];
    float fl_bad[8]; int fl_id[8]; float qn[8]; int excl[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        best[r].bad = INFINITY; best[r].id = INT_MAX;
        const int64_t qs = q0 + warp * 8 + r;
        fl_bad[r] = -INFINITY; fl_id[r] = -1; qn[r] = 0.f; excl[r] = -1;
        if (qs < nq_eff) {
            const int64_t qi = p.qsel ? p.qsel[p.qsel_base + qs] : qs;
            if (p.col_off > 0) {
                const float s = p.final_scores[qi * p.k_total + p.col_off - 1];
                fl_bad[r] = p.metric == PB200_METRIC_IP ? -s : s;
                fl_id[r] = p.final_ids[qi * p.k_total + p.col_off - 1];
                if (fl_id[r] < 0) { fl_bad[r] = INFINITY; fl_id[r] = INT_MAX; }  // list exhausted
            }
            if (p.metric == PB200_METRIC_L2) qn[r] = p.qn[qi];
            if (p.exclude) excl[r] = p.exclude[qi];
        }
    }

    for (int64_t x0 = xs; x0 < xe; x0 += TX) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < p.d; k0 += TK) {
            // 64 rows x 16 k per operand = 1024 elements each; 4 per thread
            {
                const int r = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kc = k0 + kk + i;
                    const int64_t qs = q0 + r, xi = x0 + r;
                    const int64_t qi = (p.qsel && qs < nq_eff) ? p.qsel[p.qsel_base + qs] : qs;
                    q_s[kk + i][r] = (qs < nq_eff && kc < p.d) ? __ldg(p.q + qi * p.d + kc) : 0.f;
                    x_s[kk + i][r] = (xi < xe && kc < p.d) ? __ldg(p.x + xi * p.d + kc) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&q_s[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&x_s[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[ty * 4 + i][tx * 4 + j] = acc[i][j];
        __syncthreads();
        // each warp scans its 8 query rows; lane covers items lane and lane+32 of the tile
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (q0 + warp * 8 + r >= nq_eff) break;  // warp-uniform
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = lane + 32 * h;
                const int64_t xi = x0 + col;
                const float dot = sc[warp * 8 + r][col];
                float bad;
                if (p.metric == PB200_METRIC_IP) bad = -dot;
                else bad = fmaxf(qn[r] + __ldg(p.xn + (xi < xe ? xi : xs)) - 2.f * dot, 0.f);
                const int gid = (int)xi + p.id_offset;
                bool valid = xi < xe && gid != excl[r] && bad == bad;
                valid = valid && better(fl_bad[r], fl_id[r], bad, gid);  // strictly after floor
                topk_offer(best[r], bad, gid, valid, k, lane);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t qi = q0 + warp * 8 + r;   // slot index: partial lists are slot-major
        if (qi >= nq_eff) break;
        const int64_t o = (qi * p.splits + blockIdx.y) * 32 + lane;
        p.part_bad[o] = best[r].bad;
        p.part_ids[o] = best[r].id == INT_MAX ? -1 : best[r].id;
    }
}

// One warp per query: merge c candidates (bad or score form) into ranks [col_off, col_off+k).
struct MergeParams {
    const float* __restrict__ vals; const int32_t* __restrict__ ids;   // [nq, c]
    int64_t nq; int c; int vals_are_bad; int largest;
    int k_pass, k_total, col_off, use_floor;
    float* __restrict__ out_scores; int32_t* __restrict__ out_ids;
    int out_hamming;  // unused here (kept for symmetry with lsh.cu)
    const int32_t* __restrict__ qsel; const int32_t* __restrict__ qsel_count; int64_t qsel_base;
};

__global__ void __launch_bounds__(256) topk_merge_kernel(const MergeParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // slot
    if (qi >= sel_count(p.qsel, p.qsel_count, p.qsel_base, p.nq)) return;
    const int64_t qo = p.qsel ? p.qsel[p.qsel_base + qi] : qi;                            // output row
    TopkLane e; e.bad = INFINITY; e.id = INT_MAX;
    float fb = -INFINITY; int fi = -1;
    if (p.use_floor && p.col_off > 0) {
        const float s = p.out_scores[qo * p.k_total + p.col_off - 1];
        fi = p.out_ids[qo * p.k_total + p.col_off - 1];
        fb = p.largest ? -s : s;
        if (fi < 0) { fb = INFINITY; fi = INT_MAX; }
    }
    for (int base = 0; base < p.c; base += 32) {
        const int j = base + lane;
        float bad = INFINITY; int id = -1;
        if (j < p.c) {
            const float v = p.vals[qi * p.c + j];
            id = p.ids[qi * p.c + j];
            bad = p.vals_are_bad ? v : (p.largest ? -v : v);
        }
        bool valid = j < p.c && id >= 0 && bad == bad && better(fb, fi, bad, id);
        topk_offer(e, bad, id, valid, p.k_pass, lane);
    }
    if (lane < p.k_pass) {
        const int64_t o = qo * p.k_total + p.col_off + lane;
        const bool has = e.id != INT_MAX;
        p.out_ids[o] = has ? e.id : -1;
        p.out_scores[o] = has ? (p.largest ? -e.bad : e.bad) : (p.largest ? -INFINITY : INFINITY);
    }
}

static int choose_splits(int64_t nq, int64_t nx) {
    const int64_t qblocks = ceil_div(nq, TQ);
    int64_t s = ceil_div(2 * kSMs, qblocks);
    const int64_t max_by_len = nx / 512 > 0 ? nx / 512 : 1;
    if (s > max_by_len) s = max_by_len;
    if (s > 64) s = 64;
    return (int)(s < 1 ? 1 : s);
}

// ---- N2: rank of a target item (hit-rate@k / MRR, utils/evaluation.py:5-73) ----
// rank[p] = 1 + #{j : s_j > s_gt  or  (s_j == s_gt and j < gt)},  s_j = <e[q_p], e[j]>, the
// position of the ground truth in the reference's descending sort.  tau is computed with the
// tile kernel's own arithmetic (sequential fmaf over d), so column gt compares equal to it.
__global__ void rank_tau_kernel(const float* __restrict__ e, int d, const int32_t* __restrict__ qid,
                                const int32_t* __restrict__ gid, int64_t np, float* tau, int32_t* rank) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x * blockDim.x) {
        const float* a = e + (int64_t)qid[p] * d;
        const float* b = e + (int64_t)gid[p] * d;
        float acc = 0.f;
        for (int k = 0; k < d; ++k) acc = fmaf(__ldg(a + k), __ldg(b + k), acc);
        tau[p] = acc;
        rank[p] = 1;
    }
}

__global__ void __launch_bounds__(256) rank_count_kernel(const float* __restrict__ e, int64_t n, int d,
                                                         const int32_t* __restrict__ qid,
                                                         const int32_t* __restrict__ gid, int64_t np,
                                                         const float* __restrict__ tau, int64_t split_len,
                                                         int32_t* rank) {
    __shared__ __align__(16) float q_s[TK][TQ + 4];
    __shared__ __align__(16) float x_s[TK][TX + 4];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t p0 = (int64_t)blockIdx.x * TQ;
    const int64_t xs = (int64_t)blockIdx.y * split_len, xe = min(n, xs + split_len);
    float t[4]; int g[4]; int cnt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = p0 + ty * 4 + i;
        t[i] = p < np ? tau[p] : INFINITY;
        g[i] = p < np ? gid[p] : -1;
        cnt[i] = 0;
    }
    for (int64_t x0 = xs; x0 < xe; x0 += TX) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < d; k0 += TK) {
            {
                const int r = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kc = k0 + kk + i;
                    const int64_t p = p0 + r, xi = x0 + r;
                    q_s[kk + i][r] = (p < np && kc < d) ? __ldg(e + (int64_t)qid[p] * d + kc) : 0.f;
                    x_s[kk + i][r] = (xi < xe && kc < d) ? __ldg(e + xi * d + kc) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&q_s[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&x_s[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t xi = x0 + tx * 4 + j;
                const float s = acc[i][j];
                cnt[i] += (xi < xe) && (s > t[i] || (s == t[i] && (int)xi < g[i]));
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = cnt[i];
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);   // the 16 threads of a row
        const int64_t p = p0 + ty * 4 + i;
        if (tx == 0 && p < np && c) atomicAdd(rank + p, c);
    }
}

int topk_merge_run(const float* vals, const int32_t* ids, int64_t nq, int c, int vals_are_bad, int largest,
                   int k, float* out_scores, int32_t* out_ids, const int32_t* qsel,
                   const int32_t* qsel_count, int64_t qsel_base, cudaStream_t stream) {
    MergeParams m{};
    m.vals = vals; m.ids = ids; m.nq = nq; m.c = c; m.vals_are_bad = vals_are_bad; m.largest = largest;
    m.k_pass = k; m.k_total = k; m.col_off = 0; m.use_floor = 0;
    m.out_scores = out_scores; m.out_ids = out_ids;
    m.qsel = qsel; m.qsel_count = qsel_count; m.qsel_base = qsel_base;
    topk_merge_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, stream>>>(m);
    return check_launch("topk_merge_kernel");
}

int row_sqnorm_run(const float* x, int64_t n, int d, float* out, cudaStream_t stream) {
    if (n <= 0) return PB200_OK;
    row_sqnorm_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, stream>>>(x, n, d, out);
    return check_launch("row_sqnorm_kernel");
}

// All passes (32 ranks each) of the fp32 tile kernel + merge for nq query slots.
int topk_fp32_run(const float* queries, int64_t nq, const float* items, int64_t nx, int dim, int k,
                  int metric, const float* qn, const float* xn, const int32_t* exclude_ids,
                  int32_t id_offset, float* out_scores, int32_t* out_ids, float* part_bad,
                  int32_t* part_ids, int splits, const int32_t* qsel, const int32_t* qsel_count,
                  int64_t qsel_base, cudaStream_t stream) {
    const int64_t split_len = ceil_div(ceil_div(nx > 0 ? nx : 1, splits), TX) * TX;
    for (int col = 0; col < k; col += 32) {
        TopkParams p{};
        p.q = queries; p.nq = nq; p.x = items; p.nx = nx; p.d = dim;
        p.k_pass = k - col < 32 ? k - col : 32; p.metric = metric; p.qn = qn; p.xn = xn;
        p.exclude = exclude_ids; p.id_offset = id_offset;
        p.final_scores = out_scores; p.final_ids = out_ids; p.k_total = k; p.col_off = col;
        p.splits = splits; p.split_len = split_len; p.part_bad = part_bad; p.part_ids = part_ids;
        p.qsel = qsel; p.qsel_count = qsel_count; p.qsel_base = qsel_base;
        dim3 grid((unsigned)ceil_div(nq, TQ), splits);
        topk_tile_kernel<<<grid, 256, 0, stream>>>(p);
        int rc = check_launch("topk_tile_kernel");
        if (rc) return rc;
        MergeParams m{};
        m.vals = part_bad; m.ids = part_ids; m.nq = nq; m.c = splits * 32; m.vals_are_bad = 1;
        m.largest = metric == PB200_METRIC_IP; m.k_pass = p.k_pass; m.k_total = k;
        m.col_off = col; m.use_floor = 0;  // partial lists are already beyond the floor
        m.out_scores = out_scores; m.out_ids = out_ids;
        m.qsel = qsel; m.qsel_count = qsel_count; m.qsel_base = qsel_base;
        topk_merge_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, stream>>>(m);
        rc = check_launch("topk_merge_kernel");
        if (rc) return rc;
    }
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_topk_workspace_bytes(int64_t nq, int64_t nx, int dim, int k) {
    (void)dim; (void)k;
    const int s = choose_splits(nq, nx);
    return align_up((size_t)(nq + nx) * 4, 256) + 2 * align_up((size_t)nq * s * 32 * 4, 256);
}

extern "C" int pb200_topk(const float* queries, int64_t nq, const float* items, int64_t nx,
                          int dim, int k, int metric, const int32_t* exclude_ids,
                          int32_t id_offset, float* out_scores, int32_t* out_ids, void* workspace,
                          size_t workspace_bytes, pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(nq >= 0 && nx >= 0 && dim > 0 && k > 0 && k <= 1024, "topk: bad sizes (k<=1024)");
    PB_REQUIRE(metric == PB200_METRIC_IP || metric == PB200_METRIC_L2, "topk: unknown metric");
    PB_REQUIRE(nx + (int64_t)id_offset < 2147483647ll, "topk: ids overflow int32");
    if (nq == 0) return PB200_OK;
    PB_REQUIRE(queries && out_scores && out_ids && workspace && (items || nx == 0),
               "topk: null pointer");
    const size_t need = pb200_topk_workspace_bytes(nq, nx, dim, k);
    if (workspace_bytes < need) {
        set_error("topk: workspace %zu B < required %zu B", workspace_bytes, need);
        return PB200_ERR_WORKSPACE;
    }
    const int splits = choose_splits(nq, nx);
    char* ws = static_cast<char*>(workspace);
    float* qn = reinterpret_cast<float*>(ws);
    float* xn = qn + nq;
    size_t off = align_up((size_t)(nq + nx) * 4, 256);
    float* part_bad = reinterpret_cast<float*>(ws + off);
    off += align_up((size_t)nq * splits * 32 * 4, 256);
    int32_t* part_ids = reinterpret_cast<int32_t*>(ws + off);
    if (metric == PB200_METRIC_L2) {
        row_sqnorm_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, stream>>>(queries, nq, dim, qn);
        int rc = check_launch("row_sqnorm_kernel");
        if (rc) return rc;
        if (nx > 0) {
            row_sqnorm_kernel<<<(unsigned)ceil_div(nx, 8), 256, 0, stream>>>(items, nx, dim, xn);
            rc = check_launch("row_sqnorm_kernel");
            if (rc) return rc;
        }
    }
    return topk_fp32_run(queries, nq, items, nx, dim, k, metric, qn, xn, exclude_ids, id_offset,
                         out_scores, out_ids, part_bad, part_ids, splits, nullptr, nullptr, 0, stream);
}

extern "C" int pb200_topk_merge(const float* scores, const int32_t* ids, int64_t nq, int c, int k,
                                int largest, float* out_scores, int32_t* out_ids,
                                pb200_stream_t stream) {
    PB_REQUIRE(nq >= 0 && c > 0 && k > 0 && k <= 1024, "topk_merge: bad sizes");
    if (nq == 0) return PB200_OK;
    PB_REQUIRE(scores && ids && out_scores && out_ids, "topk_merge: null pointer");
    for (int col = 0; col < k; col += 32) {
        MergeParams m{};
        m.vals = scores; m.ids = ids; m.nq = nq; m.c = c; m.vals_are_bad = 0;
        m.largest = largest != 0; m.k_pass = k - col < 32 ? k - col : 32; m.k_total = k;
        m.col_off = col; m.use_floor = 1; m.out_scores = out_scores; m.out_ids = out_ids;
        topk_merge_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, (cudaStream_t)stream>>>(m);
        int rc = check_launch("topk_merge_kernel");
        if (rc) return rc;
    }
    return PB200_OK;
}

extern "C" size_t pb200_rank_of_target_workspace_bytes(int64_t num_pairs) {
    return align_up((size_t)(num_pairs > 0 ? num_pairs : 1) * 4, 256);
}

extern "C" int pb200_rank_of_target(const float* embeddings, int64_t n, int dim, const int32_t* query_ids,
                                    const int32_t* target_ids, int64_t num_pairs, int32_t* out_rank,
                                    void* workspace, size_t workspace_bytes, pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(n > 0 && dim > 0 && num_pairs >= 0 && n < 2147483647ll, "rank_of_target: bad sizes");
    if (num_pairs == 0) return PB200_OK;
    PB_REQUIRE(embeddings && query_ids && target_ids && out_rank && workspace, "rank_of_target: null pointer");
    if (workspace_bytes < pb200_rank_of_target_workspace_bytes(num_pairs)) {
        set_error("rank_of_target: workspace too small");
        return PB200_ERR_WORKSPACE;
    }
    float* tau = static_cast<float*>(workspace);
    const int64_t tb = ceil_div(num_pairs, 256) < kSMs * 8 ? ceil_div(num_pairs, 256) : kSMs * 8;
    rank_tau_kernel<<<(unsigned)tb, 256, 0, stream>>>(embeddings, dim, query_ids, target_ids, num_pairs, tau, out_rank);
    int rc = check_launch("rank_tau_kernel");
    if (rc) return rc;
    const int splits = choose_splits(num_pairs, n);
    const int64_t split_len = ceil_div(ceil_div(n, splits), TX) * TX;
    dim3 grid((unsigned)ceil_div(num_pairs, TQ), splits);
    rank_count_kernel<<<grid, 256, 0, stream>>>(embeddings, n, dim, query_ids, target_ids, num_pairs, tau, split_len,
                                                out_rank);
    return check_launch("rank_count_kernel");
}
