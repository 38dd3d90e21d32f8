// ivf.cu -- IVF-Flat ("Weak AND") index: inverted-list build, Lloyd centroid update, and the
// list-scan search kernel.
//
// Replaces WeakANDIndex.build/search (reference utils/nearest_neighbors.py:94-139), i.e.
// faiss.IndexIVFFlat(IndexFlatL2(d), d, nlist) with nprobe = min(nlist, 20): assign each
// vector to its nearest centroid, keep per-list (id, raw fp32 vector) in insertion order,
// and per query scan the nprobe nearest lists for the k smallest squared L2 distances
// (direct difference form, like faiss fvec_L2sqr); short results are padded with id -1.
// Centroid assignment and the quantizer top-nprobe reuse pb200_topk (topk.cu).
#include <cub/device/device_radix_sort.cuh>

#include "ivf.cuh"

namespace pb200 {

__global__ void ivf_iota_kernel(const int32_t* __restrict__ assign, int64_t n, uint32_t* keys,
                                uint32_t* vals) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = (uint32_t)assign[i];
        vals[i] = (uint32_t)i;
    }
}

// list_offsets from the sorted keys, list_ids, and the list-contiguous copy of the vectors
__global__ void ivf_scatter_kernel(const uint32_t* __restrict__ skeys,
                                   const uint32_t* __restrict__ svals, int64_t n, int nlist,
                                   const float* __restrict__ x, int d, int32_t* list_offsets,
                                   int32_t* list_ids, float* list_vecs) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = wid; p < n; p += nw) {
        const uint32_t key = skeys[p];
        const uint32_t src = svals[p];
        if (lane == 0) {
            list_ids[p] = (int32_t)src;
            const int64_t prev = p > 0 ? (int64_t)skeys[p - 1] : -1;
            for (int64_t l = prev + 1; l <= (int64_t)key; ++l) list_offsets[l] = (int32_t)p;
            if (p == n - 1)
                for (int64_t l = (int64_t)key + 1; l <= nlist; ++l) list_offsets[l] = (int32_t)n;
        }
        for (int c = lane; c < d; c += 32) list_vecs[p * d + c] = x[(int64_t)src * d + c];
    }
}

__global__ void ivf_fill_offsets_kernel(int32_t* list_offsets, int nlist) {
    for (int i = threadIdx.x; i <= nlist; i += blockDim.x) list_offsets[i] = 0;
}

// one block per list; thread c owns column c; members summed in list order (deterministic)
__global__ void ivf_centroid_kernel(const float* __restrict__ list_vecs,
                                    const int32_t* __restrict__ list_offsets, int d,
                                    float* centroids) {
    const int l = blockIdx.x;
    const int a = list_offsets[l], b = list_offsets[l + 1];
    if (b <= a) return;  // empty list keeps its centroid
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float s = 0.f;
        for (int p = a; p < b; ++p) s += list_vecs[(int64_t)p * d + c];
        centroids[(int64_t)l * d + c] = s / (float)(b - a);
    }
}

struct IvfSearchParams {
    const float* __restrict__ q; int64_t nq; int d;
    const int32_t* __restrict__ probes; int nprobe;
    const int32_t* __restrict__ list_offsets; const int32_t* __restrict__ list_ids;
    const float* __restrict__ list_vecs;
    int k;
    float* __restrict__ out_dist; int32_t* __restrict__ out_ids;
    // optional query selection (pb200_ivf_search_tc re-runs its uncertified queries here)
    const int32_t* __restrict__ qsel; const int32_t* __restrict__ qsel_count; int64_t qsel_base;
    // per-probe mode (few selected queries: one warp per (query, probe) instead of a 20-list
    // chain per warp); partial lists [slot][nprobe][32] are merged by topk_merge_kernel
    int per_probe; float* __restrict__ part_bad; int32_t* __restrict__ part_ids;
    // optional floor per query (k > 32 in passes of 32): only candidates strictly worse than (floor_dist, floor_id)
    // under the (distance, id) order are considered
    const float* __restrict__ floor_dist; const int32_t* __restrict__ floor_id;
};

// One warp per query.  Each warp stages 32 list vectors at a time in shared memory with
// coalesced row reads, then every lane scores one vector (stride d+1: conflict-free).
__global__ void __launch_bounds__(128) ivf_search_kernel(const IvfSearchParams p) {
    extern __shared__ float is[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = p.d, ds = d + 1;
    float* qs = is + (size_t)warp * (d + 32 * ds);   // [d]
    float* ts = qs + d;                              // [32][ds]
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t slot = p.per_probe ? w / p.nprobe : w;
    if (slot >= p.nq) return;
    int64_t qi = slot;
    if (p.qsel) {
        if (p.qsel_base + slot >= *p.qsel_count) return;
        qi = p.qsel[p.qsel_base + slot];
    }
    const int pi0 = p.per_probe ? (int)(w % p.nprobe) : 0;
    const int pi1 = p.per_probe ? pi0 + 1 : p.nprobe;
    for (int c = lane; c < d; c += 32) qs[c] = p.q[qi * d + c];
    TopkLane e; e.bad = INFINITY; e.id = INT_MAX;
    const bool floored = p.floor_dist != nullptr;
    const float fl_b = floored ? p.floor_dist[qi] : 0.f;
    const int fl_i = floored ? p.floor_id[qi] : 0;
    for (int pi = pi0; pi < pi1; ++pi) {
        const int l = p.probes[qi * p.nprobe + pi];
        if (l < 0) continue;  // warp-uniform
        const int a = p.list_offsets[l], b = p.list_offsets[l + 1];
        for (int base = a; base < b; base += 32) {
            const int cnt = min(32, b - base);
            __syncwarp();
            for (int r = 0; r < cnt; ++r)
                for (int c = lane; c < d; c += 32)
                    ts[r * ds + c] = __ldg(p.list_vecs + (int64_t)(base + r) * d + c);
            __syncwarp();
            float dist = 0.f;
            if (lane < cnt) {
                const float* v = ts + lane * ds;
#pragma unroll 4
                for (int c = 0; c < d; ++c) { const float t = qs[c] - v[c]; dist = fmaf(t, t, dist); }
            }
            const int id = lane < cnt ? p.list_ids[base + lane] : -1;
            topk_offer(e, dist, id, lane < cnt && (!floored || better(fl_b, fl_i, dist, id)), p.k, lane);
        }
    }
    if (p.per_probe) {
        const bool has = lane < p.k && e.id != INT_MAX;
        p.part_ids[w * 32 + lane] = has ? e.id : -1;
        p.part_bad[w * 32 + lane] = has ? e.bad : INFINITY;
        return;
    }
    if (lane < p.k) {
        const bool has = e.id != INT_MAX;
        p.out_ids[qi * p.k + lane] = has ? e.id : -1;
        p.out_dist[qi * p.k + lane] = has ? e.bad : INFINITY;
    }
}

struct IvfWs { uint32_t *k_in, *k_out, *v_in, *v_out; void* temp; size_t temp_bytes, total; };
static IvfWs ivf_carve(void* base, int64_t n) {
    IvfWs w{};
    cub::DeviceRadixSort::SortPairs(nullptr, w.temp_bytes, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t)n, 0, 32);
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
    const size_t e = (size_t)(n > 0 ? n : 1);
    w.k_in = (uint32_t*)take(e * 4); w.k_out = (uint32_t*)take(e * 4);
    w.v_in = (uint32_t*)take(e * 4); w.v_out = (uint32_t*)take(e * 4);
    w.temp = take(w.temp_bytes);
    w.total = off;
    return w;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_ivf_build_workspace_bytes(int64_t n, int nlist) {
    (void)nlist;
    return ivf_carve(nullptr, n).total;
}

extern "C" int pb200_ivf_build(const float* x, int64_t n, int dim, const int32_t* assign, int nlist,
                               int32_t* list_offsets, int32_t* list_ids, float* list_vecs,
                               void* workspace, size_t workspace_bytes, pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(n >= 0 && n < 2147483647ll && dim > 0 && nlist > 0, "ivf_build: bad sizes");
    PB_REQUIRE(list_offsets, "ivf_build: null pointer");
    if (n == 0) {
        ivf_fill_offsets_kernel<<<1, 256, 0, stream>>>(list_offsets, nlist);
        return check_launch("ivf_fill_offsets_kernel");
    }
    PB_REQUIRE(x && assign && list_ids && list_vecs && workspace, "ivf_build: null pointer");
    IvfWs w = ivf_carve(workspace, n);
    if (workspace_bytes < w.total) {
        set_error("ivf_build: workspace %zu B < required %zu B", workspace_bytes, w.total);
        return PB200_ERR_WORKSPACE;
    }
    const unsigned blocks = (unsigned)(ceil_div(n, 256) < kSMs * 8 ? ceil_div(n, 256) : kSMs * 8);
    ivf_iota_kernel<<<blocks, 256, 0, stream>>>(assign, n, w.k_in, w.v_in);
    int rc = check_launch("ivf_iota_kernel");
    if (rc) return rc;
    int end_bit = 1;
    while (end_bit < 32 && (1ll << end_bit) < nlist) ++end_bit;
    size_t tb = w.temp_bytes;
    // stable LSD sort: ids stay ascending inside a list, like faiss's append order
    PB_CUDA(cub::DeviceRadixSort::SortPairs(w.temp, tb, w.k_in, w.k_out, w.v_in, w.v_out,
                                            (uint32_t)n, 0, end_bit, stream));
    count_launch(3);
    const unsigned wblocks = (unsigned)(ceil_div(n, 8) < kSMs * 8 ? ceil_div(n, 8) : kSMs * 8);
    ivf_scatter_kernel<<<wblocks, 256, 0, stream>>>(w.k_out, w.v_out, n, nlist, x, dim,
                                                    list_offsets, list_ids, list_vecs);
    return check_launch("ivf_scatter_kernel");
}

extern "C" int pb200_ivf_centroid_update(const float* list_vecs, const int32_t* list_offsets,
                                         int nlist, int dim, float* centroids,
                                         pb200_stream_t stream) {
    PB_REQUIRE(list_vecs && list_offsets && centroids && nlist > 0 && dim > 0,
               "ivf_centroid_update: bad arguments");
    ivf_centroid_kernel<<<nlist, 128, 0, (cudaStream_t)stream>>>(list_vecs, list_offsets, dim, centroids);
    return check_launch("ivf_centroid_kernel");
}

namespace pb200 {
int ivf_search_run(const float* queries, int64_t nq, int dim, const int32_t* probes, int nprobe,
                   const int32_t* list_offsets, const int32_t* list_ids, const float* list_vecs, int k,
                   float* out_dist, int32_t* out_ids, const int32_t* qsel, const int32_t* qsel_count,
                   int64_t qsel_base, float* part_bad, int32_t* part_ids, cudaStream_t stream,
                   const float* floor_dist, const int32_t* floor_id) {
    const int per_probe = part_bad != nullptr;
    IvfSearchParams p{queries, nq, dim, probes, nprobe, list_offsets, list_ids, list_vecs, k,
                      out_dist, out_ids, qsel, qsel_count, qsel_base, per_probe, part_bad, part_ids,
                      floor_dist, floor_id};
    const int wpb = 4;
    const size_t smem = (size_t)wpb * (dim + 32 * (dim + 1)) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("ivf_search: dim=%d too large for the shared-memory tile", dim);
        return PB200_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(ivf_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    const int64_t warps = per_probe ? nq * nprobe : nq;
    ivf_search_kernel<<<(unsigned)ceil_div(warps, wpb), wpb * 32, smem, stream>>>(p);
    return check_launch("ivf_search_kernel");
}
}  // namespace pb200

extern "C" int pb200_ivf_search_ex(const float* queries, int64_t nq, int dim, const int32_t* probes,
                                   int nprobe, const int32_t* list_offsets, const int32_t* list_ids,
                                   const float* list_vecs, int k, const float* floor_dist, const int32_t* floor_ids,
                                   float* out_dist, int32_t* out_ids, pb200_stream_t stream) {
    PB_REQUIRE(nq >= 0 && dim > 0 && nprobe > 0, "ivf_search: bad sizes");
    PB_REQUIRE(k > 0 && k <= 32, "ivf_search: k must be in [1, 32] per pass (larger k: passes with a floor)");
    PB_REQUIRE((floor_dist == nullptr) == (floor_ids == nullptr), "ivf_search: floor needs both distance and id");
    if (nq == 0) return PB200_OK;
    PB_REQUIRE(queries && probes && list_offsets && list_ids && list_vecs && out_dist && out_ids,
               "ivf_search: null pointer");
    return ivf_search_run(queries, nq, dim, probes, nprobe, list_offsets, list_ids, list_vecs, k, out_dist,
                          out_ids, nullptr, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream, floor_dist, floor_ids);
}

extern "C" int pb200_ivf_search(const float* queries, int64_t nq, int dim, const int32_t* probes,
                                int nprobe, const int32_t* list_offsets, const int32_t* list_ids,
                                const float* list_vecs, int k, float* out_dist, int32_t* out_ids,
                                pb200_stream_t stream) {
    return pb200_ivf_search_ex(queries, nq, dim, probes, nprobe, list_offsets, list_ids, list_vecs, k, nullptr, nullptr,
                               out_dist, out_ids, stream);
}
