// csr_build.cu -- S0: edge list -> stable CSR + row-local cumulative weights.
//
// Replaces RandomWalkSampler._prepare_adjacency_list (reference utils/random_walk.py:33-50),
// which appends (dst, weight) to adj_list[src] edge by edge: the order inside a row is the
// edge order, and that order defines the CDF the walk samples from.  Here: a stable LSD
// radix sort of (src, edge index) -- cub::DeviceRadixSort, a library call, build-time only
// and off the per-step hot path -- followed by hand-written gather / row-boundary / prefix
// kernels.  All outputs land in caller-provided buffers.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace pb200 {

__global__ void probe_weights_kernel(const float* __restrict__ w, int64_t E, int32_t* flags) {
    int worst = 0, bad = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * blockDim.x) {
        const float v = w[e];
        if (!(v >= 0.0f) || isinf(v)) { ++bad; continue; }
        int s = 0;
        double q = (double)v;
        while (s < 11 && (q != floor(q) || q > 4294967295.0)) { q *= 2.0; ++s; }
        if (q > 4294967295.0) s = 11;
        worst = max(worst, s);
    }
    worst = __reduce_max_sync(kFull, worst);
    bad = __reduce_add_sync(kFull, bad);
    if ((threadIdx.x & 31) == 0) {
        if (worst) atomicMax(&flags[0], worst);
        if (bad) atomicAdd(&flags[1], bad);
    }
}

__global__ void csr_prep_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                int64_t E, int64_t N, uint32_t* keys, uint32_t* vals,
                                int32_t* status) {
    int bad = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src[e], d = dst[e];
        const bool ok = s >= 0 && s < N && d >= 0 && d < N;
        bad += !ok;
        keys[e] = ok ? (uint32_t)s : 0u;
        vals[e] = (uint32_t)e;
    }
    bad = __reduce_add_sync(kFull, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&status[0], bad);
}

// col, quantised weights and row boundaries from the sorted (src, edge) pairs
template <bool kQuant>
__global__ void csr_gather_kernel(const uint32_t* __restrict__ skeys,
                                  const uint32_t* __restrict__ svals,
                                  const int64_t* __restrict__ dst, const float* __restrict__ w,
                                  int64_t E, int64_t N, int quant_shift, int64_t* row_ptr,
                                  int32_t* col, uint64_t* q_out, double* wd_out,
                                  int32_t* status) {
    const double scale = (double)(1u << (quant_shift > 0 ? quant_shift : 0));
    int bad = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E;
         p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t key = skeys[p];
        const uint32_t e = svals[p];
        col[p] = (int32_t)dst[e];
        const float wv = w ? w[e] : 1.0f;
        if (kQuant) {
            const double q = (double)wv * scale;
            const bool ok = q >= 0.0 && q == floor(q) && q <= 4294967295.0;
            bad += !ok;
            q_out[p] = ok ? (uint64_t)q : 0ull;
        } else {
            bad += !(wv >= 0.0f);
            wd_out[p] = (double)wv;
        }
        // row boundaries: rows (prev, key] start at p
        const int64_t prev = p > 0 ? (int64_t)skeys[p - 1] : -1;
        for (int64_t v = prev + 1; v <= (int64_t)key; ++v) row_ptr[v] = p;
        if (p == E - 1)
            for (int64_t v = (int64_t)key + 1; v <= N; ++v) row_ptr[v] = E;
    }
    bad = __reduce_add_sync(kFull, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&status[1], bad);
}

__global__ void fill_row_ptr_kernel(int64_t* row_ptr, int64_t N, int64_t value) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= N;
         v += (int64_t)gridDim.x * blockDim.x)
        row_ptr[v] = value;
}

// cum[p] = G[p] - G[row_start - 1] (exact in uint64), narrowed to uint32
__global__ void csr_localize_kernel(const uint64_t* __restrict__ g,
                                    const uint32_t* __restrict__ skeys,
                                    const int64_t* __restrict__ row_ptr, int64_t E,
                                    uint32_t* cum, int32_t* status) {
    int over = 0, zero = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E;
         p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t row = skeys[p];
        const int64_t r0 = row_ptr[row];
        const uint64_t base = r0 > 0 ? g[r0 - 1] : 0ull;
        const uint64_t local = g[p] - base;
        over += local > 0xFFFFFFFFull;
        cum[p] = (uint32_t)local;
        if (p + 1 == row_ptr[row + 1]) zero += (local == 0);
    }
    over = __reduce_add_sync(kFull, over);
    zero = __reduce_add_sync(kFull, zero);
    if ((threadIdx.x & 31) == 0) {
        if (over) atomicAdd(&status[2], over);
        if (zero) atomicAdd(&status[3], zero);
    }
}

// generic weights: sequential float64 prefix per row (one thread per row; build-time only)
__global__ void csr_prefix_f64_kernel(const double* __restrict__ wd,
                                      const int64_t* __restrict__ row_ptr, int64_t N,
                                      double* cum, int32_t* status) {
    int zero = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < N;
         v += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        const int64_t a = row_ptr[v], b = row_ptr[v + 1];
        for (int64_t p = a; p < b; ++p) { acc += wd[p]; cum[p] = acc; }
        zero += (b > a && !(acc > 0.0));
    }
    zero = __reduce_add_sync(kFull, zero);
    if ((threadIdx.x & 31) == 0 && zero) atomicAdd(&status[3], zero);
}

struct CsrWorkspace {
    uint32_t *keys_in, *keys_out, *vals_in, *vals_out;
    uint64_t* g;  // quantised weights -> global inclusive prefix (in place); or double staging
    void* cub_temp;
    size_t cub_bytes, total;
};

static CsrWorkspace carve(void* base, int64_t E, int64_t N) {
    (void)N;
    CsrWorkspace w{};
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t)E, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (uint32_t)E);
    w.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
    const size_t e = (size_t)(E > 0 ? E : 1);
    w.keys_in = (uint32_t*)take(e * 4);
    w.keys_out = (uint32_t*)take(e * 4);
    w.vals_in = (uint32_t*)take(e * 4);
    w.vals_out = (uint32_t*)take(e * 4);
    w.g = (uint64_t*)take(e * 8);
    w.cub_temp = take(w.cub_bytes);
    w.total = off;
    return w;
}

static inline unsigned grid_for(int64_t n, int block = 256) {
    int64_t b = ceil_div(n, block);
    const int64_t cap = (int64_t)kSMs * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_edge_weight_probe(const float* edge_weights, int64_t num_edges,
                                       int32_t* flags_out, pb200_stream_t stream) {
    PB_REQUIRE(num_edges >= 0 && flags_out, "edge_weight_probe: bad arguments");
    if (num_edges == 0 || !edge_weights) return PB200_OK;
    probe_weights_kernel<<<grid_for(num_edges), 256, 0, (cudaStream_t)stream>>>(
        edge_weights, num_edges, flags_out);
    return check_launch("probe_weights_kernel");
}

extern "C" size_t pb200_csr_build_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
    return carve(nullptr, num_edges, num_nodes).total;
}

extern "C" int pb200_csr_build(const int64_t* edge_index, const float* edge_weights,
                               int64_t num_edges, int64_t num_nodes, int quant_shift,
                               int64_t* row_ptr, int32_t* col, void* cum, int32_t* status_out,
                               void* workspace, size_t workspace_bytes, pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t E = num_edges, N = num_nodes;
    PB_REQUIRE(E >= 0 && N >= 0 && E < 4294967296ll && N < 2147483647ll,
               "csr_build: E=%lld N=%lld out of range", (long long)E, (long long)N);
    PB_REQUIRE(quant_shift <= 10, "csr_build: quant_shift must be <= 10");
    PB_REQUIRE(row_ptr && status_out, "csr_build: null output");
    if (E == 0) {
        fill_row_ptr_kernel<<<grid_for(N + 1), 256, 0, stream>>>(row_ptr, N, 0);
        return check_launch("fill_row_ptr_kernel");
    }
    PB_REQUIRE(edge_index && col && cum && workspace, "csr_build: null pointer");
    CsrWorkspace ws = carve(workspace, E, N);
    if (workspace_bytes < ws.total) {
        set_error("csr_build: workspace %zu B < required %zu B", workspace_bytes, ws.total);
        return PB200_ERR_WORKSPACE;
    }
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + E;
    csr_prep_kernel<<<grid_for(E), 256, 0, stream>>>(src, dst, E, N, ws.keys_in, ws.vals_in,
                                                     status_out);
    int rc = check_launch("csr_prep_kernel");
    if (rc) return rc;
    int end_bit = 1;
    while (end_bit < 32 && (1ll << end_bit) < N) ++end_bit;
    size_t tb = ws.cub_bytes;
    PB_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_temp, tb, ws.keys_in, ws.keys_out, ws.vals_in,
                                            ws.vals_out, (uint32_t)E, 0, end_bit, stream));
    count_launch(4);
    const bool quant = quant_shift >= 0;
    if (quant)
        csr_gather_kernel<true><<<grid_for(E), 256, 0, stream>>>(
            ws.keys_out, ws.vals_out, dst, edge_weights, E, N, quant_shift, row_ptr, col, ws.g,
            nullptr, status_out);
    else
        csr_gather_kernel<false><<<grid_for(E), 256, 0, stream>>>(
            ws.keys_out, ws.vals_out, dst, edge_weights, E, N, 0, row_ptr, col, nullptr,
            reinterpret_cast<double*>(ws.g), status_out);
    rc = check_launch("csr_gather_kernel");
    if (rc) return rc;
    if (quant) {
        tb = ws.cub_bytes;
        PB_CUDA(cub::DeviceScan::InclusiveSum(ws.cub_temp, tb, ws.g, ws.g, (uint32_t)E, stream));
        count_launch(2);
        csr_localize_kernel<<<grid_for(E), 256, 0, stream>>>(ws.g, ws.keys_out, row_ptr, E,
                                                             static_cast<uint32_t*>(cum),
                                                             status_out);
        return check_launch("csr_localize_kernel");
    }
    csr_prefix_f64_kernel<<<grid_for(N), 256, 0, stream>>>(reinterpret_cast<double*>(ws.g),
                                                           row_ptr, N, static_cast<double*>(cum),
                                                           status_out);
    return check_launch("csr_prefix_f64_kernel");
}
