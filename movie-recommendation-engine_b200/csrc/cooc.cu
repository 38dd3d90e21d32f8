// cooc.cu -- N4: item-item co-occurrence graph (reference data/graph_builder.py:59-116).
//
// The reference counts, in a python dict, every pair of movies that a user rated together (all users,
// all position pairs i < j of the user's rows) and keeps the pairs seen at least `threshold` times, in
// dict order = order of FIRST co-occurrence (users in ascending userId order, position pairs
// lexicographic).  Here: one thread block per movie a walks users(a) x items(user) with dense per-block
// accumulators {count[b], first user[b]} in L2/HBM -- sum_u deg(u)^2 atomic increments in total, the same
// work as the reference's loop but ~10^10 of them per second -- emits every pair (a < b) that reaches the
// threshold with the ordering key (first user, min position, max position), and the pairs are radix
// sorted by that key (cub, build-time only) into the reference's edge order: (lo -> hi), (hi -> lo) per pair.
// Precondition (checked by the host mirror): a user rates a movie at most once.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pb200 {

struct CoocArgs {
    const int64_t* __restrict__ urow_ptr;  // [U+1] user -> rows (ratings-table order inside a user)
    const int32_t* __restrict__ uitems;    // [R]
    const int64_t* __restrict__ irow_ptr;  // [M+1] movie -> its users in ascending user rank
    const int32_t* __restrict__ iusers;    // [R]
    const int32_t* __restrict__ ipos;      // [R] position of the movie inside that user's list
    int64_t M;
    int threshold, bits_p;
    int32_t* acc_cnt;                      // [gridDim.x, M]
    int32_t* acc_first;                    // [gridDim.x, M]
    unsigned long long* out_key;
    int32_t* out_a; int32_t* out_b; int32_t* out_cnt;
    unsigned long long capacity;
    unsigned long long* out_count;
};

__global__ void __launch_bounds__(256) cooc_kernel(const CoocArgs p) {
    int32_t* cnt = p.acc_cnt + (size_t)blockIdx.x * p.M;
    int32_t* first = p.acc_first + (size_t)blockIdx.x * p.M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int64_t i = threadIdx.x; i < p.M; i += blockDim.x) { cnt[i] = 0; first[i] = 0x7FFFFFFF; }
    __syncthreads();
    for (int64_t a = blockIdx.x; a < p.M; a += gridDim.x) {
        const int64_t e0 = p.irow_ptr[a], e1 = p.irow_ptr[a + 1];
        for (int64_t e = e0 + warp; e < e1; e += nwarp) {          // pass 1: count, first user
            const int u = p.iusers[e];
            const int64_t t0 = p.urow_ptr[u], t1 = p.urow_ptr[u + 1];
            for (int64_t t = t0 + lane; t < t1; t += 32) {
                const int b = p.uitems[t];
                if (b > a) { atomicAdd(&cnt[b], 1); atomicMin(&first[b], u); }
            }
        }
        __syncthreads();
        for (int64_t e = e0 + warp; e < e1; e += nwarp) {          // pass 2: emit at the first co-occurrence
            const int u = p.iusers[e];
            const uint32_t pa = (uint32_t)p.ipos[e];
            const int64_t t0 = p.urow_ptr[u], t1 = p.urow_ptr[u + 1];
            for (int64_t t = t0 + lane; t < t1; t += 32) {
                const int b = p.uitems[t];
                if (b > a && cnt[b] >= p.threshold && first[b] == u) {
                    const uint32_t pb = (uint32_t)(t - t0);
                    const unsigned long long lo = pa < pb ? pa : pb, hi = pa < pb ? pb : pa;
                    const unsigned long long idx = atomicAdd(p.out_count, 1ull);
                    if (idx < p.capacity) {
                        p.out_key[idx] = ((unsigned long long)u << (2 * p.bits_p)) | (lo << p.bits_p) | hi;
                        p.out_a[idx] = (int32_t)a; p.out_b[idx] = b; p.out_cnt[idx] = cnt[b];
                    }
                }
            }
        }
        __syncthreads();
        for (int64_t e = e0 + warp; e < e1; e += nwarp) {          // pass 3: reset what this row touched
            const int u = p.iusers[e];
            const int64_t t0 = p.urow_ptr[u], t1 = p.urow_ptr[u + 1];
            for (int64_t t = t0 + lane; t < t1; t += 32) {
                const int b = p.uitems[t];
                if (b > a) { cnt[b] = 0; first[b] = 0x7FFFFFFF; }
            }
        }
        __syncthreads();
    }
}

__global__ void cooc_iota_kernel(uint32_t* v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = (uint32_t)i;
}

// pair k (in key order) -> edges 2k: (a -> b), 2k+1: (b -> a); a < b
__global__ void cooc_edges_kernel(const uint32_t* __restrict__ order, const int32_t* __restrict__ a,
                                  const int32_t* __restrict__ b, const int32_t* __restrict__ cnt, int64_t P,
                                  int64_t* edge_index, float* edge_weight) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = order[k];
        edge_index[2 * k] = a[r]; edge_index[2 * k + 1] = b[r];
        edge_index[2 * P + 2 * k] = b[r]; edge_index[2 * P + 2 * k + 1] = a[r];
        edge_weight[2 * k] = (float)cnt[r]; edge_weight[2 * k + 1] = (float)cnt[r];
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_cooc_pairs(const int64_t* urow_ptr, const int32_t* uitems, const int64_t* irow_ptr,
                                const int32_t* iusers, const int32_t* ipos, int64_t num_users, int64_t num_items,
                                int threshold, int bits_pos, int32_t* acc_cnt, int32_t* acc_first, int num_blocks,
                                uint64_t* out_key, int32_t* out_a, int32_t* out_b, int32_t* out_cnt,
                                uint64_t capacity, uint64_t* out_count, pb200_stream_t stream) {
    PB_REQUIRE(num_users >= 0 && num_items > 0 && threshold >= 1 && bits_pos >= 1 && num_blocks >= 1,
               "cooc_pairs: bad sizes");
    int bits_u = 1;
    while (bits_u < 63 && (1ll << bits_u) < num_users) ++bits_u;
    PB_REQUIRE(bits_u + 2 * bits_pos <= 64, "cooc_pairs: %d user bits + 2 x %d position bits exceed the 64-bit order key",
               bits_u, bits_pos);
    PB_REQUIRE(urow_ptr && uitems && irow_ptr && iusers && ipos && acc_cnt && acc_first && out_key && out_a && out_b &&
               out_cnt && out_count, "cooc_pairs: null pointer");
    PB_CUDA(cudaMemsetAsync(out_count, 0, sizeof(uint64_t), (cudaStream_t)stream));
    CoocArgs p{urow_ptr, uitems, irow_ptr, iusers, ipos, num_items, threshold, bits_pos, acc_cnt, acc_first,
               reinterpret_cast<unsigned long long*>(out_key), out_a, out_b, out_cnt, (unsigned long long)capacity,
               reinterpret_cast<unsigned long long*>(out_count)};
    cooc_kernel<<<(unsigned)num_blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("cooc_kernel");
}

extern "C" size_t pb200_cooc_edges_workspace_bytes(int64_t num_pairs) {
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(num_pairs > 0 ? num_pairs : 1));
    const size_t n = (size_t)(num_pairs > 0 ? num_pairs : 1);
    return align_up(sort_bytes, 256) + align_up(n * 8, 256) + 2 * align_up(n * 4, 256);
}

extern "C" int pb200_cooc_edges(const uint64_t* keys, const int32_t* a, const int32_t* b, const int32_t* cnt,
                                int64_t num_pairs, int64_t* edge_index, float* edge_weight, void* workspace,
                                size_t workspace_bytes, pb200_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PB_REQUIRE(num_pairs >= 0 && num_pairs < 2147483647ll, "cooc_edges: bad pair count");
    if (num_pairs == 0) return PB200_OK;
    PB_REQUIRE(keys && a && b && cnt && edge_index && edge_weight && workspace, "cooc_edges: null pointer");
    PB_REQUIRE(workspace_bytes >= pb200_cooc_edges_workspace_bytes(num_pairs), "cooc_edges: workspace too small");
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)num_pairs);
    char* base = static_cast<char*>(workspace);
    void* temp = base;
    unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(base + align_up(sort_bytes, 256));
    uint32_t* idx_in = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(keys_out) + align_up((size_t)num_pairs * 8, 256));
    uint32_t* idx_out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(idx_in) + align_up((size_t)num_pairs * 4, 256));
    const unsigned blocks = (unsigned)(ceil_div(num_pairs, 256) < kSMs * 8 ? ceil_div(num_pairs, 256) : kSMs * 8);
    cooc_iota_kernel<<<blocks, 256, 0, stream>>>(idx_in, num_pairs);
    int rc = check_launch("cooc_iota_kernel");
    if (rc) return rc;
    PB_CUDA(cub::DeviceRadixSort::SortPairs(temp, sort_bytes, reinterpret_cast<const unsigned long long*>(keys), keys_out,
                                            idx_in, idx_out, (int)num_pairs, 0, 64, stream));
    count_launch(4);
    cooc_edges_kernel<<<blocks, 256, 0, stream>>>(idx_out, a, b, cnt, num_pairs, edge_index, edge_weight);
    return check_launch("cooc_edges_kernel");
}
