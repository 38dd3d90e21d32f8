// pool.cu -- P1-P4: neighbourhood pooling as a stand-alone operator.
//
// Replaces the per-node Python loops of ImportancePooling (reference model/pinsage.py:101-150),
// ImportancePoolingLayer / WeightedMeanPoolingLayer / MaxPoolingLayer (model/layers.py:87-236)
// and WeightedAggregator / MeanAggregator (model/aggregators.py:13-91).
// One warp per output row; neighbour rows are read as coalesced 16-byte vectors
// (a 256-float row = two 512 B warp requests).  HBM/L2-bandwidth bound:
// 4*dim*(valid+1) + 8*T bytes per row.
#include <cstdlib>
#include <cstring>

#include "pool.cuh"

namespace pb200 {

__device__ __forceinline__ float round_tf32(float f) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
    return __uint_as_float(r);
}

// Row-sharded x (multi-GPU): rows [r * shard_rows, (r+1) * shard_rows) live in rank r's memory,
// mapped into this process by CUDA IPC; a neighbour row is read straight from its owner over
// NVLink (16-byte vector loads, ~1.3 valid rows per node on the bipartite graph) instead of
// all-gathering the whole h matrix every layer.  world == 0: plain single-buffer x.
// cyclic: row i lives on rank i % world at local row i / world (items dealt round-robin: on a
// popularity-sorted catalogue contiguous blocks give rank 0 all the heavy rows); else blocks.
struct PeerMap {
    const float* base[PB200_MAX_PEERS];
    int64_t shard_rows;
    int world;
    int cyclic;
};
// (r2, 2 GPUs: ld.global.nc / .cg / plain loads of the peer rows all take the same time -- 22.8 us vs 12.6 us
// with local pointers at 7,803 rows -- the remote round trip, not the cache path, is the cost.)

__device__ __forceinline__ const float* peer_row(const PeerMap& pm, int id, int dim) {
    int owner, local;
    if (pm.cyclic) { local = id / pm.world; owner = id - local * pm.world; }
    else { owner = (int)(id / pm.shard_rows); local = (int)(id - owner * pm.shard_rows); }
    return pm.base[owner] + (int64_t)local * dim;
}

template <bool kVec>
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ x, int dim,
                                                   ListArgs a, int64_t n, float* __restrict__ out,
                                                   bool round_out, const PeerMap pm) {
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    int* s_id = smem + (size_t)warp * 2 * a.T;
    float* s_w = reinterpret_cast<float*>(s_id + a.T);
    const bool is_max = a.mode == PB200_POOL_MAX;
    for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < n; row += (int64_t)gridDim.x * wpb) {
        const int nv = prepare_list(a, row, s_id, s_w, lane);
        float* o = out + row * dim;
        if (kVec) {
            // All column chunks of a neighbour row are requested together (dim <= 512: 4 x 16 B per lane in
            // flight), and two neighbour rows per iteration: a row shard on another GPU costs one NVLink round
            // trip per REQUEST WAVE, not per chunk (r2: pool_sharded 31 us vs 11 us with local pointers when
            // the chunks were fetched one after the other).
            for (int c0 = lane * 4; c0 < dim; c0 += 512) {
                float4 acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    acc[k] = is_max && nv ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < nv; r += 2) {
                    const bool two = r + 1 < nv;
                    const float* xa = pm.world > 0 ? peer_row(pm, s_id[r], dim) : x + (int64_t)s_id[r] * dim;
                    const float* xb = !two ? xa : (pm.world > 0 ? peer_row(pm, s_id[r + 1], dim) : x + (int64_t)s_id[r + 1] * dim);
                    float4 va[4], vb[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = c0 + 128 * k;
                        va[k] = c < dim ? __ldg(reinterpret_cast<const float4*>(xa + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        vb[k] = (two && c < dim) ? __ldg(reinterpret_cast<const float4*>(xb + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const float wa = is_max ? 0.f : s_w[r], wb = (is_max || !two) ? 0.f : s_w[r + 1];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (is_max) {
                            acc[k].x = fmaxf(acc[k].x, va[k].x); acc[k].y = fmaxf(acc[k].y, va[k].y);
                            acc[k].z = fmaxf(acc[k].z, va[k].z); acc[k].w = fmaxf(acc[k].w, va[k].w);
                            if (two) {
                                acc[k].x = fmaxf(acc[k].x, vb[k].x); acc[k].y = fmaxf(acc[k].y, vb[k].y);
                                acc[k].z = fmaxf(acc[k].z, vb[k].z); acc[k].w = fmaxf(acc[k].w, vb[k].w);
                            }
                        } else {                         // same order of additions as before: row r, then row r + 1
                            acc[k].x = fmaf(wa, va[k].x, acc[k].x); acc[k].y = fmaf(wa, va[k].y, acc[k].y);
                            acc[k].z = fmaf(wa, va[k].z, acc[k].z); acc[k].w = fmaf(wa, va[k].w, acc[k].w);
                            if (two) {
                                acc[k].x = fmaf(wb, vb[k].x, acc[k].x); acc[k].y = fmaf(wb, vb[k].y, acc[k].y);
                                acc[k].z = fmaf(wb, vb[k].z, acc[k].z); acc[k].w = fmaf(wb, vb[k].w, acc[k].w);
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = c0 + 128 * k;
                    if (c < dim) {
                        float4 a4 = acc[k];
                        if (round_out) {
                            a4.x = round_tf32(a4.x); a4.y = round_tf32(a4.y);
                            a4.z = round_tf32(a4.z); a4.w = round_tf32(a4.w);
                        }
                        *reinterpret_cast<float4*>(o + c) = a4;
                    }
                }
            }
        } else {
            for (int c = lane; c < dim; c += 32) {
                float acc = is_max && nv ? -INFINITY : 0.f;
                for (int r = 0; r < nv; ++r) {
                    const float* xr;
                    if (pm.world > 0) {
                        xr = peer_row(pm, s_id[r], dim);
                    } else {
                        xr = x + (int64_t)s_id[r] * dim;
                    }
                    const float v = __ldg(xr + c);
                    acc = is_max ? fmaxf(acc, v) : fmaf(s_w[r], v, acc);
                }
                o[c] = round_out ? round_tf32(acc) : acc;
            }
        }
        __syncwarp();
    }
}

}  // namespace pb200

using namespace pb200;

static int pool_launch(const float* x, const PeerMap& pm, int64_t num_rows, int dim, const int32_t* ids,
                       const float* weights, const int32_t* list_len, const int32_t* weight_len,
                       int64_t n, int max_neighbors, int mode, float* out, cudaStream_t stream) {
    PB_REQUIRE(n >= 0 && dim > 0 && max_neighbors > 0 && num_rows >= 0, "pool: bad sizes");
    const bool round_out = mode & PB200_POOL_ROUND_TF32;
    mode &= ~PB200_POOL_ROUND_TF32;
    PB_REQUIRE(mode >= PB200_POOL_PINSAGE && mode <= PB200_POOL_MAX, "pool: unknown mode %d", mode);
    if (n == 0) return PB200_OK;
    PB_REQUIRE((x || pm.world > 0) && ids && out, "pool: null pointer");
    ListArgs a{ids, weights, list_len, weight_len, max_neighbors, mode, num_rows};
    const int wpb = 8;
    const size_t smem = (size_t)wpb * 2 * max_neighbors * sizeof(int32_t);
    PB_REQUIRE(smem <= 200 * 1024, "pool: max_neighbors=%d too large", max_neighbors);
    bool vec = dim % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
    for (int r = 0; r < pm.world; ++r) vec = vec && ((uintptr_t)pm.base[r] % 16 == 0);
    auto kern = vec ? pool_kernel<true> : pool_kernel<false>;
    if (smem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div(n, wpb);
    const int64_t cap = (int64_t)kSMs * 16;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, wpb * 32, smem, stream>>>(x, dim, a, n, out, round_out, pm);
    return check_launch("pool_kernel");
}

extern "C" int pb200_pool(const float* x, int64_t num_rows, int dim, const int32_t* ids,
                          const float* weights, const int32_t* list_len,
                          const int32_t* weight_len, int64_t n, int max_neighbors, int mode,
                          float* out, pb200_stream_t stream) {
    PeerMap pm{};
    return pool_launch(x, pm, num_rows, dim, ids, weights, list_len, weight_len, n, max_neighbors, mode,
                       out, (cudaStream_t)stream);
}

extern "C" int pb200_pool_sharded_ex(const float* const* shard_ptrs, int world, int64_t shard_rows,
                                     int64_t num_rows, int dim, const int32_t* ids, const float* weights,
                                     const int32_t* list_len, const int32_t* weight_len, int64_t n,
                                     int max_neighbors, int mode, int layout, float* out, pb200_stream_t stream) {
    PB_REQUIRE(shard_ptrs && world >= 1 && world <= PB200_MAX_PEERS && shard_rows > 0,
               "pool_sharded: need 1..%d shard pointers and shard_rows > 0", PB200_MAX_PEERS);
    PB_REQUIRE(num_rows <= (int64_t)world * shard_rows, "pool_sharded: num_rows exceeds world * shard_rows");
    PB_REQUIRE(layout == PB200_SHARD_BLOCKS || layout == PB200_SHARD_CYCLIC, "pool_sharded: unknown layout %d", layout);
    PeerMap pm{};
    for (int r = 0; r < world; ++r) {
        PB_REQUIRE(shard_ptrs[r], "pool_sharded: shard pointer %d is null", r);
        pm.base[r] = shard_ptrs[r];
    }
    pm.shard_rows = shard_rows; pm.world = world; pm.cyclic = layout == PB200_SHARD_CYCLIC;
    return pool_launch(nullptr, pm, num_rows, dim, ids, weights, list_len, weight_len, n, max_neighbors,
                       mode, out, (cudaStream_t)stream);
}

extern "C" int pb200_pool_sharded(const float* const* shard_ptrs, int world, int64_t shard_rows,
                                  int64_t num_rows, int dim, const int32_t* ids, const float* weights,
                                  const int32_t* list_len, const int32_t* weight_len, int64_t n,
                                  int max_neighbors, int mode, float* out, pb200_stream_t stream) {
    return pb200_pool_sharded_ex(shard_ptrs, world, shard_rows, num_rows, dim, ids, weights, list_len, weight_len, n,
                                 max_neighbors, mode, PB200_SHARD_BLOCKS, out, stream);
}

// ---- barrier between the ranks of one box on peer memory (a plain kernel: CUDA-graph capturable) ----
// flags[r] points to rank r's flag array (uint32[world], peer memory).  Rank `rank` publishes its
// next sequence number into slot [rank] of every rank's array and waits until all slots of its
// own array have reached it.  The wait is bounded (~seconds): a dead peer must not hang the GPU;
// on time-out *error_flag is set and the kernel returns.
namespace pb200 {
__global__ void peer_barrier_kernel(uint32_t* const* __restrict__ flags, uint32_t* seq_counter, int rank,
                                    int world, uint32_t* error_flag, unsigned long long max_spins) {
    const int t = threadIdx.x;
    const uint32_t seq = *seq_counter + 1u;
    __threadfence_system();                       // this rank's earlier writes before its flag
    if (t < world) {
        volatile uint32_t* theirs = flags[t] + rank;
        *theirs = seq;
        volatile uint32_t* mine = flags[rank] + t;
        unsigned long long spins = 0;
        while ((int32_t)(*mine - seq) < 0) {
            if (++spins > max_spins) { atomicOr(error_flag, 1u << (t & 31)); break; }   // bit t: peer t never arrived
            if (spins > 256) __nanosleep(40);      // the usual wait is a few microseconds: poll hard first
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) *seq_counter = seq;
}
}  // namespace pb200

extern "C" int pb200_peer_barrier_ex(uint32_t* const* flag_ptrs_dev, uint32_t* seq_counter, int rank, int world,
                                     uint32_t* error_flag, uint64_t max_spins, pb200_stream_t stream) {
    PB_REQUIRE(flag_ptrs_dev && seq_counter && error_flag && world >= 1 && world <= PB200_MAX_PEERS &&
               rank >= 0 && rank < world && max_spins > 0, "peer_barrier: bad arguments");
    pb200::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag_ptrs_dev, seq_counter, rank, world,
                                                                     error_flag, (unsigned long long)max_spins);
    return check_launch("peer_barrier_kernel");
}

extern "C" int pb200_peer_barrier(uint32_t* const* flag_ptrs_dev, uint32_t* seq_counter, int rank, int world,
                                  uint32_t* error_flag, pb200_stream_t stream) {
    return pb200_peer_barrier_ex(flag_ptrs_dev, seq_counter, rank, world, error_flag, 1ull << 27, stream);
}

// ---- peer buffers: cudaMalloc'ed exchange buffers shared between the ranks of one box by CUDA IPC ----
extern "C" int pb200_peer_alloc(size_t bytes, void** ptr_out) {
    PB_REQUIRE(ptr_out && bytes > 0, "peer_alloc: bad arguments");
    PB_CUDA(cudaMalloc(ptr_out, bytes));
    return PB200_OK;
}
extern "C" int pb200_peer_free(void* ptr) {
    if (ptr) PB_CUDA(cudaFree(ptr));
    return PB200_OK;
}
extern "C" int pb200_peer_export(const void* ptr, uint8_t handle_out[PB200_PEER_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == PB200_PEER_HANDLE_BYTES, "IPC handle size");
    PB_REQUIRE(ptr && handle_out, "peer_export: null pointer");
    cudaIpcMemHandle_t h;
    PB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle_out, &h, sizeof(h));
    return PB200_OK;
}
extern "C" int pb200_peer_open(const uint8_t handle[PB200_PEER_HANDLE_BYTES], void** ptr_out) {
    PB_REQUIRE(handle && ptr_out, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    PB_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return PB200_OK;
}
extern "C" int pb200_peer_close(void* ptr) {
    if (ptr) PB_CUDA(cudaIpcCloseMemHandle(ptr));
    return PB200_OK;
}
