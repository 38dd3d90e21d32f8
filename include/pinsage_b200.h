/*
 * pinsage_b200.h -- C ABI of libpinsage_b200.so: the B200 (sm_100a) replacement for the
 * PinSage inference + retrieval hot path of anisanazim/Movie-Recommendation-Engine.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes; every data pointer is a DEVICE pointer unless
 *     the parameter name ends in _host; no torch types.
 *   - returns 0 (PB200_OK) or a negative pb200_status; never throws; the message for the
 *     last failure on the calling thread is pb200_last_error().
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises the device or allocates device memory: scratch comes from the caller
 *     (`workspace`, sized by the matching *_workspace_bytes query).
 *   - matrices are row-major and contiguous; ids are int32 on the device.
 *
 * The reference has no FFI: its "operator API" is the Python classes cited on each entry
 * point below (file:line under the reference repo).  INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 */
#ifndef PINSAGE_B200_H
#define PINSAGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ABI_VERSION 1

typedef void* pb200_stream_t; /* cudaStream_t */

typedef enum {
    PB200_OK = 0,
    PB200_ERR_INVALID_ARG = -1, /* bad size / null pointer / unsupported combination */
    PB200_ERR_CUDA = -2,        /* a CUDA runtime call or kernel launch failed */
    PB200_ERR_WORKSPACE = -3,   /* caller-provided workspace too small */
    PB200_ERR_UNSUPPORTED = -4  /* shape outside what the kernels implement */
} pb200_status;

int pb200_abi_version(void);
const char* pb200_last_error(void);
/* number of kernel launches issued by this library on the calling process so far
 * (bench.py reports the delta over the timed region as "gpu_launches") */
int64_t pb200_launch_count(void);
/* Device-wide L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes; the
 * CUDA default is 64).  The walk kernel fetches one random 32-byte block per step: at 64 every miss
 * moves two DRAM sectors.  Affects every kernel of the process on the current device; the Python
 * mirror sets 32 when a RandomWalkSampler is built unless PB200_L2_FETCH overrides it. */
int pb200_set_l2_fetch_granularity(int bytes);
int pb200_get_l2_fetch_granularity(void);

/* ------------------------------------------------------------------------------------
 * S0  RandomWalkSampler.__init__/_prepare_adjacency_list   (utils/random_walk.py:11-50)
 * Stable (edge-order preserving) CSR of a directed edge list + row-local cumulative
 * weights, replacing the Python adjacency list.
 * ------------------------------------------------------------------------------------ */

/* flags_out[0] = smallest s in [0,10] such that w*2^s is integral for every edge (11 if
 * none: use the float64 prefix path); flags_out[1] = number of negative or NaN weights
 * (saturating); flags_out must be zero-initialised by the caller. */
int pb200_edge_weight_probe(const float* edge_weights, int64_t num_edges, int32_t* flags_out,
                            pb200_stream_t stream);

size_t pb200_csr_build_workspace_bytes(int64_t num_edges, int64_t num_nodes);

/* edge_index: int64 [2, E] (row 0 = src, row 1 = dst), edge_weights: float32 [E] or NULL
 * (=1.0, random_walk.py:45-48).  quant_shift >= 0: cum is uint32 [E], row-local inclusive
 * prefix of w * 2^quant_shift (exact integers); quant_shift < 0: cum is float64 [E],
 * sequential row-local prefix.  Outputs: row_ptr int64 [N+1], col int32 [E], cum.
 * status_out int32 [4] (zero-initialised by the caller): [0] #edges with an id outside
 * [0,N), [1] #weights not representable, [2] #rows whose total overflows uint32,
 * [3] #non-empty rows with zero total weight.  Requires E < 2^32. */
int pb200_csr_build(const int64_t* edge_index, const float* edge_weights, int64_t num_edges,
                    int64_t num_nodes, int quant_shift, int64_t* row_ptr, int32_t* col,
                    void* cum, int32_t* status_out, void* workspace, size_t workspace_bytes,
                    pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * S1-S3  _single_walk / sample_neighbors / batch_sample_neighbors
 *                                                       (utils/random_walk.py:52-142)
 * One warp per start node: num_walks weighted walks of <= walk_length steps driven by
 * Philox4x32-10 (key = seed, counter = (start, walk, step/2, epoch)), visit counting in a
 * per-warp shared-memory hash table, warp-level top-T by (count desc, first visit asc),
 * weights = count / sum(kept counts).
 *   starts int32 [n]; out_ids int32 [n,T] (-1 padded); out_counts int32 [n,T];
 *   out_weights float32 [n,T] (= (float)((double)count/total)); out_nvalid int32 [n];
 *   trace_out (optional, may be NULL) int32 [n, W, L], -1 where a walk had stopped.
 * cum_kind: 0 = uint32 quanta, 1 = float64.  Limits: W*L <= 65535, W*L <= ~13000 (smem).
 * ------------------------------------------------------------------------------------ */
int pb200_walk_topt(const int64_t* row_ptr, const int32_t* col, const void* cum, int cum_kind,
                    int64_t num_nodes, const int32_t* starts, int64_t n, int num_walks,
                    int walk_length, int num_neighbors, uint64_t seed, uint32_t epoch,
                    int32_t* out_ids, int32_t* out_counts, float* out_weights,
                    int32_t* out_nvalid, int32_t* trace_out, pb200_stream_t stream);

/* Sampling index over the CSR (uint32-quanta graphs): an implicit 8-ary search tree per row
 * whose nodes are single 256-bit loads -- meta uint32 [4*N] {leaf block offset, degree, row
 * total, upper-level block offset}, leaf uint32 [16 * leaf_blocks] {8 cumulative weights | 8
 * neighbour ids}, idx uint32 [8 * idx_blocks] separator keys (top level first).  It selects
 * exactly the edge the flat search selects; a step costs ~3 dependent loads and one DRAM
 * access instead of ~log2(degree).
 *   1. pb200_walk_index_sizes: sizes_out int64 [2] = {leaf_blocks, idx_blocks} (device);
 *      offsets are left in `workspace` for step 2 (same buffer, unmodified in between);
 *   2. the caller allocates meta / idx (32 B aligned) / leaf (64 B aligned);
 *   3. pb200_walk_index_build fills them. */
size_t pb200_walk_index_workspace_bytes(int64_t num_nodes);
int pb200_walk_index_sizes(const int64_t* row_ptr, int64_t num_nodes, int64_t* sizes_out,
                           void* workspace, size_t workspace_bytes, pb200_stream_t stream);
int pb200_walk_index_build(const int64_t* row_ptr, const int32_t* col, const void* cum,
                           int64_t num_nodes, const void* workspace, uint32_t* meta, uint32_t* idx,
                           uint32_t* leaf, pb200_stream_t stream);

/* Leaf formats of the sampling index.  WIDE: 64-byte blocks {8 x u32 cumulative weight, 8 x u32 id}.
 * COMPACT: 32-byte blocks {8 x u8 (block separator - cumulative weight), 8 x u16 id low, 8 x u8 id
 * high}: one 256-bit load per walk step instead of two and half the DRAM bytes; valid when
 * num_nodes < 2^24 and pb200_walk_index_leaf_range reports <= 255 (leaf buffer: 32 bytes per
 * leaf block instead of 64).  The selected edges are identical. */
#define PB200_LEAF_WIDE 0
#define PB200_LEAF_COMPACT 1
/* BUCKET: direct-addressed index (csrc/walk_bucket.cu) -- meta uint32 [4*N] {first bucket, degree,
 * row total S, shift s}; bucket j of a row is one 32-byte block with every edge overlapping
 * [j 2^s, (j+1) 2^s) of the row's weight axis {8 x u8 min(cum - j 2^s, 2^s), 8 x id byte 0, 8 x id
 * byte 1, 8 x id byte 2}.  A walk step is meta -> bucket (t >> s): two dependent loads.  Needs every
 * weight >= 1 quantum and num_nodes <= 2^24; idx is unused (may be NULL).  Same selected edges. */
#define PB200_LEAF_BUCKET 2
/*   1. pb200_walk_bucket_plan: per-row shift + bucket counts -> meta {_, degree, S, s}; info_out
 *      uint64 [2] (device) = {total buckets, number of zero-weight edges (format unusable if != 0)};
 *      row offsets stay in `workspace` for step 3;
 *   2. the caller allocates leaf (32 bytes per bucket, 32 B aligned);
 *   3. pb200_walk_bucket_fill writes the buckets and meta[.].x. */
size_t pb200_walk_bucket_workspace_bytes(int64_t num_nodes);
int pb200_walk_bucket_plan(const int64_t* row_ptr, const void* cum, int64_t num_nodes, uint32_t* meta,
                           uint64_t* info_out, void* workspace, size_t workspace_bytes,
                           pb200_stream_t stream);
int pb200_walk_bucket_fill(const int64_t* row_ptr, const int32_t* col, const void* cum,
                           int64_t num_nodes, const void* workspace, uint32_t* meta, uint32_t* leaf,
                           uint64_t total_buckets, pb200_stream_t stream);
/* BUCKET32: the bucket format for graphs with more than 2^24 nodes -- SIX slots per 32-byte block {6 x u8 rel,
 * 2 unused bytes (128), 6 x u32 neighbour id}; the shift rule then forbids 7 edges per bucket.  The _ex entry
 * points take the format (PB200_LEAF_BUCKET or PB200_LEAF_BUCKET32). */
#define PB200_LEAF_BUCKET32 3
int pb200_walk_bucket_plan_ex(const int64_t* row_ptr, const void* cum, int64_t num_nodes, uint32_t* meta,
                              uint64_t* info_out, void* workspace, size_t workspace_bytes, int leaf_format,
                              pb200_stream_t stream);
int pb200_walk_bucket_fill_ex(const int64_t* row_ptr, const int32_t* col, const void* cum,
                              int64_t num_nodes, const void* workspace, uint32_t* meta, uint32_t* leaf,
                              uint64_t total_buckets, int leaf_format, pb200_stream_t stream);
/* max over all 8-edge leaf blocks of (last - first cumulative weight) -> *max_range_out (device u32) */
int pb200_walk_index_leaf_range(const int64_t* row_ptr, const void* cum, int64_t num_nodes,
                                uint32_t* max_range_out, pb200_stream_t stream);
int pb200_walk_index_build_ex(const int64_t* row_ptr, const int32_t* col, const void* cum,
                              int64_t num_nodes, const void* workspace, uint32_t* meta, uint32_t* idx,
                              uint32_t* leaf, int leaf_format, pb200_stream_t stream);
/* pb200_walk_topt on the sampling index (same outputs, bit-identical results). */
int pb200_walk_topt_indexed(const uint32_t* meta, const uint32_t* idx, const uint32_t* leaf,
                            int64_t num_nodes, const int32_t* starts, int64_t n, int num_walks,
                            int walk_length, int num_neighbors, uint64_t seed, uint32_t epoch,
                            int32_t* out_ids, int32_t* out_counts, float* out_weights,
                            int32_t* out_nvalid, int32_t* trace_out, pb200_stream_t stream);

/* pb200_walk_topt_indexed with the sampling epoch advanced ON THE DEVICE: the effective epoch is
 * epoch + *epoch_dev (epoch_dev: device uint32, may be NULL).  Lets a whole embedding step be
 * captured in a CUDA graph whose replays still draw fresh walks: pb200_u32_add(epoch_dev, L)
 * as the last node of the graph plays the role of the advancing global RNG stream
 * (utils/random_walk.py:79; per-layer resampling at model/pinsage.py:271-275). */
int pb200_walk_topt_indexed_ex(const uint32_t* meta, const uint32_t* idx, const uint32_t* leaf,
                               int leaf_format, int64_t num_nodes, const int32_t* starts, int64_t n,
                               int num_walks, int walk_length, int num_neighbors, uint64_t seed,
                               uint32_t epoch, const uint32_t* epoch_dev, int32_t* out_ids,
                               int32_t* out_counts, float* out_weights, int32_t* out_nvalid,
                               int32_t* trace_out, pb200_stream_t stream);
/* num_epochs independent samples per start node (epochs epoch .. epoch + num_epochs - 1, as PinSage.get_embeddings
 * draws one sample per layer, model/pinsage.py:271-275) -- ONE launch over (epoch, start) pairs on the bucket
 * index, a loop of launches otherwise.  Outputs carry a leading epoch dimension: out_ids / out_counts /
 * out_weights [num_epochs, n, T], out_nvalid [num_epochs, n], trace_out [num_epochs, n, W, L]. */
int pb200_walk_topt_indexed_multi(const uint32_t* meta, const uint32_t* idx, const uint32_t* leaf, int leaf_format,
                                  int64_t num_nodes, const int32_t* starts, int64_t n, int num_walks,
                                  int walk_length, int num_neighbors, uint64_t seed, uint32_t epoch,
                                  const uint32_t* epoch_dev, int num_epochs, int32_t* out_ids, int32_t* out_counts,
                                  float* out_weights, int32_t* out_nvalid, int32_t* trace_out, pb200_stream_t stream);
int pb200_u32_add(uint32_t* counter, uint32_t delta, pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * N4  GraphBuilder.build_item_similarity_graph          (data/graph_builder.py:59-116)
 * Item-item co-occurrence graph: pairs of movies rated together by >= threshold users, both directions,
 * weight = the count, pairs in the order of their FIRST co-occurrence (users ascending, position pairs
 * i < j of the user's rows lexicographic) -- the reference's dict order.
 *   pb200_cooc_pairs: urow_ptr / uitems = user -> movies CSR (rows of the ratings table in table order),
 *     irow_ptr / iusers / ipos = movie -> (user rank ascending, position of the movie in that user's row);
 *     acc_cnt / acc_first: int32 [num_blocks, num_items] scratch; emits up to `capacity` pairs (a < b) with their
 *     order key, count; *out_count (device uint64) = number of qualifying pairs (may exceed capacity: retry).
 *   pb200_cooc_edges: sorts the pairs by key (cub radix sort) and writes edge_index int64 [2, 2P]
 *     = (a -> b), (b -> a) per pair and edge_weight float32 [2P].
 * ------------------------------------------------------------------------------------ */
int pb200_cooc_pairs(const int64_t* urow_ptr, const int32_t* uitems, const int64_t* irow_ptr, const int32_t* iusers,
                     const int32_t* ipos, int64_t num_users, int64_t num_items, int threshold, int bits_pos,
                     int32_t* acc_cnt, int32_t* acc_first, int num_blocks, uint64_t* out_key, int32_t* out_a,
                     int32_t* out_b, int32_t* out_cnt, uint64_t capacity, uint64_t* out_count, pb200_stream_t stream);
size_t pb200_cooc_edges_workspace_bytes(int64_t num_pairs);
int pb200_cooc_edges(const uint64_t* keys, const int32_t* a, const int32_t* b, const int32_t* cnt, int64_t num_pairs,
                     int64_t* edge_index, float* edge_weight, void* workspace, size_t workspace_bytes,
                     pb200_stream_t stream);

/* N4  RandomWalkSampler.compute_ppr_matrix / precompute_top_neighbors
 *                                                                  (utils/random_walk.py:144-229)
 * pb200_ppr_push: per source, num_iterations in-place push sweeps over all nodes in index order (alpha =
 * teleport probability) on the CSR of pb200_csr_build; ppr_out / residual_ws: float64 [num_sources, vec_len]
 * (vec_len >= num_nodes: the reference sizes the vectors max(edge_index.max() + 1, max(nodes) + 1)).
 * pb200_topk_rows_f64: per row the k largest strictly positive scores, ties by smaller index; -1 / 0.0 padding. */
int pb200_ppr_push(const int64_t* row_ptr, const int32_t* col, const void* cum, int cum_kind, int quant_shift,
                   int64_t num_nodes, int64_t vec_len, const int32_t* sources, int64_t num_sources, double alpha,
                   int num_iterations, double* ppr_out, double* residual_ws, pb200_stream_t stream);
int pb200_topk_rows_f64(const double* scores, int64_t num_rows, int64_t row_len, int k, int32_t* out_ids,
                        double* out_scores, pb200_stream_t stream);

/* Counting stage alone, given traces (parity "given the same walk traces"):
 * trace int32 [n, V] (V = W*L visits in walk-major order, -1 = none). */
int pb200_count_topt(const int32_t* trace, int64_t n, int visits_per_start, int num_neighbors,
                     int32_t* out_ids, int32_t* out_counts, float* out_weights,
                     int32_t* out_nvalid, pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * P1-P4  neighbourhood pooling on padded ragged lists
 *   ids int32 [n,T] (left-aligned lists, entries >= len ignored), weights float32 [n,T] or
 *   NULL, list_len int32 [n] (#ids per row), weight_len int32 [n] or NULL (= list_len).
 * ------------------------------------------------------------------------------------ */
typedef enum {
    /* model/pinsage.py:101-150 ImportancePooling: drop ids > M-1 WITH their weights,
     * missing weight = 1, fp32 renormalise iff sum > 0 */
    PB200_POOL_PINSAGE = 0,
    /* model/layers.py:87-133 ImportancePoolingLayer (and WeightedMeanPoolingLayer with
     * weights): drop ids >= M, use the FIRST len(valid) weights, zero sum -> uniform */
    PB200_POOL_LAYERS = 1,
    /* model/aggregators.py:49-91 WeightedAggregator: no id filtering (ids must be in
     * range), weights[:len], zero sum -> mean */
    PB200_POOL_AGGREGATOR = 2,
    /* mean over ids < M (layers.py:189-191) */
    PB200_POOL_MEAN = 3,
    /* max over ids < M (layers.py:205-236 MaxPoolingLayer) */
    PB200_POOL_MAX = 4
} pb200_pool_mode;
/* OR into `mode`: store the pooled rows rounded to TF32 (they only feed a tensor-core layer) */
#define PB200_POOL_ROUND_TF32 0x100

int pb200_pool(const float* x, int64_t num_rows, int dim, const int32_t* ids,
               const float* weights, const int32_t* list_len, const int32_t* weight_len,
               int64_t n, int max_neighbors, int mode, float* out, pb200_stream_t stream);

/* Multi-GPU form of pb200_pool: x is split in contiguous row shards of shard_rows rows, shard r
 * living in rank r's memory (shard_ptrs[r]: a device pointer valid in THIS process -- the
 * rank's own buffer or a peer buffer opened with pb200_peer_open).  Neighbour rows are read
 * from their owners over NVLink peer access; this replaces the per-layer all-gather of h
 *  by ~1.3 remote rows per node.  shard_ptrs is a HOST array of `world` pointers. */
#define PB200_MAX_PEERS 16
int pb200_pool_sharded(const float* const* shard_ptrs, int world, int64_t shard_rows, int64_t num_rows,
                       int dim, const int32_t* ids, const float* weights, const int32_t* list_len,
                       const int32_t* weight_len, int64_t n, int max_neighbors, int mode, float* out,
                       pb200_stream_t stream);
/* Same with a choice of row layout.  BLOCKS: row i on rank i / shard_rows (local row i % shard_rows).  CYCLIC: row i
 * on rank i % world (local row i / world) -- rows dealt round-robin, which balances a popularity-sorted catalogue
 * across ranks (contiguous blocks give rank 0 all the heavy rows of the walk kernel). */
#define PB200_SHARD_BLOCKS 0
#define PB200_SHARD_CYCLIC 1
int pb200_pool_sharded_ex(const float* const* shard_ptrs, int world, int64_t shard_rows, int64_t num_rows,
                          int dim, const int32_t* ids, const float* weights, const int32_t* list_len,
                          const int32_t* weight_len, int64_t n, int max_neighbors, int mode, int layout,
                          float* out, pb200_stream_t stream);

/* Exchange buffers for pb200_pool_sharded: plain cudaMalloc allocations exported / opened
 * through CUDA IPC (one process per GPU on one box).  Handles are opaque 64-byte blobs that the
 * host side passes between ranks (torch.distributed all_gather_object). */
#define PB200_PEER_HANDLE_BYTES 64
int pb200_peer_alloc(size_t bytes, void** ptr_out);
int pb200_peer_free(void* ptr);
int pb200_peer_export(const void* ptr, uint8_t handle_out[PB200_PEER_HANDLE_BYTES]);
int pb200_peer_open(const uint8_t handle[PB200_PEER_HANDLE_BYTES], void** ptr_out);
int pb200_peer_close(void* ptr);

/* Barrier between the ranks of one box on peer memory -- one tiny kernel, so it can sit inside a
 * captured CUDA graph.  flag_ptrs_dev: DEVICE array of `world` pointers, entry r = rank r's flag
 * array (uint32[world], zero-initialised peer buffer, valid in this process); seq_counter /
 * error_flag: device uint32 owned by this rank (zero-initialised).  Everything this rank queued
 * before the barrier is visible to its peers' kernels queued after theirs.  The wait is bounded;
 * *error_flag becomes 1 if a peer never arrives. */
int pb200_peer_barrier(uint32_t* const* flag_ptrs_dev, uint32_t* seq_counter, int rank, int world,
                       uint32_t* error_flag, pb200_stream_t stream);
/* Same with an explicit bound on the wait: max_spins polls of ~40 ns each (pb200_peer_barrier uses 2^27,
 * about 5 s).  On time-out bit t of *error_flag is set for every peer t that never arrived; the caller
 * must check the flag before trusting anything computed after the barrier (the Python mirror raises). */
int pb200_peer_barrier_ex(uint32_t* const* flag_ptrs_dev, uint32_t* seq_counter, int rank, int world,
                          uint32_t* error_flag, uint64_t max_spins, pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * G1-G3  fused [gather -> importance sum -> concat -> dense -> epilogue]
 *   out[m,:] = epi( [A1[m,:K1] | A2row(m)] . W^T + bias )          out: float32 [n, N]
 *   A1 float32 [n,K1]; W float32 [N, K1+K2] in nn.Linear layout; bias float32 [N] or NULL.
 *   A2row(m) is  (a) absent (K2 = 0),  (b) dense A2[m,:K2]  (pool_x == NULL), or
 *                (c) pooled on the fly from pool_x [pool_rows, K2] with the list arguments
 *                    of pb200_pool (model/pinsage.py:232-240).
 *   flags: PB200_EPI_RELU, PB200_EPI_L2NORM (F.normalize eps 1e-12), PB200_EPI_LAYERNORM
 *   (gamma/beta = ln_gamma/ln_beta, eps 1e-5; aggregators.py:276-283).
 *   precision: PB200_PREC_FP32 / PB200_PREC_TF32 / PB200_PREC_AUTO (below).  The tensor-core
 *   path covers n_out <= 256, k1 % 4 == k2 % 4 == 0, 16-byte aligned operands, no LayerNorm.
 * ------------------------------------------------------------------------------------ */
#define PB200_EPI_RELU 1
#define PB200_EPI_L2NORM 2
#define PB200_EPI_LAYERNORM 4
#define PB200_EPI_ROUND_TF32 8 /* round the stored outputs to TF32 (an intermediate activation that
                                  only feeds the next tensor-core layer) */
#define PB200_IN_A2_TF32 32    /* same for a dense a2 */
#define PB200_IN_A1_TF32 16    /* a1 is already TF32-representable (written with PB200_EPI_ROUND_TF32 or
                                  pb200_round_tf32): the tensor-core path streams it with cp.async */
#define PB200_PREC_FP32 0 /* CUDA-core fp32 FMA (exact-fp32 reference kernel) */
#define PB200_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM; UNSUPPORTED if the shape is not covered */
#define PB200_PREC_AUTO 2 /* TF32 tensor cores where the shape is covered, CUDA-core fp32 otherwise */

/* out[i] = in[i] rounded to TF32 (round-to-nearest, ties away: cvt.rna.tf32.f32); in == out ok.
 * The tensor core reads only the 19 high bits of a TF32 operand, so weights handed to the
 * PB200_PREC_TF32 / AUTO path should be pre-rounded once (activations are rounded in-kernel). */
int pb200_round_tf32(const float* in, float* out, int64_t n, pb200_stream_t stream);

int pb200_gather_dense(const float* a1, int k1, const float* a2, int k2, const float* pool_x,
                       int64_t pool_rows, const int32_t* ids, const float* weights,
                       const int32_t* list_len, const int32_t* weight_len, int max_neighbors,
                       int pool_mode, const float* w, const float* bias, const float* ln_gamma,
                       const float* ln_beta, int64_t n, int n_out, int flags, int precision,
                       float* out, pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * E1/E2  exact search: GEMM fused with a streaming per-query top-k
 *   E1 (utils/evaluation.py:119-130): inner product, descending, exclude_ids[q] removed
 *   E2 (utils/nearest_neighbors.py:174-181, faiss IndexFlatL2): squared L2 via
 *      |q|^2+|x|^2-2<q,x> clipped at 0, ascending.
 *   queries float32 [nq,d], items float32 [nx,d]; exclude_ids int32 [nq] or NULL (-1 = none);
 *   id_offset is added to every returned id (item-sharded search);
 *   out_scores float32 [nq,k], out_ids int32 [nq,k] (-1 / +-inf padded when nx < k).
 *   Ties are broken by ascending id, so results do not depend on tiling or sharding.
 * ------------------------------------------------------------------------------------ */
#define PB200_METRIC_IP 0
#define PB200_METRIC_L2 1

size_t pb200_topk_workspace_bytes(int64_t nq, int64_t nx, int dim, int k);
int pb200_topk(const float* queries, int64_t nq, const float* items, int64_t nx, int dim, int k,
               int metric, const int32_t* exclude_ids, int32_t id_offset, float* out_scores,
               int32_t* out_ids, void* workspace, size_t workspace_bytes, pb200_stream_t stream);

/* Same contract and BITWISE the same results as pb200_topk, computed on the tensor cores:
 * tcgen05 kind::tf32 scoring GEMM fused with a streaming per-query shortlist (ks = 16 or 32
 * items per query and item split), exact fp32 re-rank of the shortlist with pb200_topk's own
 * arithmetic, and a certificate (|tf32 - fp32 score| <= |q| max|x~ - x| + |q~ - q| max|x~| + slack,
 * from the measured TF32 rounding residuals of the operands) that no item outside the
 * shortlist can enter the top-k; queries that fail it are re-run by the fp32
 * kernel from a device-side list (no host sync).  Covers dim % 4 == 0, dim <= 256,
 * k (+1 with exclude_ids) <= 24 (pb200_topk_tc_supported); PB200_ERR_UNSUPPORTED otherwise.
 * queries / items 16-byte aligned.  stats_out: optional DEVICE int32[1] = queries re-run in fp32. */
int pb200_topk_tc_supported(int64_t nq, int64_t nx, int dim, int k, int has_exclude);
size_t pb200_topk_tc_workspace_bytes(int64_t nq, int64_t nx, int dim, int k);
int pb200_topk_tc(const float* queries, int64_t nq, const float* items, int64_t nx, int dim, int k,
                  int metric, const int32_t* exclude_ids, int32_t id_offset, float* out_scores,
                  int32_t* out_ids, void* workspace, size_t workspace_bytes, int32_t* stats_out,
                  pb200_stream_t stream);

/* N2: rank of a ground-truth item in the descending similarity order of a
 * query item -- what calculate_hit_rate (utils/evaluation.py:5-36: rank <= k) and calculate_mrr
 * (:38-73: 1 / (rank / scale)) derive from their per-pair matmul + topk / sort.
 * out_rank[p] = 1 + #{j : s_j > s_gt or (s_j == s_gt and j < gt)}, s_j = <e[query_ids[p]], e[j]>
 * (fp32, the query itself is NOT excluded, like the reference).  ids must be in [0, n). */
size_t pb200_rank_of_target_workspace_bytes(int64_t num_pairs);
int pb200_rank_of_target(const float* embeddings, int64_t n, int dim, const int32_t* query_ids,
                         const int32_t* target_ids, int64_t num_pairs, int32_t* out_rank,
                         void* workspace, size_t workspace_bytes, pb200_stream_t stream);

/* Merge of per-shard candidate lists (multi-GPU: after the NCCL all-gather).
 * scores float32 [nq,c], ids int32 [nq,c] (id < 0 = padding); largest != 0 for IP. */
int pb200_topk_merge(const float* scores, const int32_t* ids, int64_t nq, int c, int k,
                     int largest, float* out_scores, int32_t* out_ids, pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * L1-L3  LSHIndex (utils/nearest_neighbors.py:7-68; faiss.IndexLSH(d, nbits, rotate=True))
 * ------------------------------------------------------------------------------------ */
/* codes[v] = sign bits of (proj . x_v), packed LSB first; proj float32 [nbits, d];
 * nbits % 32 == 0; codes uint8 [n, nbits/8]; proj_out (optional) float32 [n, nbits]. */
int pb200_lsh_encode(const float* x, int64_t n, int dim, const float* proj, int nbits,
                     uint8_t* codes, float* proj_out, pb200_stream_t stream);

/* Same codes as pb200_lsh_encode, bit for bit, with the projection on the tensor cores (tcgen05
 * TF32 kernel of pb200_gather_dense + a sign-packing kernel); projections smaller than the
 * measured TF32 error bound are recomputed in fp32 with pb200_lsh_encode's arithmetic.
 * dim % 4 == 0, 16-byte aligned x / proj / workspace. */
size_t pb200_lsh_encode_tc_workspace_bytes(int64_t n, int dim, int nbits);
int pb200_lsh_encode_tc(const float* x, int64_t n, int dim, const float* proj, int nbits,
                        uint8_t* codes, void* workspace, size_t workspace_bytes, pb200_stream_t stream);

/* exhaustive Hamming top-k over all stored codes (what the reference computes).
 * out_dist float32 [nq,k] (Hamming counts as floats, like faiss), out_ids int32 [nq,k]. */
int pb200_hamming_topk(const uint8_t* codes_q, int64_t nq, const uint8_t* codes_x, int64_t nx,
                       int code_bytes, int k, int32_t id_offset, float* out_dist,
                       int32_t* out_ids, pb200_stream_t stream);

/* Same contract and the same results as pb200_hamming_topk on the tensor cores: codes expanded
 * to +-1 bf16 vectors, <a,b> = nbits - 2 hamming(a,b) is exact in fp32, so the fused
 * tcgen05 (kind::f16) GEMM + per-row shortlist kernel of pb200_topk_tc gives the exact top-k.
 * code_bytes <= 64, k <= 32 (pb200_hamming_topk_tc_supported). */
int pb200_hamming_topk_tc_supported(int64_t nq, int64_t nx, int code_bytes, int k);
size_t pb200_hamming_topk_tc_workspace_bytes(int64_t nq, int64_t nx, int code_bytes, int k);
int pb200_hamming_topk_tc(const uint8_t* codes_q, int64_t nq, const uint8_t* codes_x, int64_t nx,
                          int code_bytes, int k, int32_t id_offset, float* out_dist, int32_t* out_ids,
                          void* workspace, size_t workspace_bytes, pb200_stream_t stream);

/* bucketed mode: num_tables keys of (8*code_bytes/num_tables) in {8,16} bits.
 * bucket_offsets int32 [num_tables, 2^key_bits + 1], bucket_ids int32 [num_tables, nx]. */
size_t pb200_lsh_tables_workspace_bytes(int64_t nx, int code_bytes, int num_tables);
int pb200_lsh_build_tables(const uint8_t* codes_x, int64_t nx, int code_bytes, int num_tables,
                           int32_t* bucket_offsets, int32_t* bucket_ids, void* workspace,
                           size_t workspace_bytes, pb200_stream_t stream);
/* probe the query's bucket in every table, exact dedup (first matching table wins),
 * re-rank by full-code Hamming (vectors == NULL) or inner product, warp top-k.
 * out_ncand (optional) int32 [nq]: unique candidates examined per query. */
int pb200_lsh_search_tables(const uint8_t* codes_q, int64_t nq, const uint8_t* codes_x,
                            int64_t nx, int code_bytes, int num_tables,
                            const int32_t* bucket_offsets, const int32_t* bucket_ids,
                            const float* queries, const float* vectors, int dim, int k,
                            float* out_scores, int32_t* out_ids, int32_t* out_ncand,
                            pb200_stream_t stream);
/* Same with a floor per query in output units (Hamming distance ascending, or dot score descending): passes for k > 32. */
int pb200_lsh_search_tables_ex(const uint8_t* codes_q, int64_t nq, const uint8_t* codes_x, int64_t nx, int code_bytes,
                               int num_tables, const int32_t* bucket_offsets, const int32_t* bucket_ids,
                               const float* queries, const float* vectors, int dim, int k, const float* floor_scores,
                               const int32_t* floor_ids, float* out_scores, int32_t* out_ids, int32_t* out_ncand,
                               pb200_stream_t stream);

/* ------------------------------------------------------------------------------------
 * I1/I2  WeakANDIndex (utils/nearest_neighbors.py:70-139; faiss.IndexIVFFlat, L2)
 * ------------------------------------------------------------------------------------ */
/* inverted lists from assignments: list_offsets int32 [nlist+1], list_ids int32 [n]
 * (ascending id inside a list, like faiss appends), list_vecs float32 [n,d] = x[list_ids]. */
size_t pb200_ivf_build_workspace_bytes(int64_t n, int nlist);
int pb200_ivf_build(const float* x, int64_t n, int dim, const int32_t* assign, int nlist,
                    int32_t* list_offsets, int32_t* list_ids, float* list_vecs, void* workspace,
                    size_t workspace_bytes, pb200_stream_t stream);
/* centroid update of one Lloyd iteration: centroids[c] = mean(list c) (unchanged if empty) */
int pb200_ivf_centroid_update(const float* list_vecs, const int32_t* list_offsets, int nlist,
                              int dim, float* centroids, pb200_stream_t stream);
/* probes int32 [nq, nprobe] = nearest centroids (from pb200_topk on the centroids); scans
 * those lists, k smallest squared L2 (direct difference form), -1 / +inf padded. */
int pb200_ivf_search(const float* queries, int64_t nq, int dim, const int32_t* probes,
                     int nprobe, const int32_t* list_offsets, const int32_t* list_ids,
                     const float* list_vecs, int k, float* out_dist, int32_t* out_ids,
                     pb200_stream_t stream);
/* pb200_ivf_search restricted to candidates strictly WORSE than (floor_dist[q], floor_ids[q]) under the
 * (distance asc, id asc) order (NULL: no floor): k > 32 is served in passes of 32, each floored by the last
 * result of the previous pass (faiss accepts any k). */
int pb200_ivf_search_ex(const float* queries, int64_t nq, int dim, const int32_t* probes, int nprobe,
                        const int32_t* list_offsets, const int32_t* list_ids, const float* list_vecs, int k,
                        const float* floor_dist, const int32_t* floor_ids, float* out_dist, int32_t* out_ids,
                        pb200_stream_t stream);

/* Same results as pb200_ivf_search (bit for bit) on the tensor cores.  The caller prepares, once
 * per index, a padded list-ordered layout of the vectors (np rows, np % 128 == 0, every list
 * starting at a multiple of 128):
 *   xp        float32 [np, dim]  TF32-rounded vectors (zero rows as padding)
 *   hxp       float32 [np + 128] 0.5 |x|^2 (+inf for padding rows)
 *   src_pos   int32   [np]       row of list_vecs / list_ids for every padded row, -1 = padding
 *   tile_list int32   [np / 128] list id of every 128-row tile
 *   xstats    float32 [3] (device) max |x|^2, max |rna_tf32(x) - x|_2, max |rna_tf32(x)|_2
 * The scoring GEMM covers all tiles; a query row scans a tile only if it probes the tile's
 * list; exact fp32 re-rank with pb200_ivf_search's own distance arithmetic; queries whose result
 * cannot be certified are re-run by the list-scan kernel from a device-side list.
 * nlist <= 128, dim % 4 == 0, dim <= 256, k <= 24.  stats_out: optional DEVICE int32[1]. */
int pb200_ivf_search_tc_supported(int64_t nq, int64_t np, int dim, int k, int nlist);
size_t pb200_ivf_search_tc_workspace_bytes(int64_t nq, int64_t np, int dim, int k, int nlist, int nprobe);
int pb200_ivf_search_tc(const float* queries, int64_t nq, int dim, const int32_t* probes, int nprobe,
                        int nlist, const int32_t* list_offsets, const int32_t* list_ids,
                        const float* list_vecs, const float* xp, const float* hxp,
                        const int32_t* src_pos, const int32_t* tile_list, int64_t np,
                        const uint32_t* xstats, int k, float* out_dist, int32_t* out_ids,
                        void* workspace, size_t workspace_bytes, int32_t* stats_out,
                        pb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PINSAGE_B200_H */
