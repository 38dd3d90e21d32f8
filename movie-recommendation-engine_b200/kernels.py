"""Tensor-in / tensor-out wrappers over the C ABI (one function per entry point).

These are the only callers of ``_native.lib()``.  They allocate outputs and workspaces with
torch's caching allocator, pass raw pointers + the current stream, and never synchronise
unless a host-side flag has to be read (build-time only).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _native as N
from ._native import check, lib, ptr, stream_ptr


@dataclass
class CSR:
    """Device-resident graph: stable CSR + row-local cumulative weights (S0)."""
    row_ptr: torch.Tensor      # int64 [N+1]
    col: torch.Tensor          # int32 [E]
    cum: torch.Tensor          # uint32-as-int32 [E] (quanta) or float64 [E]
    cum_kind: int              # 0 = uint32 quanta, 1 = float64
    quant_shift: int           # weights were multiplied by 2**quant_shift (cum_kind 0)
    num_nodes: int
    num_edges: int
    # sampling index (8-ary search tree per row; uint32-quanta graphs only)
    meta: Optional[torch.Tensor] = None     # int32 [N, 4]
    idx: Optional[torch.Tensor] = None      # int32 [idx_blocks, 8]
    leaf: Optional[torch.Tensor] = None     # int32 [leaf_blocks, 16] (wide) or [leaf_blocks, 8] (compact)
    leaf_format: int = 0                    # N.LEAF_WIDE / N.LEAF_COMPACT

    @property
    def device(self):
        return self.row_ptr.device

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.row_ptr, self.col, self.cum))

    def index_nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.meta, self.idx, self.leaf)
                   if t is not None)


def _aligned_empty(n_int32, align, dev):
    """int32 buffer whose data pointer is `align`-byte aligned (torch gives >= 256 B for fresh
    allocations; asserted rather than assumed)."""
    t = torch.empty(max(n_int32, 1), dtype=torch.int32, device=dev)
    assert t.data_ptr() % align == 0
    return t


def _build_bucket_index(csr, wide_ids=None):
    """Direct-addressed bucket index (pb200_walk_bucket_*): meta -> one 32-byte bucket per walk step.
    wide_ids: None = 8 slots with 24-bit ids when the graph has at most 2^24 nodes, else 6 slots with
    32-bit ids (PB200_LEAF_BUCKET32); True forces the 32-bit form.  Returns False (csr untouched) when the
    format does not apply: a zero-weight edge, or weights so heavy that the buckets would take more
    than ~48 (64 for the 6-slot form) bytes per edge."""
    dev = csr.device
    st = stream_ptr(dev)
    Nn, E = csr.num_nodes, csr.num_edges
    if E == 0:
        return False
    if wide_ids is None:
        wide_ids = Nn > (1 << 24)
    if not wide_ids and Nn > (1 << 24):
        return False
    fmt = N.LEAF_BUCKET32 if wide_ids else N.LEAF_BUCKET
    ws_bytes = lib().pb200_walk_bucket_workspace_bytes(Nn)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    meta = _aligned_empty(4 * Nn, 16, dev)
    info = torch.zeros(2, dtype=torch.int64, device=dev)
    check(lib().pb200_walk_bucket_plan_ex(ptr(csr.row_ptr), ptr(csr.cum), Nn, ptr(meta), ptr(info), ptr(ws),
                                          ws_bytes, fmt, st), "walk_bucket_plan")
    buckets, zero_edges = info.tolist()               # build-time sync
    if zero_edges or buckets >= (1 << 32) or buckets * 32 > (64 if wide_ids else 48) * E + 64 * Nn:
        return False
    leaf = _aligned_empty(8 * max(buckets, 1), 32, dev)
    check(lib().pb200_walk_bucket_fill_ex(ptr(csr.row_ptr), ptr(csr.col), ptr(csr.cum), Nn, ptr(ws), ptr(meta),
                                          ptr(leaf), buckets, fmt, st), "walk_bucket_fill")
    csr.meta = meta.view(Nn, 4)
    csr.idx = None
    csr.leaf = leaf.view(-1, 8)[:buckets]
    csr.leaf_format = fmt
    return True


def build_walk_index(csr, leaf=None):
    """Adds a sampling index to a uint32-quanta CSR.  `leaf` (or PB200_WALK_LEAF): "auto" (bucket index
    when it applies -- 24-bit ids up to 2^24 nodes, 32-bit ids beyond -- else the 8-ary tree with compact or
    wide leaves), "bucket", "bucket32" (the 32-bit-id bucket form on any graph), "compact", "wide"."""
    if csr.cum_kind != 0 or csr.num_nodes == 0 or csr.num_edges == 0:     # nothing to index: the flat kernel serves
        return csr
    import os
    want = (leaf or os.environ.get("PB200_WALK_LEAF", "auto")).lower()
    if want in ("auto", "bucket", "bucket32"):
        if _build_bucket_index(csr, wide_ids=True if want == "bucket32" else None):
            return csr
        if want != "auto":
            raise N.NativeError("the bucket sampling index does not apply to this graph (zero-weight edges "
                                "or very heavy weights)")
    dev = csr.device
    st = stream_ptr(dev)
    Nn = csr.num_nodes
    ws_bytes = lib().pb200_walk_index_workspace_bytes(Nn)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    sizes = torch.zeros(2, dtype=torch.int64, device=dev)
    check(lib().pb200_walk_index_sizes(ptr(csr.row_ptr), Nn, ptr(sizes), ptr(ws), ws_bytes, st),
          "walk_index_sizes")
    rng = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().pb200_walk_index_leaf_range(ptr(csr.row_ptr), ptr(csr.cum) if csr.num_edges else ptr(csr.row_ptr),
                                            Nn, ptr(rng), st), "walk_index_leaf_range")
    leaf_blocks, idx_blocks = sizes.tolist()          # build-time sync
    # compact 32-byte leaves (one 256-bit load per walk step) when ids fit 24 bits and every
    # 8-edge block spans <= 255 weight quanta; PB200_WALK_LEAF=wide forces the 64-byte format
    compact = (Nn < (1 << 24) and 0 <= int(rng.item()) <= 255 and want != "wide")
    words = 8 if compact else 16
    meta = _aligned_empty(4 * Nn, 16, dev)
    idx = _aligned_empty(8 * max(idx_blocks, 1), 32, dev)
    leaf = _aligned_empty(words * max(leaf_blocks, 1), 64, dev)
    check(lib().pb200_walk_index_build_ex(ptr(csr.row_ptr), ptr(csr.col),
                                          ptr(csr.cum) if csr.num_edges else ptr(csr.row_ptr), Nn,
                                          ptr(ws), ptr(meta), ptr(idx), ptr(leaf), N.LEAF_COMPACT if compact else N.LEAF_WIDE,
                                          st),
          "walk_index_build")
    csr.meta = meta.view(Nn, 4)
    csr.idx = idx.view(-1, 8)[:idx_blocks]
    csr.leaf = leaf.view(-1, words)[:leaf_blocks]
    csr.leaf_format = N.LEAF_COMPACT if compact else N.LEAF_WIDE
    return csr


def csr_build(edge_index, edge_weights=None, num_nodes=None, device=None, force_float=False,
              index=True):
    """index: True / "auto" / "bucket" / "compact" / "wide" (see build_walk_index), False = CSR only."""
    dev = N.device_of(edge_index, edge_weights, device=device)
    ei = N.dev_tensor(edge_index, torch.int64, dev)
    if ei.dim() != 2 or ei.size(0) != 2:
        raise ValueError("edge_index must have shape [2, num_edges]")
    E = ei.size(1)
    if num_nodes is None:                       # reference: edge_index.max() + 1
        num_nodes = int(ei.max().item()) + 1 if E else 0
    w = None if edge_weights is None else N.dev_tensor(edge_weights, torch.float32, dev)
    if w is not None and w.numel() != E:
        raise ValueError("edge_weights must have one entry per edge")
    st = stream_ptr(dev)
    quant_shift = 0
    if w is not None and E:
        flags = torch.zeros(2, dtype=torch.int32, device=dev)
        check(lib().pb200_edge_weight_probe(ptr(w), E, ptr(flags), st), "edge_weight_probe")
        shift, bad = flags.tolist()              # build-time sync
        if bad:
            raise ValueError(f"{bad} edge weights are negative, NaN or infinite")
        quant_shift = -1 if (shift > 10 or force_float) else shift
    row_ptr = torch.empty(num_nodes + 1, dtype=torch.int64, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    cum = torch.empty(E, dtype=torch.int32 if quant_shift >= 0 else torch.float64, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    for attempt in range(2):
        ws_bytes = lib().pb200_csr_build_workspace_bytes(E, num_nodes)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        check(lib().pb200_csr_build(ptr(ei), ptr(w), E, num_nodes, quant_shift, ptr(row_ptr),
                                    ptr(col), ptr(cum) if E else None, ptr(status), ptr(ws),
                                    ws_bytes, st), "csr_build")
        bad_id, bad_w, overflow, zero_rows = status.tolist()
        del ws
        if bad_id:
            raise IndexError(f"{bad_id} edges reference node ids outside [0, {num_nodes})")
        if overflow and quant_shift >= 0 and attempt == 0:
            # a row's total does not fit uint32 quanta: fall back to the float64 prefix path
            quant_shift = -1
            cum = torch.empty(E, dtype=torch.float64, device=dev)
            status.zero_()
            continue
        if bad_w:
            raise ValueError(f"{bad_w} edge weights not representable")
        if zero_rows:
            raise ValueError(f"{zero_rows} nodes have out-edges whose weights sum to zero "
                             "(the reference's np.random.choice would raise on NaN probabilities)")
        break
    csr = CSR(row_ptr, col, cum, 0 if quant_shift >= 0 else 1, quant_shift, num_nodes, E)
    return build_walk_index(csr, leaf=index if isinstance(index, str) else None) if index else csr


def u32_add(counter, delta):
    """counter (device int32/uint32 scalar tensor) += delta, in stream order."""
    check(lib().pb200_u32_add(ptr(counter), int(delta) & 0xFFFFFFFF, stream_ptr(counter.device)), "u32_add")


def walk_topt(csr: CSR, starts, num_walks, walk_length, num_neighbors, seed, epoch=0,
              return_trace=False, use_index=True, epoch_dev=None, num_epochs=None):
    """Walks + visit counts + top-T.  num_epochs=E: E independent samples per start node (epochs epoch ..
    epoch + E - 1) in one call -- one launch on the bucket index; every output gains a leading [E] dim."""
    dev = csr.device
    s = N.dev_tensor(starts, torch.int32, dev)
    n = s.numel()
    T = int(num_neighbors)
    E = 1 if num_epochs is None else int(num_epochs)
    lead = () if num_epochs is None else (E,)
    ids = torch.empty(lead + (n, T), dtype=torch.int32, device=dev)
    counts = torch.empty(lead + (n, T), dtype=torch.int32, device=dev)
    weights = torch.empty(lead + (n, T), dtype=torch.float32, device=dev)
    nvalid = torch.empty(lead + (n,), dtype=torch.int32, device=dev)
    trace = torch.empty(lead + (n, num_walks, walk_length), dtype=torch.int32, device=dev) \
        if return_trace else None
    indexed = use_index and csr.meta is not None
    if (epoch_dev is not None or num_epochs is not None) and not indexed:
        raise N.NativeError("walk_topt: a device-side epoch / num_epochs needs the sampling index (use_index=True)")
    if indexed:
        check(lib().pb200_walk_topt_indexed_multi(ptr(csr.meta), ptr(csr.idx), ptr(csr.leaf),
                                                  int(getattr(csr, "leaf_format", N.LEAF_WIDE)),
                                                  csr.num_nodes, ptr(s), n, int(num_walks),
                                                  int(walk_length), T, int(seed) & (2**64 - 1),
                                                  int(epoch) & 0xFFFFFFFF, ptr(epoch_dev), E, ptr(ids),
                                                  ptr(counts), ptr(weights), ptr(nvalid), ptr(trace),
                                                  stream_ptr(dev)),
              "walk_topt_indexed")
        return (ids, counts, weights, nvalid, trace) if return_trace else \
            (ids, counts, weights, nvalid)
    # a graph without edges: col / cum are empty tensors (null pointers); every row is a dead end, nothing is read
    check(lib().pb200_walk_topt(ptr(csr.row_ptr), ptr(csr.col) if csr.num_edges else ptr(csr.row_ptr),
                                ptr(csr.cum) if csr.num_edges else ptr(csr.row_ptr), csr.cum_kind, csr.num_nodes, ptr(s), n,
                                int(num_walks), int(walk_length), T, int(seed) & (2**64 - 1),
                                int(epoch) & 0xFFFFFFFF, ptr(ids), ptr(counts), ptr(weights),
                                ptr(nvalid), ptr(trace), stream_ptr(dev)), "walk_topt")
    return (ids, counts, weights, nvalid, trace) if return_trace else (ids, counts, weights, nvalid)


def ppr_push(csr: CSR, sources, alpha=0.15, num_iterations=10, vec_len=None):
    """Dense PPR push scores (pb200_ppr_push): float64 [len(sources), vec_len] on the device."""
    dev = csr.device
    s = N.dev_tensor(sources, torch.int32, dev).reshape(-1)
    vec_len = csr.num_nodes if vec_len is None else max(int(vec_len), csr.num_nodes)
    ppr = torch.empty((s.numel(), vec_len), dtype=torch.float64, device=dev)
    res = torch.empty((s.numel(), vec_len), dtype=torch.float64, device=dev)
    check(lib().pb200_ppr_push(ptr(csr.row_ptr), ptr(csr.col) if csr.num_edges else ptr(csr.row_ptr),
                               ptr(csr.cum) if csr.num_edges else ptr(csr.row_ptr), csr.cum_kind,
                               max(csr.quant_shift, 0), csr.num_nodes, vec_len, ptr(s), s.numel(), float(alpha),
                               int(num_iterations), ptr(ppr), ptr(res), stream_ptr(dev)), "ppr_push")
    return ppr


def topk_rows_f64(scores, k):
    """Per row: the k largest strictly positive float64 scores, ties by smaller index (ids -1 / score 0 padded)."""
    dev = N.device_of(scores)
    sc = N.dev_tensor(scores, torch.float64, dev)
    S, n = sc.shape
    ids = torch.empty((S, k), dtype=torch.int32, device=dev)
    vals = torch.empty((S, k), dtype=torch.float64, device=dev)
    check(lib().pb200_topk_rows_f64(ptr(sc), S, n, int(k), ptr(ids), ptr(vals), stream_ptr(dev)), "topk_rows_f64")
    return ids, vals


def item_cooccurrence_graph(user_rank, movie_idx, num_users, num_items, threshold, device=None):
    """Item-item co-occurrence graph (pb200_cooc_*; reference data/graph_builder.py:59-116) from the ratings
    table given as (user rank, movie index) per row, in table order.  Returns (src int64 [2P], dst int64 [2P],
    weight float32 [2P]) on the device, pairs in the reference's dict order."""
    dev = N.device_of(user_rank, movie_idx, device=device)
    u = N.dev_tensor(user_rank, torch.int64, dev).reshape(-1)
    m = N.dev_tensor(movie_idx, torch.int64, dev).reshape(-1)
    R = u.numel()
    if R and int(torch.unique(u * num_items + m).numel()) != R:
        raise ValueError("build_item_similarity_graph: a user rates the same movie more than once "
                         "(unsupported: MovieLens ratings are unique per (user, movie))")
    st = stream_ptr(dev)
    # user -> movies (table order inside a user), then movie -> users (ascending rank) from the edges in that order
    ucsr = csr_build(torch.stack([u, m]), None, num_nodes=max(num_users, num_items), device=dev, index=False)
    urow = ucsr.row_ptr[:num_users + 1].contiguous()
    R = ucsr.num_edges
    rows = torch.repeat_interleave(torch.arange(num_users, device=dev), (urow[1:] - urow[:-1]))
    pos = (torch.arange(R, device=dev) - urow[:-1][rows]).to(torch.int32)
    # movie -> its (user, position) edges, users ascending: a stable sort of the user-CSR edges by movie.  The
    # CSR is built over EDGE NUMBERS, which gives the permutation; user rank and position follow by lookup.
    ecsr = csr_build(torch.stack([ucsr.col.to(torch.int64), torch.arange(R, device=dev)]), None,
                     num_nodes=max(num_items, R, 1), device=dev, index=False)
    irow = ecsr.row_ptr[:num_items + 1].contiguous()
    perm = ecsr.col.to(torch.int64)
    iusers = rows[perm].to(torch.int32).contiguous()
    ipos = pos[perm].contiguous()
    max_deg = int((urow[1:] - urow[:-1]).max().item()) if num_users else 1
    bits_p = max(1, int(max_deg - 1).bit_length())
    nblocks = 148 * 2
    acc_cnt = torch.empty((nblocks, num_items), dtype=torch.int32, device=dev)
    acc_first = torch.empty((nblocks, num_items), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    cap = max(1024, min(4 * R, 1 << 26))
    while True:
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        pa = torch.empty(cap, dtype=torch.int32, device=dev)
        pbb = torch.empty(cap, dtype=torch.int32, device=dev)
        pc = torch.empty(cap, dtype=torch.int32, device=dev)
        check(lib().pb200_cooc_pairs(ptr(urow), ptr(ucsr.col), ptr(irow), ptr(iusers), ptr(ipos), num_users, num_items,
                                     int(threshold), bits_p, ptr(acc_cnt), ptr(acc_first), nblocks, ptr(keys), ptr(pa),
                                     ptr(pbb), ptr(pc), cap, ptr(count), st), "cooc_pairs")
        P = int(count.item())                      # build-time sync
        if P <= cap:
            break
        cap = P
    ei = torch.empty((2, 2 * P), dtype=torch.int64, device=dev)
    ew = torch.empty(2 * P, dtype=torch.float32, device=dev)
    if P:
        ws_bytes = lib().pb200_cooc_edges_workspace_bytes(P)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().pb200_cooc_edges(ptr(keys), ptr(pa), ptr(pbb), ptr(pc), P, ptr(ei), ptr(ew), ptr(ws), ws_bytes, st),
              "cooc_edges")
    return ei[0], ei[1], ew


def count_topt(trace, num_neighbors):
    dev = N.device_of(trace)
    tr = N.dev_tensor(trace, torch.int32, dev)
    n = tr.size(0)
    V = tr.numel() // max(n, 1)
    T = int(num_neighbors)
    ids = torch.empty((n, T), dtype=torch.int32, device=dev)
    counts = torch.empty((n, T), dtype=torch.int32, device=dev)
    weights = torch.empty((n, T), dtype=torch.float32, device=dev)
    nvalid = torch.empty(n, dtype=torch.int32, device=dev)
    check(lib().pb200_count_topt(ptr(tr), n, V, T, ptr(ids), ptr(counts), ptr(weights),
                                 ptr(nvalid), stream_ptr(dev)), "count_topt")
    return ids, counts, weights, nvalid


def pool(x, ids, weights, list_len, weight_len, mode):
    dev = N.device_of(x)
    x = N.dev_tensor(x, torch.float32, dev)
    n, T = ids.shape
    out = torch.empty((n, x.size(1)), dtype=torch.float32, device=dev)
    check(lib().pb200_pool(ptr(x), x.size(0), x.size(1), ptr(ids), ptr(weights), ptr(list_len),
                           ptr(weight_len), n, T, mode, ptr(out), stream_ptr(dev)), "pool")
    return out


def pool_sharded(shard_ptr_array, world, shard_rows, num_rows, dim, ids, weights, list_len, weight_len, mode,
                 dev, layout=N.SHARD_BLOCKS):
    """pb200_pool over a row-sharded x whose shards live on the ranks of one box (peer memory)."""
    n, T = ids.shape
    out = torch.empty((n, dim), dtype=torch.float32, device=dev)
    check(lib().pb200_pool_sharded_ex(shard_ptr_array, world, shard_rows, num_rows, dim, ptr(ids), ptr(weights),
                                      ptr(list_len), ptr(weight_len), n, T, mode, int(layout), ptr(out),
                                      stream_ptr(dev)), "pool_sharded")
    return out


_TF32_CACHE = {}


def tf32_weight(w):
    """`w` rounded to TF32 (pb200_round_tf32), cached per tensor object and version: the
    tensor-core path reads only the high 19 bits of its operands, so weights are rounded once,
    not per call.  Entries die with their source tensor (weak reference), so a new tensor that
    happens to reuse freed memory can never hit a stale entry."""
    import weakref
    key = id(w)
    hit = _TF32_CACHE.get(key)
    if hit is not None and hit[0]() is w and hit[1] == (w._version, w.data_ptr()):
        return hit[2]
    src = w.detach().contiguous()
    out = torch.empty_like(src)
    check(lib().pb200_round_tf32(ptr(src), ptr(out), src.numel(), stream_ptr(src.device)),
          "round_tf32")
    try:
        ref = weakref.ref(w, lambda _r, k=key: _TF32_CACHE.pop(k, None))
    except TypeError:
        return out
    _TF32_CACHE[key] = (ref, (w._version, w.data_ptr()), out)
    return out


def gather_dense(a1, w, bias=None, a2=None, pool_x=None, lists=None, pool_mode=N.POOL_PINSAGE,
                 flags=0, precision=N.PREC_FP32, ln_gamma=None, ln_beta=None, n=None, out=None):
    """out = epi([a1 | a2-or-pooled] @ w.T + bias).  lists = (ids, weights, list_len, weight_len).
    out: optional contiguous float32 [n, n_out] device buffer to write into (e.g. a peer-memory shard)."""
    dev = N.device_of(w)
    if precision != N.PREC_FP32:
        w = tf32_weight(w)
    k1 = 0 if a1 is None else a1.size(1)
    if n is None:
        n = a1.size(0) if a1 is not None else (a2.size(0) if a2 is not None else lists[0].size(0))
    k2 = 0
    if a2 is not None:
        k2 = a2.size(1)
    elif pool_x is not None:
        k2 = pool_x.size(1)
    n_out, K = w.shape
    if K != k1 + k2:
        raise ValueError(f"weight has {K} input columns, inputs provide {k1}+{k2}")
    ids = wts = ll = wl = None
    T = 0
    if pool_x is not None:
        ids, wts, ll, wl = lists
        T = ids.size(1)
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, n_out) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 [{n}, {n_out}] tensor")
    check(lib().pb200_gather_dense(ptr(a1), k1, ptr(a2), k2, ptr(pool_x),
                                   0 if pool_x is None else pool_x.size(0), ptr(ids), ptr(wts),
                                   ptr(ll), ptr(wl), T, pool_mode, ptr(w), ptr(bias),
                                   ptr(ln_gamma), ptr(ln_beta), n, n_out, flags, precision,
                                   ptr(out), stream_ptr(dev)), "gather_dense")
    return out


# under precision="auto" small problems stay on the CUDA-core kernels: a 256-row query group per
# CTA would be mostly padding, and the operand preparation / re-rank passes are not amortised
# over a handful of items (e.g. the IVF coarse quantiser: 100 centroids)
TOPK_TC_MIN_QUERIES = 256
TOPK_TC_MIN_ITEMS = 2048


def topk(queries, items, k, metric, exclude_ids=None, id_offset=0, precision="auto", stats=None):
    """Exact top-k.  precision: "fp32" = CUDA-core kernel (pb200_topk); "tf32" = tensor-core
    shortlist + exact fp32 re-rank + certificate (pb200_topk_tc, bitwise the same result);
    "auto" = "tf32" where the shape is covered and the batch is large enough.  `stats`, if a
    dict, receives {"path": ..., "fp32_reruns": device int32 tensor}."""
    dev = N.device_of(items, queries)
    q = N.dev_tensor(queries, torch.float32, dev)
    x = N.dev_tensor(items, torch.float32, dev)
    if q.dim() == 1:
        q = q[None]
    nq, d = q.shape
    nx = x.size(0)
    if x.size(1) != d:
        raise RuntimeError(f"dimension mismatch: queries {d}, items {x.size(1)}")
    if precision not in ("auto", "fp32", "tf32"):
        raise ValueError(f"precision must be 'auto', 'fp32' or 'tf32', got {precision!r}")
    ex = None if exclude_ids is None else N.dev_tensor(exclude_ids, torch.int32, dev)
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    use_tc = False
    if precision != "fp32" and nq > 0:
        ok = bool(lib().pb200_topk_tc_supported(nq, nx, d, k, 0 if ex is None else 1))
        ok = ok and q.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0
        if precision == "tf32" and not ok:
            raise N.NativeError("topk: precision='tf32' needs dim % 4 == 0, dim <= 256, "
                                f"k (+1 with exclude_ids) <= 24 (got dim={d}, k={k})")
        use_tc = ok and (precision == "tf32" or (nq >= TOPK_TC_MIN_QUERIES and nx >= TOPK_TC_MIN_ITEMS))
    if use_tc:
        ws_bytes = lib().pb200_topk_tc_workspace_bytes(nq, nx, d, k)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        reruns = torch.zeros(1, dtype=torch.int32, device=dev) if stats is not None else None
        check(lib().pb200_topk_tc(ptr(q), nq, ptr(x), nx, d, k, metric, ptr(ex), int(id_offset),
                                  ptr(scores), ptr(ids), ptr(ws), ws_bytes, ptr(reruns),
                                  stream_ptr(dev)), "topk_tc")
        if stats is not None:
            stats.update(path="tf32", fp32_reruns=reruns)
        return scores, ids
    ws_bytes = lib().pb200_topk_workspace_bytes(nq, nx, d, k)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    check(lib().pb200_topk(ptr(q), nq, ptr(x), nx, d, k, metric, ptr(ex), int(id_offset),
                           ptr(scores), ptr(ids), ptr(ws), ws_bytes, stream_ptr(dev)), "topk")
    if stats is not None:
        stats.update(path="fp32", fp32_reruns=None)
    return scores, ids


def rank_of_target(embeddings, query_ids, target_ids):
    """1-based rank of target_ids[p] in the descending similarity order of query_ids[p] (int32 [P])."""
    dev = N.device_of(embeddings)
    e = N.dev_tensor(embeddings, torch.float32, dev)
    q = N.dev_tensor(query_ids, torch.int32, dev).reshape(-1)
    g = N.dev_tensor(target_ids, torch.int32, dev).reshape(-1)
    if q.numel() != g.numel():
        raise RuntimeError("query and ground-truth index lists differ in length")
    rank = torch.empty(q.numel(), dtype=torch.int32, device=dev)
    ws_bytes = lib().pb200_rank_of_target_workspace_bytes(q.numel())
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    check(lib().pb200_rank_of_target(ptr(e), e.size(0), e.size(1), ptr(q), ptr(g), q.numel(), ptr(rank), ptr(ws),
                                     ws_bytes, stream_ptr(dev)), "rank_of_target")
    return rank


def topk_merge(scores, ids, k, largest):
    dev = N.device_of(scores)
    s = N.dev_tensor(scores, torch.float32, dev)
    i = N.dev_tensor(ids, torch.int32, dev)
    nq, c = s.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    check(lib().pb200_topk_merge(ptr(s), ptr(i), nq, c, k, 1 if largest else 0, ptr(out_s),
                                 ptr(out_i), stream_ptr(dev)), "topk_merge")
    return out_s, out_i


def lsh_encode(x, proj, return_projection=False, precision="auto"):
    """Sign-random-projection codes.  precision: "fp32" = CUDA-core projection (pb200_lsh_encode);
    "tc" = tensor-core projection + exact fp32 recompute of the projections near zero
    (pb200_lsh_encode_tc, bit-identical codes); "auto" = "tc" for >= 256 vectors when covered."""
    dev = N.device_of(x, proj)
    x = N.dev_tensor(x, torch.float32, dev)
    proj = N.dev_tensor(proj, torch.float32, dev)
    n, d = x.shape
    nbits = proj.size(0)
    if precision not in ("auto", "fp32", "tc"):
        raise ValueError(f"precision must be 'auto', 'fp32' or 'tc', got {precision!r}")
    codes = torch.empty((n, nbits // 8), dtype=torch.uint8, device=dev)
    ok = d % 4 == 0 and x.data_ptr() % 16 == 0 and proj.data_ptr() % 16 == 0 and not return_projection
    if precision == "tc" and not ok:
        raise N.NativeError("lsh_encode: precision='tc' needs dim % 4 == 0 and no projection output")
    if ok and (precision == "tc" or (precision == "auto" and n >= TOPK_TC_MIN_QUERIES)):
        ws_bytes = lib().pb200_lsh_encode_tc_workspace_bytes(n, d, nbits)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        check(lib().pb200_lsh_encode_tc(ptr(x), n, d, ptr(proj), nbits, ptr(codes), ptr(ws), ws_bytes,
                                        stream_ptr(dev)), "lsh_encode_tc")
        return codes
    y = torch.empty((n, nbits), dtype=torch.float32, device=dev) if return_projection else None
    check(lib().pb200_lsh_encode(ptr(x), n, d, ptr(proj), nbits, ptr(codes), ptr(y),
                                 stream_ptr(dev)), "lsh_encode")
    return (codes, y) if return_projection else codes


def hamming_topk(codes_q, codes_x, k, id_offset=0, precision="auto"):
    """Exhaustive Hamming top-k.  precision: "simt" = xor + popcount kernel (pb200_hamming_topk);
    "tc" = +-1 bf16 GEMM on the tensor cores fused with the shortlist (pb200_hamming_topk_tc, exact:
    same distances and ids); "auto" = "tc" where covered and the query batch is large enough."""
    dev = N.device_of(codes_x)
    nq, cb = codes_q.shape
    nx = codes_x.size(0)
    if precision not in ("auto", "simt", "tc"):
        raise ValueError(f"precision must be 'auto', 'simt' or 'tc', got {precision!r}")
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    use_tc = False
    if precision != "simt" and nq > 0:
        ok = bool(lib().pb200_hamming_topk_tc_supported(nq, nx, cb, k))
        if precision == "tc" and not ok:
            raise N.NativeError(f"hamming_topk: precision='tc' needs code_bytes <= 64, k <= 32 (got {cb}, {k})")
        use_tc = ok and (precision == "tc" or (nq >= TOPK_TC_MIN_QUERIES and nx >= TOPK_TC_MIN_ITEMS))
    if use_tc:
        ws_bytes = lib().pb200_hamming_topk_tc_workspace_bytes(nq, nx, cb, k)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        check(lib().pb200_hamming_topk_tc(ptr(codes_q), nq, ptr(codes_x), nx, cb, k, int(id_offset),
                                          ptr(dist), ptr(ids), ptr(ws), ws_bytes, stream_ptr(dev)),
              "hamming_topk_tc")
        return dist, ids
    check(lib().pb200_hamming_topk(ptr(codes_q), nq, ptr(codes_x), nx, cb, k,
                                   int(id_offset), ptr(dist), ptr(ids), stream_ptr(dev)),
          "hamming_topk")
    return dist, ids


def lsh_build_tables(codes_x, num_tables):
    dev = N.device_of(codes_x)
    nx, cb = codes_x.shape
    key_bits = 8 * cb // num_tables
    offsets = torch.empty((num_tables, (1 << key_bits) + 1), dtype=torch.int32, device=dev)
    bucket_ids = torch.empty((num_tables, max(nx, 1)), dtype=torch.int32, device=dev)
    ws_bytes = lib().pb200_lsh_tables_workspace_bytes(nx, cb, num_tables)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    check(lib().pb200_lsh_build_tables(ptr(codes_x), nx, cb, num_tables, ptr(offsets),
                                       ptr(bucket_ids), ptr(ws), ws_bytes, stream_ptr(dev)),
          "lsh_build_tables")
    return offsets, bucket_ids


def _passes_of_32(k, run_pass, nq, dev, worst):
    """k results per query in passes of <= 32, each floored by the last result of the previous pass."""
    scores = torch.full((nq, k), worst, dtype=torch.float32, device=dev)
    ids = torch.full((nq, k), -1, dtype=torch.int32, device=dev)
    floor_s = floor_i = None
    extra = None
    for lo in range(0, k, 32):
        kk = min(32, k - lo)
        s, i, extra_p = run_pass(kk, floor_s, floor_i)
        if extra is None:
            extra = extra_p
        if floor_i is not None:                       # a query that was already short stays padded
            dead = (floor_i < 0)[:, None]
            s = torch.where(dead, torch.full_like(s, worst), s)
            i = torch.where(dead, torch.full_like(i, -1), i)
        scores[:, lo:lo + kk], ids[:, lo:lo + kk] = s, i
        floor_s, floor_i = s[:, kk - 1].contiguous(), i[:, kk - 1].contiguous()
    return scores, ids, extra


def lsh_search_tables(codes_q, codes_x, num_tables, offsets, bucket_ids, k, queries=None,
                      vectors=None):
    """Bucketed LSH search; any k (passes of 32 with a floor beyond that, like pb200_topk)."""
    dev = N.device_of(codes_x)
    nq, cb = codes_q.shape
    d = 0 if vectors is None else vectors.size(1)
    dot = vectors is not None

    def run(kk, floor_s, floor_i):
        scores = torch.empty((nq, kk), dtype=torch.float32, device=dev)
        ids = torch.empty((nq, kk), dtype=torch.int32, device=dev)
        ncand = torch.empty(nq, dtype=torch.int32, device=dev)
        check(lib().pb200_lsh_search_tables_ex(ptr(codes_q), nq, ptr(codes_x), codes_x.size(0), cb,
                                               num_tables, ptr(offsets), ptr(bucket_ids), ptr(queries),
                                               ptr(vectors), d, kk, ptr(floor_s), ptr(floor_i), ptr(scores), ptr(ids),
                                               ptr(ncand), stream_ptr(dev)), "lsh_search_tables")
        return scores, ids, ncand
    if k <= 32:
        return run(k, None, None)
    return _passes_of_32(k, run, nq, dev, float("-inf") if dot else float("inf"))


def ivf_build(x, assign, nlist):
    dev = N.device_of(x)
    n, d = x.shape
    offsets = torch.empty(nlist + 1, dtype=torch.int32, device=dev)
    list_ids = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    list_vecs = torch.empty((max(n, 1), d), dtype=torch.float32, device=dev)
    ws_bytes = lib().pb200_ivf_build_workspace_bytes(n, nlist)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    check(lib().pb200_ivf_build(ptr(x), n, d, ptr(assign), nlist, ptr(offsets), ptr(list_ids),
                                ptr(list_vecs), ptr(ws), ws_bytes, stream_ptr(dev)), "ivf_build")
    return offsets, list_ids[:n], list_vecs[:n]


def ivf_centroid_update(list_vecs, offsets, centroids):
    dev = N.device_of(centroids)
    check(lib().pb200_ivf_centroid_update(ptr(list_vecs), ptr(offsets), centroids.size(0),
                                          centroids.size(1), ptr(centroids), stream_ptr(dev)),
          "ivf_centroid_update")
    return centroids


def ivf_tc_layout(offsets, list_ids, list_vecs, nlist):
    """Padded list-ordered layout for pb200_ivf_search_tc (built once per index): every list starts
    at a multiple of 128 rows.  Returns (xp, hxp, src_pos, tile_list, xstats) or None when the
    shape is not covered (nlist > 128, dim % 4 != 0, dim > 256, empty index)."""
    dev = list_vecs.device
    n, d = list_vecs.shape
    if n == 0 or nlist > 128 or d % 4 or d > 256:
        return None
    offs = offsets.to(torch.int64)
    lens = offs[1:] - offs[:-1]
    padded = (lens + 127) // 128 * 128
    pstart = torch.cumsum(padded, 0) - padded
    np_rows = int(padded.sum().item())
    if np_rows == 0:
        return None
    # list l occupies padded rows [pstart[l], pstart[l] + lens[l]): row p of list order goes to
    # pstart[list(p)] + (p - offs[list(p)])
    lid = torch.repeat_interleave(torch.arange(nlist, device=dev), lens)
    pos = torch.arange(n, device=dev, dtype=torch.int64)
    dst = pstart[lid] + (pos - offs[:-1][lid])
    src_pos = torch.full((np_rows,), -1, dtype=torch.int32, device=dev)
    src_pos[dst] = pos.to(torch.int32)
    tile_list = torch.repeat_interleave(torch.arange(nlist, device=dev, dtype=torch.int32), padded // 128).contiguous()
    vec = torch.zeros((np_rows, d), dtype=torch.float32, device=dev)
    vec[dst] = list_vecs
    xp = torch.empty_like(vec)
    check(lib().pb200_round_tf32(ptr(vec), ptr(xp), vec.numel(), stream_ptr(dev)), "round_tf32")
    xn = (vec.double() * vec.double()).sum(1)
    hxp = torch.full((np_rows + 128,), float("inf"), dtype=torch.float32, device=dev)
    hxp[dst] = (0.5 * xn[dst]).float()
    resid = (xp.double() - vec.double()).norm(dim=1).max() * 1.0001
    xstats = torch.stack([xn.max() * 1.0001, resid, xp.double().norm(dim=1).max() * 1.0001]).float().contiguous()
    return xp, hxp.contiguous(), src_pos.contiguous(), tile_list, xstats


def ivf_search_tc(queries, probes, offsets, list_ids, list_vecs, layout, nlist, k, stats=None):
    """pb200_ivf_search_tc: tensor-core scoring of the padded list-ordered vectors, probe-masked
    shortlists, exact re-rank; equal to ivf_search bit for bit."""
    dev = N.device_of(list_vecs)
    xp, hxp, src_pos, tile_list, xstats = layout
    nq, d = queries.shape
    np_rows = xp.size(0)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    ws_bytes = lib().pb200_ivf_search_tc_workspace_bytes(nq, np_rows, d, k, nlist, probes.size(1))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    reruns = torch.zeros(1, dtype=torch.int32, device=dev) if stats is not None else None
    check(lib().pb200_ivf_search_tc(ptr(queries), nq, d, ptr(probes), probes.size(1), nlist, ptr(offsets),
                                    ptr(list_ids), ptr(list_vecs), ptr(xp), ptr(hxp), ptr(src_pos),
                                    ptr(tile_list), np_rows, ptr(xstats), k, ptr(dist), ptr(ids), ptr(ws),
                                    ws_bytes, ptr(reruns), stream_ptr(dev)), "ivf_search_tc")
    if stats is not None:
        stats.update(path="tf32", list_scan_reruns=reruns)
    return dist, ids


def ivf_search_tc_supported(nq, np_rows, d, k, nlist):
    return bool(lib().pb200_ivf_search_tc_supported(nq, np_rows, d, k, nlist))


def ivf_search(queries, probes, offsets, list_ids, list_vecs, k):
    """IVF list scan; any k (passes of 32 with a floor beyond that)."""
    dev = N.device_of(list_vecs)
    nq, d = queries.shape

    def run(kk, floor_s, floor_i):
        dist = torch.empty((nq, kk), dtype=torch.float32, device=dev)
        ids = torch.empty((nq, kk), dtype=torch.int32, device=dev)
        check(lib().pb200_ivf_search_ex(ptr(queries), nq, d, ptr(probes), probes.size(1), ptr(offsets),
                                        ptr(list_ids), ptr(list_vecs), kk, ptr(floor_s), ptr(floor_i), ptr(dist),
                                        ptr(ids), stream_ptr(dev)), "ivf_search")
        return dist, ids, None
    if k <= 32:
        return run(k, None, None)[:2]
    return _passes_of_32(k, run, nq, dev, float("inf"))[:2]
