"""Multi-GPU layout of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards and what is exchanged:
  * walks       start nodes are split in contiguous blocks; the CSR is replicated; no
                collective (Philox counters make a node's sample independent of its shard).
  * conv layers rows are split the same way; layer l+1 gathers rows of h^(l) from arbitrary
                nodes, so each layer ends with ONE all-gather of the row shards of h.
  * search      queries are split, the index is replicated, results are all-gathered; or items
                are split (exact search on catalogues that do not fit one GPU), every rank
                returns a local top-k with global ids, and the all-gathered lists are merged
                by pb200_topk_merge under the (score, id) total order -- identical to the
                unsharded result bit for bit.
Works with the NCCL backend on GPUs and with gloo on CPU tensors (host-logic tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_size(n, world_size):
    return (n + world_size - 1) // world_size


def shard_range(n, rank, world_size):
    """Contiguous block [lo, hi) of rank; all blocks have shard_size rows except the tail."""
    s = shard_size(n, world_size)
    lo = min(rank * s, n)
    return lo, min(lo + s, n)


def all_gather_rows(local, n_total, group=None):
    """Row shards (shard_range layout) -> the full [n_total, ...] tensor on every rank."""
    rank, ws = world(group)
    if ws == 1:
        return local
    s = shard_size(n_total, ws)
    if local.size(0) != s:                                  # pad the tail shard
        pad = torch.zeros((s - local.size(0),) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        local = torch.cat([local, pad])
    out = torch.empty((s * ws,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


def all_gather_cols(local, group=None):
    """[nq, c] per rank -> [nq, c * world] (per-shard candidate lists side by side)."""
    rank, ws = world(group)
    if ws == 1:
        return local
    nq, c = local.shape
    out = torch.empty((ws * nq, c), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out.view(ws, nq, c).permute(1, 0, 2).reshape(nq, ws * c).contiguous()


def get_embeddings_sharded(model, x_local, sampler, num_items, num_neighbors=10, group=None):
    """PinSage.get_embeddings with rows split across ranks.  x_local: this rank's rows of the
    feature matrix (shard_range layout).  Returns this rank's rows of the embeddings."""
    from . import _native as N
    from . import kernels as K
    from . import neighbor_lists as NL
    rank, ws = world(group)
    lo, hi = shard_range(num_items, rank, ws)
    dev = model._device()
    xd = N.dev_tensor(x_local, torch.float32, dev)
    nodes = torch.arange(lo, hi, dtype=torch.int32, device=dev)
    batches = []
    for _ in range(model.num_layers):                       # same epochs on every rank
        ids, _c, w, nv = sampler._sample(nodes, num_neighbors, check=False)
        batches.append(NL.from_walk(ids, w, nv))
    P = lambda lin: (lin.weight, lin.bias)     # Parameter objects: identity keys the TF32 weight cache
    RND = 0 if model.precision == N.PREC_FP32 else N.EPI_ROUND_TF32     # see PinSage.forward
    PRE = 0 if model.precision == N.PREC_FP32 else N.IN_A1_TF32
    h_loc = K.gather_dense(xd, *P(model.input_proj), flags=N.EPI_RELU | RND, precision=model.precision)
    for i in range(model.num_layers):
        h_full = all_gather_rows(h_loc, num_items, group)   # the one exchange per layer
        wf, bf = model._folded_layer(i)
        if model.precision != N.PREC_FP32 and not model.fuse_pool:      # see PinSage.forward
            h_neigh = K.pool(h_full, *batches[i].as_args(), N.POOL_PINSAGE | N.POOL_ROUND_TF32)
            h_loc = K.gather_dense(h_full[lo:hi], wf, bf, a2=h_neigh,
                                   flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE | N.IN_A2_TF32,
                                   precision=model.precision)
            continue
        h_loc = K.gather_dense(h_full[lo:hi], wf, bf, pool_x=h_full, lists=batches[i].as_args(),
                               pool_mode=N.POOL_PINSAGE, flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE,
                               precision=model.precision)
    return K.gather_dense(h_loc, *P(model.output_proj), flags=N.EPI_L2NORM | PRE,
                          precision=model.precision)


def exact_search_item_sharded(queries, items_local, item_offset, k, metric, exclude_ids=None,
                              group=None):
    """Every rank scores all queries against its item shard; lists are all-gathered and merged."""
    from . import _native as N
    from . import kernels as K
    s, i = K.topk(queries, items_local, k, metric, exclude_ids=exclude_ids, id_offset=item_offset)
    rank, ws = world(group)
    if ws == 1:
        return s, i
    return K.topk_merge(all_gather_cols(s, group), all_gather_cols(i, group), k,
                        largest=metric == N.METRIC_IP)


def search_query_sharded(search_fn, queries, group=None):
    """search_fn(q_local) -> (scores [n,k], ids [n,k]) on this rank's block of queries; results
    of all ranks are all-gathered in query order."""
    rank, ws = world(group)
    n = queries.size(0)
    lo, hi = shard_range(n, rank, ws)
    s, i = search_fn(queries[lo:hi])
    return all_gather_rows(s, n, group), all_gather_rows(i, n, group)
