"""Drop-in for the exact-search part of the reference's ``utils/evaluation.py``.

generate_recommendations (reference :106-132): sim = q . E^T, sim[query] = -inf, topk ->
pb200_topk (inner product, query excluded).  ``generate_recommendations_batch`` is the same
for many queries in one launch.  Hit-rate / MRR / evaluate_embeddings (:5-104, SURVEY.md 8(f)
N2) are functions of the rank of the ground truth in the query's similarity order, which
pb200_rank_of_target computes for all pairs in one pass.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .. import kernels as K


def generate_recommendations_batch(item_embeddings, query_indices, k=10, exclude_query=True):
    dev = N.device_of(item_embeddings)
    emb = N.dev_tensor(item_embeddings, torch.float32, dev)
    qi = torch.as_tensor(query_indices, dtype=torch.int64).reshape(-1)
    if qi.numel() and (int(qi.min()) < -emb.size(0) or int(qi.max()) >= emb.size(0)):
        raise IndexError("index out of range")
    qi = torch.where(qi < 0, qi + emb.size(0), qi).to(dev)
    _scores, ids = K.topk(emb[qi].contiguous(), emb, k, N.METRIC_IP,
                          exclude_ids=qi.to(torch.int32) if exclude_query else None)
    return ids.cpu().numpy().astype(np.int64)


def generate_recommendations(item_embeddings, query_idx, k=10, exclude_query=True):
    if k > item_embeddings.size(0):
        raise RuntimeError("selected index k out of range")       # what torch.topk raises
    return generate_recommendations_batch(item_embeddings, [int(query_idx)], k, exclude_query)[0]


def _ranks(item_embeddings, query_indices, ground_truth_indices):
    """1-based rank of every ground-truth item in its query's descending similarity order
    (pb200_rank_of_target) -- the quantity both reference metrics are functions of."""
    dev = N.device_of(item_embeddings)
    emb = N.dev_tensor(item_embeddings, torch.float32, dev)
    n = emb.size(0)
    q = torch.as_tensor(np.asarray(query_indices), dtype=torch.int64).reshape(-1)
    g = torch.as_tensor(np.asarray(ground_truth_indices), dtype=torch.int64).reshape(-1)
    for t in (q, g):
        if t.numel() and (int(t.min()) < -n or int(t.max()) >= n):
            raise IndexError("index out of range")
    q = torch.where(q < 0, q + n, q).to(torch.int32)
    g = torch.where(g < 0, g + n, g).to(torch.int32)
    return K.rank_of_target(emb, q, g).cpu().numpy().astype(np.int64)


def calculate_hit_rate(item_embeddings, query_indices, ground_truth_indices, k=500):
    """reference :5-36: fraction of queries whose ground truth is in the top-k of q . E^T (the query
    itself is not excluded).  One rank kernel instead of a matmul + topk per pair."""
    total = len(query_indices)
    if k > item_embeddings.shape[0]:
        raise RuntimeError("selected index k out of range")        # what torch.topk raises (:28)
    if total == 0:
        raise ZeroDivisionError("division by zero")                 # hits / total (:35)
    ranks = _ranks(item_embeddings, query_indices, ground_truth_indices)
    return int((ranks <= k).sum()) / total


def calculate_mrr(item_embeddings, query_indices, ground_truth_indices, scale=100):
    """reference :38-73: mean of 1 / (rank / scale)."""
    ranks = _ranks(item_embeddings, query_indices, ground_truth_indices)
    return np.mean([1.0 / (int(r) / scale) for r in ranks])


def evaluate_embeddings(item_embeddings, test_data, k_values=[10, 50, 100, 500]):
    """reference :75-104 -- same result dict; the ranks are computed once for all metrics."""
    positive_pairs = test_data['positive_pairs']
    pairs = positive_pairs.cpu().numpy() if isinstance(positive_pairs, torch.Tensor) else np.asarray(positive_pairs)
    total = len(pairs)
    ranks = _ranks(item_embeddings, pairs[:, 0], pairs[:, 1])
    results = {}
    for k in k_values:
        if k > item_embeddings.shape[0]:
            raise RuntimeError("selected index k out of range")
        results[f'hit_rate@{k}'] = int((ranks <= k).sum()) / total
    results['mrr'] = np.mean([1.0 / (int(r) / 100) for r in ranks])
    return results
