"""Drop-in for the exact-search part of the reference's ``utils/evaluation.py``.

generate_recommendations (reference :106-132): sim = q . E^T, sim[query] = -inf, topk ->
pb200_topk (inner product, query excluded).  ``generate_recommendations_batch`` is the same
for many queries in one launch.  Hit-rate / MRR (:5-104) are metrics outside the first pass
(SURVEY.md 8(f) N2).
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


def calculate_hit_rate(*a, **k):
    raise NotImplementedError("hit-rate (reference :5-36) is outside the first pass (SURVEY 8(f) N2)")


def calculate_mrr(*a, **k):
    raise NotImplementedError("MRR (reference :38-73) is outside the first pass (SURVEY 8(f) N2)")
