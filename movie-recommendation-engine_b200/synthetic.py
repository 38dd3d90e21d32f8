"""Seeded synthetic MovieLens-shaped inputs.

Produces exactly the layout ``MovieLensDataset.build_graph`` emits
(reference data/dataset.py:105-116): unified ids (movies [0,M), users [M,M+U)),
``edge_index = [[u+M | i], [i | u+M]]`` int64 and ``edge_weights = [r | r]`` float32 with
ratings that are multiples of 0.5.  Host-side numpy; used by tests and bench.py only.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (movies, users, ratings, F, H, E, layers)
    "C1": (2_000, 5_000, 100_000, 64, 64, 64, 2),
    "C2": (62_423, 162_541, 25_000_095, 128, 256, 128, 2),
}


def bipartite_graph(M, U, R, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    p_i = 1.0 / (np.arange(M, dtype=np.float64) + 50.0)
    p_i /= p_i.sum()
    p_u = rng.lognormal(mean=0.0, sigma=1.2, size=U)
    p_u /= p_u.sum()
    n_draw = int(1.25 * R)
    ci = np.cumsum(p_i); ci /= ci[-1]
    cu = np.cumsum(p_u); cu /= cu[-1]
    keys = np.empty(0, dtype=np.int64)
    # draw, dedup, top up until R unique pairs exist (popular pairs collide often)
    for _ in range(8):
        items = np.searchsorted(ci, rng.random(n_draw), side="right").astype(np.int64)
        users = np.searchsorted(cu, rng.random(n_draw), side="right").astype(np.int64)
        np.minimum(items, M - 1, out=items)
        np.minimum(users, U - 1, out=users)
        keys = np.unique(np.concatenate([keys, users * M + items]))
        if keys.size >= R:
            break
        n_draw = int(1.5 * (R - keys.size)) + 1024
    rng.shuffle(keys)
    keys = keys[:R]
    users = keys // M
    items = keys % M
    rating = (0.5 * rng.integers(1, 11, size=keys.size)).astype(np.float32)
    u = users + M
    edge_index = np.stack([np.concatenate([u, items]), np.concatenate([items, u])]).astype(np.int64)
    edge_weights = np.concatenate([rating, rating]).astype(np.float32)
    return edge_index, edge_weights


def features(M, F, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(M, F, generator=g)


def spread_embeddings(N, d, seed=1, clusters=1024, noise=0.3):
    """Set B of SURVEY.md 8(d): clustered, L2-normalised, well spread."""
    import torch
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(clusters, d, generator=g)
    which = torch.randint(0, clusters, (N,), generator=g)
    x = centres[which] + noise * torch.randn(N, d, generator=g)
    return torch.nn.functional.normalize(x, dim=1)
