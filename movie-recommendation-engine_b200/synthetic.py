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
    # BASELINE.json configs[4]: generated on the device (bipartite_graph_device); scale it with c5_config()
    "C5": (10_000_000, 50_000_000, 2_000_000_000, 128, 256, 128, 3),
}


def c5_config(scale=1.0):
    """C5 with items, users and ratings multiplied by `scale` (mean degrees unchanged: items 200, users 40)."""
    M, U, R, F, H, E, layers = CONFIGS["C5"]
    return (max(int(M * scale), 64), max(int(U * scale), 64), max(int(R * scale), 1024), F, H, E, layers)


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


def bipartite_graph_device(M, U, R, seed=0, device="cuda"):
    """The generator above with the same formulas, on the device (torch's CUDA generator is Philox4x32-10): a
    2 G-rating graph cannot be drawn through numpy in reasonable time.  A different SAMPLE of
    the same distribution (not bit-identical to bipartite_graph); deterministic per (seed, GPU model).
    Returns (edge_index int64 [2, 2R], edge_weights float32 [2R]) on the device, build_graph's layout."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    p_i = 1.0 / (torch.arange(M, dtype=torch.float64, device=dev) + 50.0)
    ci = torch.cumsum(p_i, 0); ci /= ci[-1].clone()
    p_u = torch.exp(1.2 * torch.randn(U, generator=g, device=dev, dtype=torch.float64))
    cu = torch.cumsum(p_u, 0); cu /= cu[-1].clone()
    del p_i, p_u
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    n_draw = int(1.25 * R)
    for _ in range(8):
        chunks = []
        for lo in range(0, n_draw, 1 << 27):                    # bounded temporaries
            n = min(1 << 27, n_draw - lo)
            items = torch.searchsorted(ci, torch.rand(n, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=M - 1)
            users = torch.searchsorted(cu, torch.rand(n, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=U - 1)
            chunks.append(users * M + items)
            del items, users
        keys = torch.unique(torch.cat([keys] + chunks))
        del chunks
        if keys.numel() >= R:
            break
        n_draw = int(1.5 * (R - keys.numel())) + 1024
    order = torch.argsort(torch.rand(keys.numel(), generator=g, device=dev))      # shuffle
    keys = keys[order[:R]]
    del order, ci, cu
    users = keys // M + M
    items = keys % M
    rating = (0.5 * torch.randint(1, 11, (R,), generator=g, device=dev)).to(torch.float32)
    edge_index = torch.stack([torch.cat([users, items]), torch.cat([items, users])])
    return edge_index, torch.cat([rating, rating])


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
