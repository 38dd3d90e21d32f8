"""Ragged python neighbour lists <-> padded device tensors.

The reference passes neighbourhoods around as ``list[list[int]]`` / ``list[list[float]]``
(utils/random_walk.py:134-142 -> model/pinsage.py:108).  The kernels take left-aligned padded
int32/float32 matrices plus per-row lengths; ``NeighborBatch`` is that device-resident form and
is what the tensor fast paths exchange (no Python lists, no host round trip).
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import chain
from typing import Optional

import numpy as np
import torch

_I32_MAX = np.iinfo(np.int32).max
_I32_MIN = np.iinfo(np.int32).min


@dataclass
class NeighborBatch:
    ids: torch.Tensor                    # int32 [n, T], entries >= list_len are padding
    weights: Optional[torch.Tensor]      # float32 [n, T] or None
    list_len: torch.Tensor               # int32 [n]
    weight_len: Optional[torch.Tensor]   # int32 [n] or None (= list_len)

    def as_args(self):
        return (self.ids, self.weights, self.list_len, self.weight_len)

    def __len__(self):
        return self.ids.size(0)


def _ragged_to_padded(rows, dtype, fill):
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    T = max(int(lens.max()) if len(rows) else 0, 1)
    out = np.full((len(rows), T), fill, dtype=dtype)
    total = int(lens.sum())
    if total:
        flat = np.fromiter(chain.from_iterable(rows), dtype=np.float64 if dtype == np.float32
                           else np.int64, count=total)
        if dtype == np.int32:
            flat = np.clip(flat, _I32_MIN, _I32_MAX)  # ids beyond int32 are out of range either way
        r = np.repeat(np.arange(len(rows)), lens)
        c = np.arange(total) - np.repeat(np.cumsum(lens) - lens, lens)
        out[r, c] = flat.astype(dtype)
    return out, lens.astype(np.int32)


def _wrap_negative(ids, lens, num_rows, check_upper):
    """Python indexing semantics of the reference's ``x[list_of_ids]`` on the host copy of the lists:
    an id in [-num_rows, -1] addresses row id + num_rows, an id below -num_rows raises IndexError
    (model/pinsage.py:146, model/layers.py:119/181/227, model/aggregators.py:33/71/258).  Every
    validity filter of the reference (``idx <= max_idx``, ``n < x.size(0)``) lets negative ids
    through.  check_upper: the aggregators do not filter at all, so ids >= num_rows raise too."""
    valid = np.arange(ids.shape[1])[None, :] < lens[:, None]
    neg = valid & (ids < 0)
    if (neg & (ids < -num_rows)).any() or (check_upper and (valid & (ids >= num_rows)).any()):
        raise IndexError(f"index out of range for a matrix with {num_rows} rows")
    if neg.any():
        ids = np.where(neg, ids + num_rows, ids).astype(np.int32)
    return ids


def pad_lists(neighbors, weights, device, bare_int=False, num_rows=None, check_upper=False):
    """Host lists -> NeighborBatch on `device`.  ``zip`` semantics: rows = min(len(n), len(w)).
    bare_int=True applies model/pinsage.py:110-112 (an int entry means [int] with weight 1).
    num_rows: rows of the matrix the ids index; negative ids then follow Python indexing
    (see _wrap_negative).  Without it negative ids are dropped by the kernels."""
    if isinstance(neighbors, NeighborBatch):
        return neighbors
    n = len(neighbors) if weights is None else min(len(neighbors), len(weights))
    nbrs = list(neighbors[:n])
    wts = None if weights is None else list(weights[:n])
    if bare_int:
        for i, nb in enumerate(nbrs):
            if isinstance(nb, (int, np.integer)):
                nbrs[i] = [int(nb)]
                if wts is not None:
                    wts[i] = [1.0]
    ids, lens = _ragged_to_padded(nbrs, np.int32, -1)
    if num_rows is not None:
        ids = _wrap_negative(ids, lens, int(num_rows), check_upper)
    t_ids = torch.from_numpy(ids).to(device, non_blocking=True)
    t_len = torch.from_numpy(lens).to(device, non_blocking=True)
    if wts is None:
        return NeighborBatch(t_ids, None, t_len, None)
    w, wlens = _ragged_to_padded(wts, np.float32, 0.0)
    T = max(ids.shape[1], w.shape[1])
    if w.shape[1] != T:
        w = np.pad(w, ((0, 0), (0, T - w.shape[1])))
    if ids.shape[1] != T:
        t_ids = torch.from_numpy(np.pad(ids, ((0, 0), (0, T - ids.shape[1])), constant_values=-1)
                                 ).to(device, non_blocking=True)
    return NeighborBatch(t_ids, torch.from_numpy(w).to(device, non_blocking=True), t_len,
                         torch.from_numpy(wlens).to(device, non_blocking=True))


def from_walk(ids, weights, nvalid):
    """Kernel outputs of walk_topt -> NeighborBatch (already on the device, zero copies)."""
    return NeighborBatch(ids, weights, nvalid, None)


def to_lists(ids, counts, nvalid):
    """Device walk results -> the reference's (list[list[int]], list[list[float]]).
    weights = count / sum(kept counts) in float64, bit-identical to the python int division at
    utils/random_walk.py:113-115."""
    ids_h = ids.cpu().numpy()
    cnt_h = counts.cpu().numpy().astype(np.float64)
    nv = nvalid.cpu().numpy()
    tot = cnt_h.sum(axis=1, keepdims=True)
    w_h = np.divide(cnt_h, tot, out=np.zeros_like(cnt_h), where=tot > 0)
    ids_l, w_l = ids_h.tolist(), w_h.tolist()
    return ([row[:k] for row, k in zip(ids_l, nv)], [row[:k] for row, k in zip(w_l, nv)])
