"""Drop-in for the reference's ``model/aggregators.py`` (gather + weighted-reduce family).

MeanAggregator (:5-39), WeightedAggregator (:41-91): pb200_pool, no id filtering -- like the
reference's ``features[node_neighbors]`` an out-of-range id raises IndexError.
ImportanceAggregator (:213-287): Linear on gathered rows -> importance-weighted sum ->
LayerNorm.  Because the weights are normalised to sum 1, sum_j w_j (W x_j + b) =
W (sum_j w_j x_j) + b, so it runs as one fused launch: pooled A tile -> dense -> LayerNorm
epilogue; rows with an empty list are zeros (no LayerNorm), as in the reference (:248-251).
AttentionAggregator / MaxPoolingAggregator (per-neighbour MLPs, selectable only through an
unused config string) are outside the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import kernels as K
from .. import neighbor_lists as NL


def _checked_lists(features, neighbors, weights, dev):
    """``features[node_neighbors]`` semantics (reference :33, :71, :258): no filtering -- ids outside
    [-M, M) raise IndexError, negative ids index from the end.  Checked on the host lists (no device
    sync); a NeighborBatch of device tensors is checked on the device."""
    M = features.size(0)
    if not isinstance(neighbors, NL.NeighborBatch):
        return NL.pad_lists(neighbors, weights, dev, num_rows=M, check_upper=True)
    nb = neighbors
    valid = torch.arange(nb.ids.size(1), device=dev)[None, :] < nb.list_len[:, None]
    if bool((((nb.ids >= M) | (nb.ids < -M)) & valid).any()):
        raise IndexError(f"index out of range for features with {M} rows")
    if bool(((nb.ids < 0) & valid).any()):
        nb = NL.NeighborBatch(torch.where(valid & (nb.ids < 0), nb.ids + M, nb.ids), nb.weights,
                              nb.list_len, nb.weight_len)
    return nb


class MeanAggregator(nn.Module):
    def forward(self, features, neighbors):
        dev = N.device_of(features)
        fd = N.dev_tensor(features, torch.float32, dev)
        nb = _checked_lists(fd, neighbors, None, dev)
        out = K.pool(fd, *nb.as_args(), N.POOL_MEAN)
        return out if features.is_cuda else out.to(features.device)


class WeightedAggregator(nn.Module):
    def forward(self, features, neighbors, weights):
        dev = N.device_of(features)
        fd = N.dev_tensor(features, torch.float32, dev)
        nb = _checked_lists(fd, neighbors, weights, dev)
        out = K.pool(fd, *nb.as_args(), N.POOL_AGGREGATOR)
        return out if features.is_cuda else out.to(features.device)


class ImportanceAggregator(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.transform = nn.Linear(in_channels, out_channels)
        self.norm = nn.LayerNorm(out_channels)

    def forward(self, features, neighbors, importance_weights):
        dev = N.device_of(self.transform.weight)
        fd = N.dev_tensor(features, torch.float32, dev)
        nb = _checked_lists(fd, neighbors, importance_weights, dev)
        out = K.gather_dense(None, self.transform.weight.detach(), self.transform.bias.detach(),
                             pool_x=fd, lists=nb.as_args(), pool_mode=N.POOL_AGGREGATOR,
                             flags=N.EPI_LAYERNORM, ln_gamma=self.norm.weight.detach(),
                             ln_beta=self.norm.bias.detach(), n=len(nb))
        out.masked_fill_((nb.list_len == 0)[:, None], 0.0)      # empty list -> zeros (:248-251)
        return out if features.is_cuda else out.to(features.device)


class AttentionAggregator(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("AttentionAggregator (reference :93-160) is outside the PinSage "
                                  "importance-pooling hot path (SURVEY.md section 2 row 4)")


class MaxPoolingAggregator(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("MaxPoolingAggregator (reference :162-211) is outside the PinSage "
                                  "importance-pooling hot path (SURVEY.md section 2 row 4)")
