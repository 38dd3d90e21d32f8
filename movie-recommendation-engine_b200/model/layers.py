"""Drop-in for the reference's ``model/layers.py``, B200-native (inference).

GraphConvLayer (reference model/layers.py:5-77) collapses to ONE fused launch in eval mode (training mode, with
batch statistics, takes two launches and a column reduction: _forward_batch_stats):
    out = normalize(relu(BN_eval(linear_out([linear_self(x) | linear_neigh(neigh_x)]))))
      = normalize(relu([x | neigh_x] . W'^T + b'))
with W' = diag(s) [W_o1 W_s | W_o2 W_n],  b' = s * (W_o1 b_s + W_o2 b_n + b_o - mean) + beta,
s = gamma / sqrt(var + eps); the BatchNorm factor is dropped for single-row inputs (:68-69).
The pooling layers (:79-236) keep their list-based call signatures and differing id/weight
semantics and run as pb200_pool.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import kernels as K
from .. import neighbor_lists as NL


class GraphConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.linear_self = nn.Linear(in_channels, out_channels)
        self.linear_neigh = nn.Linear(in_channels, out_channels)
        self.linear_out = nn.Linear(2 * out_channels, out_channels)
        self.bn = nn.BatchNorm1d(out_channels)
        self._init_weights()
        self._folded = {}

    def _init_weights(self):                     # reference :34-42
        for lin in (self.linear_self, self.linear_neigh, self.linear_out):
            nn.init.xavier_uniform_(lin.weight)
        for lin in (self.linear_self, self.linear_neigh, self.linear_out):
            nn.init.zeros_(lin.bias)

    def _fold(self, with_bn):
        params = [self.linear_self.weight, self.linear_self.bias, self.linear_neigh.weight,
                  self.linear_neigh.bias, self.linear_out.weight, self.linear_out.bias,
                  self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        key = tuple(p._version for p in params) + tuple(p.data_ptr() for p in params)
        hit = self._folded.get(with_bn)
        if hit is None or hit[0] != key:
            ws, bs, wn, bn_, wo, bo = (p.detach() for p in params[:6])
            O = ws.size(0)
            wo1, wo2 = wo[:, :O].contiguous(), wo[:, O:].contiguous()
            w1 = K.gather_dense(wo1, ws.t().contiguous())
            w2 = K.gather_dense(wo2, wn.t().contiguous())
            b = K.gather_dense(wo1, bs[None, :].contiguous())[:, 0] + \
                K.gather_dense(wo2, bn_[None, :].contiguous())[:, 0] + bo
            wf = torch.cat([w1, w2], dim=1)
            if with_bn:
                s = self.bn.weight.detach() / torch.sqrt(self.bn.running_var + self.bn.eps)
                wf = wf * s[:, None]
                b = (b - self.bn.running_mean) * s + self.bn.bias.detach()
            hit = (key, wf.contiguous(), b.contiguous())
            self._folded[with_bn] = hit
        return hit[1], hit[2]

    def forward(self, x, neigh_x):
        dev = N.device_of(self.linear_self.weight)
        xd = N.dev_tensor(x, torch.float32, dev)
        nd = N.dev_tensor(neigh_x, torch.float32, dev)
        if self.training and x.size(0) > 1:
            out = self._forward_batch_stats(xd, nd)
            return out if x.is_cuda else out.to(x.device)
        wf, bf = self._fold(with_bn=x.size(0) > 1)
        out = K.gather_dense(xd, wf, bf, a2=nd, flags=N.EPI_RELU | N.EPI_L2NORM)
        return out if x.is_cuda else out.to(x.device)

    def _forward_batch_stats(self, xd, nd):
        """Training mode (what a freshly constructed module is in, reference layers.py:68-69): BatchNorm1d
        normalises with the statistics of THIS batch and moves its running statistics.  Two launches of the
        exact-fp32 dense kernel around one column reduction: y = [x | neigh] W'^T + b' (BatchNorm not folded), then
        out = normalize(relu(y * s + t)) as a diagonal GEMM with s = gamma / sqrt(var_batch + eps),
        t = beta - mean_batch * s.  No gradients: this package is inference-only."""
        wf, bf = self._fold(with_bn=False)
        y = K.gather_dense(xd, wf, bf, a2=nd, flags=0)
        var, mean = torch.var_mean(y, dim=0, unbiased=False)
        bn = self.bn
        with torch.no_grad():
            if bn.track_running_stats and bn.running_mean is not None:
                n = y.size(0)
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
                bn.running_var.mul_(1 - m).add_(var * (n / (n - 1)), alpha=m)        # running_var is unbiased
            s = bn.weight.detach() / torch.sqrt(var + bn.eps)
            t = bn.bias.detach() - mean * s
        return K.gather_dense(y, torch.diag(s).contiguous(), t.contiguous(), flags=N.EPI_RELU | N.EPI_L2NORM)


def _pool(x, neighbors, weights, mode):
    dev = N.device_of(x)
    xd = N.dev_tensor(x, torch.float32, dev)
    nb = NL.pad_lists(neighbors, weights, dev, num_rows=xd.size(0))
    out = K.pool(xd, *nb.as_args(), mode)
    return out if x.is_cuda else out.to(x.device)


class ImportancePoolingLayer(nn.Module):
    """reference :79-133: ids >= x.size(0) dropped, weights = HEAD of the weight list."""

    def forward(self, x, neighbors, weights):
        return _pool(x, neighbors, weights, N.POOL_LAYERS)


class WeightedMeanPoolingLayer(nn.Module):
    """reference :135-195: as above; weights=None (or fewer weight rows than nodes) -> mean."""

    def forward(self, x, neighbors, weights=None):
        if weights is None:
            return _pool(x, neighbors, None, N.POOL_MEAN)
        if len(weights) < len(neighbors):       # rows i >= len(weights) fall back to mean (:174)
            n = len(weights)
            head = _pool(x, neighbors[:n], weights, N.POOL_LAYERS)
            tail = _pool(x, neighbors[n:], None, N.POOL_MEAN)
            return torch.cat([head, tail], dim=0)
        return _pool(x, neighbors, weights[:len(neighbors)], N.POOL_LAYERS)


class MaxPoolingLayer(nn.Module):
    """reference :197-236."""

    def forward(self, x, neighbors):
        return _pool(x, neighbors, None, N.POOL_MAX)
