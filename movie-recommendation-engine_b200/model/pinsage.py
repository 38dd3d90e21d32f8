"""Drop-in for the reference's ``model/pinsage.py`` (PinSage, ImportancePooling), B200-native.

State-dict compatible with the reference (``input_proj``, ``convs.{i}.{lin_self,lin_neigh,
lin_update}``, ``output_proj``; checkpoints/best_model.pt loads unchanged).  The modules only
hold parameters: all arithmetic runs in libpinsage_b200.so.

forward (reference model/pinsage.py:186-251), importance branch, per layer:
    h_neigh = ImportancePooling(h, nbrs, wts)           # :232   gather + weighted sum
    h_self  = lin_self(h)                               # :235
    h = normalize(relu(lin_update([h_self | h_neigh]))) # :238-240
is ONE launch of pb200_gather_dense: the A tile [h | h_neigh] is built in shared memory by the
gather stage and multiplied by the folded weight  W' = [W_u1 W_s | W_u2],
b' = W_u1 b_s + b_u  (exact algebra); ``fold=False`` keeps the two GEMMs apart
for bisecting.  ``lin_neigh`` exists for state-dict compatibility and, as in the reference's
importance branch, is never used.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import _native as N
from .. import kernels as K
from .. import neighbor_lists as NL


class GraphConv(nn.Module):
    """Parameter container mirroring reference model/pinsage.py:8-29.  The PyG ``edge_index``
    message-passing branch (:31-92) is out of scope (needs torch_geometric; no reference script
    reaches it with a valid tensor)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_self = nn.Linear(in_channels, out_channels)
        self.lin_neigh = nn.Linear(in_channels, out_channels)
        self.lin_update = nn.Linear(2 * out_channels, out_channels)

    def forward(self, x, edge_index=None, edge_weight=None, importance_weights=None):
        if edge_index is not None:
            raise NotImplementedError("GraphConv.propagate (torch_geometric branch, reference "
                                      "model/pinsage.py:53) is outside the B200 hot path")
        dev = N.device_of(self.lin_self.weight)
        x = N.dev_tensor(x, torch.float32, dev)
        h_self = K.gather_dense(x, self.lin_self.weight, self.lin_self.bias)
        zeros = torch.zeros_like(h_self)                       # :50
        return K.gather_dense(h_self, self.lin_update.weight, self.lin_update.bias, a2=zeros,
                              flags=N.EPI_RELU | N.EPI_L2NORM)


class ImportancePooling(nn.Module):
    """reference model/pinsage.py:94-150.  Accepts the reference's python lists (bare ints
    allowed, :110-112) or a NeighborBatch of device tensors."""

    def forward(self, x, neighbors, weights=None):
        dev = N.device_of(x)
        xd = N.dev_tensor(x, torch.float32, dev)
        nb = NL.pad_lists(neighbors, weights, dev, bare_int=True, num_rows=xd.size(0))
        out = K.pool(xd, *nb.as_args(), N.POOL_PINSAGE)
        return out if not isinstance(x, torch.Tensor) or x.is_cuda else out.to(x.device)


class PinSage(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=2):
        super().__init__()
        self.num_layers = num_layers
        self.input_proj = nn.Linear(in_channels, hidden_channels)
        self.convs = nn.ModuleList(GraphConv(hidden_channels, hidden_channels)
                                   for _ in range(num_layers))
        self.importance_pooling = ImportancePooling()
        self.output_proj = nn.Linear(hidden_channels, out_channels)
        self.fold = True                 # fold lin_self into lin_update (one GEMM per layer)
        # Tensor-core path: True gathers the neighbour rows inside the GEMM kernel's producer
        # warps (one launch per layer); False runs the pooling kernel first and streams both
        # halves of K through TMA (two launches, measured faster: see DESIGN.md section 4).
        self.fuse_pool = False
        # PREC_AUTO: tcgen05 kind::tf32 tensor cores (fp32 accumulate) where the layer shape is
        # covered, CUDA-core fp32 otherwise; PREC_FP32 forces the exact-fp32 kernels.
        self.precision = N.PREC_AUTO
        self._folded = {}

    # ---- weight preparation (once per parameter version) ---------------------------------
    def _device(self):
        return N.device_of(self.input_proj.weight)

    def _folded_layer(self, i):
        conv = self.convs[i]
        ws, bs = conv.lin_self.weight, conv.lin_self.bias
        wu, bu = conv.lin_update.weight, conv.lin_update.bias
        key = (ws._version, bs._version, wu._version, bu._version, ws.data_ptr(), wu.data_ptr())
        hit = self._folded.get(i)
        if hit is None or hit[0] != key:
            H = ws.size(0)
            wu1 = wu.detach()[:, :H].contiguous()
            # W_u1 @ W_s and W_u1 @ b_s through the library's own dense kernel
            w1 = K.gather_dense(wu1, ws.detach().t().contiguous())
            bf = K.gather_dense(wu1, bs.detach()[None, :].contiguous(), bias=None)[:, 0] + bu.detach()
            wf = torch.cat([w1, wu.detach()[:, H:]], dim=1).contiguous()
            hit = (key, wf, bf.contiguous())
            self._folded[i] = hit
        return hit[1], hit[2]

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, x, edge_index=None, sampled_neighbors=None, importance_weights=None, *,
                out=None):
        """reference model/pinsage.py:186-251.  ``sampled_neighbors`` / ``importance_weights``:
        the reference's per-layer python lists, or a list of NeighborBatch (weights ignored).
        ``out``: optional (pinned) host or device tensor that receives the embeddings."""
        if edge_index is not None:
            raise NotImplementedError("the edge_index / torch_geometric branch (reference "
                                      "model/pinsage.py:243-245) is outside the B200 hot path")
        dev = self._device()
        in_dev = x.device if isinstance(x, torch.Tensor) else torch.device("cpu")
        xd = N.dev_tensor(x, torch.float32, dev)
        prec = self.precision
        P = lambda lin: (lin.weight, lin.bias)     # Parameter objects: identity keys the TF32 weight cache
        # tensor-core path: intermediate activations are stored rounded to TF32 (exactly what
        # the MMA would read anyway), which lets the next layer stream them with cp.async
        RND = 0 if prec == N.PREC_FP32 else N.EPI_ROUND_TF32
        PRE = 0 if prec == N.PREC_FP32 else N.IN_A1_TF32
        h = K.gather_dense(xd, *P(self.input_proj), flags=N.EPI_RELU | RND, precision=prec)   # :202

        tensor_path = isinstance(sampled_neighbors, (list, tuple)) and len(sampled_neighbors) > 0 \
            and isinstance(sampled_neighbors[0], NL.NeighborBatch)
        if sampled_neighbors is None or (importance_weights is None and not tensor_path):
            for i in range(self.num_layers):                                               # :205-214
                h = K.gather_dense(h, *P(self.convs[i].lin_self), flags=N.EPI_RELU | RND | PRE,
                                   precision=prec)
        else:
            per_layer = tensor_path or (isinstance(sampled_neighbors, list) and
                                        isinstance(importance_weights, list))
            shared = None
            for i in range(self.num_layers):                                               # :222-240
                if per_layer and len(sampled_neighbors) > i:
                    nb = sampled_neighbors[i] if tensor_path else \
                        NL.pad_lists(sampled_neighbors[i], importance_weights[i], dev, bare_int=True,
                                     num_rows=h.size(0))
                else:                                                                      # :226-229
                    if shared is None:
                        shared = NL.pad_lists(sampled_neighbors, importance_weights, dev,
                                              bare_int=True, num_rows=h.size(0))
                    nb = shared
                if len(nb) != h.size(0):
                    raise RuntimeError(f"layer {i}: {len(nb)} neighbour lists for {h.size(0)} rows "
                                       "(torch.cat would fail in the reference, :238)")
                if self.fold and prec != N.PREC_FP32 and not self.fuse_pool:
                    wf, bf = self._folded_layer(i)
                    h_neigh = K.pool(h, *nb.as_args(), N.POOL_PINSAGE | N.POOL_ROUND_TF32)
                    h = K.gather_dense(h, wf, bf, a2=h_neigh,
                                       flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE | N.IN_A2_TF32,
                                       precision=prec)
                elif self.fold:
                    wf, bf = self._folded_layer(i)
                    h = K.gather_dense(h, wf, bf, pool_x=h, lists=nb.as_args(),
                                       pool_mode=N.POOL_PINSAGE,
                                       flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE, precision=prec)
                else:
                    h_self = K.gather_dense(h, *P(self.convs[i].lin_self), flags=RND | PRE, precision=prec)
                    h = K.gather_dense(h_self, *P(self.convs[i].lin_update), pool_x=h,
                                       lists=nb.as_args(), pool_mode=N.POOL_PINSAGE,
                                       flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE, precision=prec)
        emb = K.gather_dense(h, *P(self.output_proj), flags=N.EPI_L2NORM | PRE, precision=prec)  # :248-249
        if out is not None:
            out.copy_(emb, non_blocking=True)
            if not out.is_cuda:
                torch.cuda.current_stream(dev).synchronize()
            return out
        return emb if in_dev.type == "cuda" else emb.to(in_dev)

    def get_embeddings(self, x, random_walk_sampler, num_neighbors=10, *, out=None):
        """reference model/pinsage.py:253-280: resample neighbours per layer, then forward.
        With this package's sampler everything stays on the device (NeighborBatch); any other
        sampler object goes through its list API exactly like the reference.  A pinned host
        ``x`` is uploaded on a side stream while the walk kernels (which do not need the
        features) run; ``out`` optionally receives the embeddings (pinned host or device)."""
        M = x.size(0)
        if hasattr(random_walk_sampler, "batch_sample_neighbors_tensor"):
            dev = self._device()
            if M > random_walk_sampler.csr.num_nodes:
                raise IndexError("list index out of range")
            in_dev = x.device
            main = torch.cuda.current_stream(dev)
            if not x.is_cuda and x.is_pinned():
                if getattr(self, "_copy_stream", None) is None:
                    self._copy_stream = torch.cuda.Stream(dev)
                with torch.cuda.stream(self._copy_stream):
                    xd = x.to(dev, dtype=torch.float32, non_blocking=True)
                    uploaded = torch.cuda.Event()
                    uploaded.record(self._copy_stream)
                xd.record_stream(main)
            else:
                xd, uploaded = x, None
            nodes = torch.arange(M, dtype=torch.int32, device=dev)
            batches = random_walk_sampler.sample_layers(nodes, num_neighbors, self.num_layers)   # one launch
            if uploaded is not None:
                main.wait_event(uploaded)
            emb = self.forward(xd, None, batches, None, out=out)
            if out is not None or in_dev.type == "cuda":
                return emb
            return emb.to(in_dev)
        all_n, all_w = [], []
        nodes = list(range(M))
        for _ in range(self.num_layers):
            n_, w_ = random_walk_sampler.batch_sample_neighbors(nodes, num_neighbors)
            all_n.append(n_)
            all_w.append(w_)
        return self.forward(x, edge_index=None, sampled_neighbors=all_n, importance_weights=all_w)
