"""Drop-in for the reference's ``utils/random_walk.py`` (RandomWalkSampler), B200-native.

Same constructor and method signatures as reference utils/random_walk.py:6-142; the Python
adjacency list becomes a device-resident stable CSR (pb200_csr_build) and the per-node
Python loops become one launch of the warp-per-start-node walk kernel (pb200_walk_topt).

Randomness: the reference draws from numpy's *global* MT19937 stream, so its output depends
on call order.  Here every draw is Philox4x32-10 keyed by ``seed`` with counter
(start node, walk, step, epoch): results do not depend on batching or on how start nodes are
sharded across GPUs.  ``epoch`` advances by one per sampling call, which plays the role of
the advancing global stream (two calls give independent samples, as in
PinSage.get_embeddings' per-layer resampling, model/pinsage.py:271-275).
"""
from __future__ import annotations

import torch

from .. import kernels as K
from .. import neighbor_lists as NL


class RandomWalkSampler:
    def __init__(self, edge_index, edge_weights=None, walk_length=2, num_walks=100, p=1.0, q=1.0,
                 *, seed=1234, device=None, num_nodes=None):
        self.edge_index = edge_index
        self.edge_weights = edge_weights
        self.walk_length = walk_length
        self.num_walks = num_walks
        self.p = p            # stored and unused, exactly like the reference (:27-28)
        self.q = q
        self.seed = int(seed)
        self.epoch = 0
        self._device = device
        self._num_nodes = num_nodes
        self._adj_list = None
        self._prepare_adjacency_list()

    # ---- S0 -------------------------------------------------------------------------
    def _prepare_adjacency_list(self):
        """reference :33-50 -- builds the device CSR instead of python lists."""
        self.csr = K.csr_build(self.edge_index, self.edge_weights, num_nodes=self._num_nodes,
                               device=self._device)
        self.device = self.csr.device

    @property
    def adj_list(self):
        """The reference attribute (list of [(dst, weight), ...] per node), materialised lazily
        from the CSR for callers that introspect it; the kernels never use it."""
        if self._adj_list is None:
            rp = self.csr.row_ptr.cpu().tolist()
            col = self.csr.col.cpu().tolist()
            cum = self.csr.cum.cpu()
            if self.csr.cum_kind == 0:
                cum = (cum.to(torch.int64) & 0xFFFFFFFF).tolist()
                scale = float(1 << self.csr.quant_shift)
            else:
                cum, scale = cum.tolist(), 1.0
            adj = []
            for v in range(self.csr.num_nodes):
                a, b = rp[v], rp[v + 1]
                prev, row = 0, []
                for e in range(a, b):
                    row.append((col[e], (cum[e] - prev) / scale))
                    prev = cum[e]
                adj.append(row)
            self._adj_list = adj
        return self._adj_list

    def _next_epoch(self):
        e = self.epoch
        self.epoch += 1
        return e

    def _check_nodes(self, nodes_t):
        if nodes_t.numel() and (int(nodes_t.min()) < 0 or int(nodes_t.max()) >= self.csr.num_nodes):
            raise IndexError("list index out of range")   # what adj_list[node] raises (:66)

    # ---- S1 -------------------------------------------------------------------------
    def _single_walk(self, start_node):
        """reference :52-83 -- one walk, returned as [start, v1, ..., vk] (k <= walk_length)."""
        start = torch.tensor([int(start_node)], dtype=torch.int32)
        self._check_nodes(start)
        *_ignored, trace = K.walk_topt(self.csr, start, 1, self.walk_length, 1, self.seed,
                                       self._next_epoch(), return_trace=True)
        return [int(start_node)] + [v for v in trace.view(-1).tolist() if v >= 0]

    # ---- S2 / S3 ----------------------------------------------------------------------
    def sample_neighbors(self, node_idx, num_neighbors=10):
        """reference :85-117."""
        n, w = self.batch_sample_neighbors([int(node_idx)], num_neighbors)
        return n[0], w[0]

    def batch_sample_neighbors(self, nodes, num_neighbors=10):
        """reference :119-142 -- returns (list[list[int]], list[list[float]])."""
        ids, counts, _w, nvalid = self._sample(nodes, num_neighbors)
        return NL.to_lists(ids, counts, nvalid)

    def batch_sample_neighbors_tensor(self, nodes, num_neighbors=10):
        """Fast path: the same sample as device tensors (NeighborBatch); no host round trip."""
        ids, _counts, weights, nvalid = self._sample(nodes, num_neighbors)
        return NL.from_walk(ids, weights, nvalid)

    def _sample(self, nodes, num_neighbors, epoch=None, check=True, epoch_dev=None):
        if not isinstance(nodes, torch.Tensor):
            nodes = torch.as_tensor(list(nodes), dtype=torch.int64)
        nodes = nodes.reshape(-1)
        if check:
            self._check_nodes(nodes)
        return K.walk_topt(self.csr, nodes, self.num_walks, self.walk_length, num_neighbors,
                           self.seed, self._next_epoch() if epoch is None else epoch, epoch_dev=epoch_dev)

    def sample_layers(self, nodes, num_neighbors, num_layers, epoch=None, epoch_dev=None):
        """`num_layers` consecutive sampling calls (one per conv layer, model/pinsage.py:271-275) as ONE
        kernel launch over (layer, start) pairs: returns [NeighborBatch] * num_layers, identical to
        num_layers calls of batch_sample_neighbors_tensor (epochs e, e+1, ...)."""
        if self.csr.meta is None:                       # float-weight graphs: no index, one launch per layer
            out = []
            for l in range(num_layers):
                ids, _c, w, nv = self._sample(nodes, num_neighbors, check=False, epoch_dev=epoch_dev,
                                              epoch=None if epoch is None else epoch + l)
                out.append(NL.from_walk(ids, w, nv))
            return out
        if epoch is None:
            epoch = self.epoch
            self.epoch += num_layers
        ids, _c, w, nv = K.walk_topt(self.csr, nodes, self.num_walks, self.walk_length, num_neighbors, self.seed,
                                     epoch, epoch_dev=epoch_dev, num_epochs=num_layers)
        return [NL.from_walk(ids[l], w[l], nv[l]) for l in range(num_layers)]

    # ---- N4: PPR push variant ---------------------------------------------------
    def _ppr_sources(self, nodes):
        if isinstance(nodes, torch.Tensor):
            nodes = nodes.tolist()
        nodes = [int(v) for v in nodes]
        # the reference sizes its vectors max(edge_index.max() + 1, max(nodes) + 1) and then indexes
        # adj_list[source]: a source beyond the adjacency list raises IndexError (:176)
        if nodes and (min(nodes) < 0 or max(nodes) >= self.csr.num_nodes):
            raise IndexError("list index out of range")
        return nodes

    def compute_ppr_dense(self, nodes, alpha=0.15, num_iterations=10):
        """compute_ppr_matrix as a device tensor: float64 [len(nodes), num_nodes] (pb200_ppr_push)."""
        return K.ppr_push(self.csr, self._ppr_sources(nodes), alpha, num_iterations)

    def compute_ppr_matrix(self, nodes, alpha=0.15, num_iterations=10):
        """reference :144-195 -- {(source, target): score} for every positive score, sources in the given
        order, targets ascending (the reference's insertion order)."""
        nodes = self._ppr_sources(nodes)
        dense = K.ppr_push(self.csr, nodes, alpha, num_iterations).cpu().numpy()
        ppr_matrix = {}
        for row, source in zip(dense, nodes):
            for target in row.nonzero()[0].tolist():
                if row[target] > 0:
                    ppr_matrix[(source, target)] = float(row[target])
        return ppr_matrix

    def precompute_top_neighbors(self, nodes, num_neighbors=10):
        """reference :197-229 -- {source: (neighbors, weights)}: the num_neighbors largest PPR scores (ties by
        smaller target id = the reference's stable reverse sort), weights normalised by their python-float sum."""
        if isinstance(nodes, torch.Tensor):
            nodes = nodes.tolist()
        sources = self._ppr_sources(nodes)
        dense = K.ppr_push(self.csr, sources, 0.15, 10)
        ids, vals = K.topk_rows_f64(dense, num_neighbors)
        ids, vals = ids.cpu().tolist(), vals.cpu().tolist()
        top_neighbors = {}
        for source, row_i, row_v in zip(nodes, ids, vals):
            neighbors = [t for t in row_i if t >= 0]
            weights = row_v[:len(neighbors)]
            if weights:
                total_weight = sum(weights)
                weights = [w / total_weight for w in weights]
            top_neighbors[source] = (neighbors, weights)
        return top_neighbors
