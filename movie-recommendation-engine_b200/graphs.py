"""CUDA-graph capture of one whole embedding step (PinSage.get_embeddings: L x sampling + forward).

Small catalogues and multi-GPU shards make the step launch bound (~15 kernels of a few tens of
microseconds each plus, across GPUs, one barrier per layer): the step is captured once and
replayed.  Replays still draw fresh walks: the sampling epoch is `base + *epoch_dev` inside the
walk kernel (pb200_walk_topt_indexed_ex) and the last node of the graph advances the device
counter by the number of layers -- replay k equals the eager call number k bit for bit.
"""
from __future__ import annotations

import torch

from . import _native as N
from . import kernels as K
from . import neighbor_lists as NL
from . import sharding as SH


class GraphedEmbeddings:
    """graph = GraphedEmbeddings(model, x_dev, sampler, T[, num_items, group]); out = graph.replay().

    x_dev: device-resident features (this rank's rows when sharded).  `replay()` returns the
    same output tensor every time (overwritten in place, stream ordered)."""

    def __init__(self, model, x_dev, sampler, num_neighbors=10, num_items=None, group=None, warmup=2):
        dev = model._device()
        if dev.type != "cuda":
            raise N.NativeError("GraphedEmbeddings needs a CUDA device")
        rank, ws = SH.world(group)
        self.layers = model.num_layers
        # the graph bakes in addresses: keep every captured input alive for the graph's lifetime
        self.model, self.x_dev, self.sampler = model, x_dev, sampler
        self.base = int(sampler.epoch)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        n_items = x_dev.size(0) if num_items is None else int(num_items)
        lo, hi = SH.shard_range(n_items, rank, ws)
        nodes = self.nodes = torch.arange(lo, hi, dtype=torch.int32, device=dev)

        def step():
            if ws > 1:
                return SH.get_embeddings_sharded(model, x_dev, sampler, n_items, num_neighbors, group,
                                                 epoch_base=self.base, epoch_dev=self.epoch_dev)
            batches = []
            for layer in range(self.layers):
                ids, _c, w, nv = sampler._sample(nodes, num_neighbors, epoch=self.base + layer, check=False,
                                                 epoch_dev=self.epoch_dev)
                batches.append(NL.from_walk(ids, w, nv))
            return model.forward(x_dev, None, batches, None)

        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # allocator, weight caches, peer buffers: all warm
            for _ in range(max(warmup, 1)):
                step()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.epoch_dev.zero_()
        self.graph = torch.cuda.CUDAGraph()
        l0 = N.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = step()
            K.u32_add(self.epoch_dev, self.layers)
        self.launches_per_replay = N.launch_count() - l0

    def replay(self):
        self.graph.replay()
        return self.out
