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
    """graph = GraphedEmbeddings(model, x, sampler, T[, num_items, group, out]); emb = graph.replay().

    x: features (this rank's rows when sharded) -- device-resident, or a PINNED host tensor: the upload
    then becomes a node of the graph on a forked stream, running under the walk kernels (which do not
    need the features), and every replay re-reads the host buffer.  out: optional pinned host tensor
    that receives the embeddings (a download node at the end of the graph).  `replay()` returns the
    same tensor every time (overwritten in place, stream ordered)."""

    def __init__(self, model, x, sampler, num_neighbors=10, num_items=None, group=None, warmup=2, out=None):
        dev = model._device()
        if dev.type != "cuda":
            raise N.NativeError("GraphedEmbeddings needs a CUDA device")
        rank, ws = SH.world(group)
        self.ws = ws
        self.layers = model.num_layers
        self.replays = 0
        host_in = isinstance(x, torch.Tensor) and not x.is_cuda
        if host_in and not x.is_pinned():
            raise N.NativeError("GraphedEmbeddings: a host feature tensor must be pinned (the upload is a graph node)")
        if out is not None and not out.is_cuda and not out.is_pinned():
            raise N.NativeError("GraphedEmbeddings: a host output tensor must be pinned")
        # the graph bakes in addresses: keep every captured input alive for the graph's lifetime
        self.model, self.x_src, self.sampler, self.out_host = model, x, sampler, out
        self.x_dev = torch.empty(x.shape, dtype=torch.float32, device=dev) if host_in else x
        self.base = int(sampler.epoch)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        n_items = x.size(0) if num_items is None else int(num_items)
        sl = SH.local_slice(n_items, rank, ws) if ws > 1 else slice(0, n_items, 1)
        nodes = self.nodes = torch.arange(sl.start, sl.stop, sl.step or 1, dtype=torch.int32, device=dev)
        side_up = torch.cuda.Stream(dev) if host_in else None

        def sample():
            return sampler.sample_layers(nodes, num_neighbors, self.layers, epoch=self.base, epoch_dev=self.epoch_dev)

        def step():
            cur = torch.cuda.current_stream(dev)
            if host_in:                                  # fork: upload under the walks
                side_up.wait_stream(cur)
                with torch.cuda.stream(side_up):
                    self.x_dev.copy_(x, non_blocking=True)
            if ws > 1:
                emb = SH.get_embeddings_sharded(model, self.x_dev, sampler, n_items, num_neighbors, group,
                                                epoch_base=self.base, epoch_dev=self.epoch_dev,
                                                check_barriers=False,
                                                before_forward=(lambda: cur.wait_stream(side_up)) if host_in else None)
            else:
                batches = sample()
                if host_in:
                    cur.wait_stream(side_up)
                emb = model.forward(self.x_dev, None, batches, None)
            if out is not None:
                out.copy_(emb, non_blocking=True)
            return emb

        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # allocator, weight caches, peer buffers: all warm
            for _ in range(max(warmup, 1)):
                step()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        if ws > 1:
            SH.check_peer_barriers()
        self.epoch_dev.zero_()
        self.graph = torch.cuda.CUDAGraph()
        l0 = N.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = step()
            K.u32_add(self.epoch_dev, self.layers)
        self.launches_per_replay = N.launch_count() - l0

    def replay(self, check=True):
        """One step.  check (sharded graphs): read the peer barriers' time-out flag afterwards (a stream
        sync) and raise if a peer never arrived; a pipelined loop passes False and calls
        sharding.check_peer_barriers(force=True) before it trusts the results."""
        self.graph.replay()
        self.replays += 1
        # eager sampling calls made after replays continue the epoch sequence instead of reusing it
        self.sampler.epoch = max(self.sampler.epoch, self.base + self.replays * self.layers)
        if check and self.ws > 1:
            SH.check_peer_barriers(force=True)
        return self.out if self.out_host is None else self.out_host
