"""Multi-GPU layout of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards and what is exchanged:
  * walks       start nodes are split in contiguous blocks; the CSR is replicated; no
                collective (Philox counters make a node's sample independent of its shard).
  * conv layers rows are split the same way; layer l+1 gathers rows of h^(l) from arbitrary
                nodes.  On GPUs of one box the pooling kernel reads those rows straight from
                their owner's memory over NVLink (CUDA IPC peer buffers, `PeerBuffers`; ~1.3
                remote rows per node instead of the whole matrix) and the only synchronisation
                is one flag barrier on peer memory per layer (pb200_peer_barrier: a kernel, so
                the whole step can be replayed as a CUDA graph; NCCL collectives inside the
                capture hung on this stack).  The all-gather
                of the row shards of h is kept for CPU/gloo runs and the exact-fp32 path.
  * search      queries are split, the index is replicated, results are all-gathered; or items
                are split (exact search on catalogues that do not fit one GPU), every rank
                returns a local top-k with global ids, and the all-gathered lists are merged
                by pb200_topk_merge under the (score, id) total order -- identical to the
                unsharded result bit for bit.
Works with the NCCL backend on GPUs and with gloo on CPU tensors (host-logic tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_size(n, world_size):
    return (n + world_size - 1) // world_size


def shard_range(n, rank, world_size):
    """Contiguous block [lo, hi) of rank; all blocks have shard_size rows except the tail."""
    s = shard_size(n, world_size)
    lo = min(rank * s, n)
    return lo, min(lo + s, n)


# How the rows of the embedding step (start nodes of the walks = rows of x, h and the embeddings) are dealt
# to ranks.  "cyclic": row i belongs to rank i % world -- the catalogues this path sees are popularity
# sorted (MovieLens ids, the synthetic generator), and the heavy rows cost the walk kernel ~20 % more
# (more distinct buckets per start row: more DRAM misses), so contiguous blocks make rank 0 the slowest
# rank of every step (r2 measurement at N = 2: 203 us vs 166 us).  "blocks": shard_range.
EMB_LAYOUT = "cyclic"


def local_slice(n, rank, world_size, layout=None):
    """The rows of an n-row matrix that rank owns, as a slice."""
    if (layout or EMB_LAYOUT) == "cyclic":
        return slice(rank, n, world_size)
    lo, hi = shard_range(n, rank, world_size)
    return slice(lo, hi)


def local_rows(t, group=None, layout=None):
    """This rank's rows of a full [n, ...] tensor."""
    rank, ws = world(group)
    return t[local_slice(t.size(0), rank, ws, layout)]


def all_gather_rows(local, n_total, group=None, layout="blocks"):
    """Row shards (shard_range layout, or cyclic) -> the full [n_total, ...] tensor on every rank."""
    rank, ws = world(group)
    if ws == 1:
        return local
    if layout == "cyclic":
        s = shard_size(n_total, ws)
        if local.size(0) != s:
            pad = torch.zeros((s - local.size(0),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            local = torch.cat([local, pad])
        out = torch.empty((ws, s) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out.view((ws * s,) + tuple(local.shape[1:])), local.contiguous(), group=group)
        # out[r, j] = row r + j * ws
        return out.transpose(0, 1).reshape((ws * s,) + tuple(local.shape[1:]))[:n_total].contiguous()
    s = shard_size(n_total, ws)
    if local.size(0) != s:                                  # pad the tail shard
        pad = torch.zeros((s - local.size(0),) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        local = torch.cat([local, pad])
    out = torch.empty((s * ws,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


def all_gather_cols(local, group=None):
    """[nq, c] per rank -> [nq, c * world] (per-shard candidate lists side by side)."""
    rank, ws = world(group)
    if ws == 1:
        return local
    nq, c = local.shape
    out = torch.empty((ws * nq, c), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out.view(ws, nq, c).permute(1, 0, 2).reshape(nq, ws * c).contiguous()


class _RawCuda:
    """A cudaMalloc'ed region viewed through __cuda_array_interface__ (float32, C order)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerBuffers:
    """`count` row-shard buffers [shard_rows, width] fp32 on every rank, allocated by
    pb200_peer_alloc and opened on all peers with CUDA IPC (one process per GPU, one box).
    `ptr_array(b)` is the host array of world pointers pb200_pool_sharded takes for buffer b;
    `local(b)` is this rank's shard as a tensor.  `barrier()` is stream ordered."""

    def __init__(self, count, shard_rows, width, dev, group=None):
        import ctypes
        from . import _native as N
        self.group, self.dev = group, dev
        self.rank, self.ws = world(group)
        self.shard_rows, self.width, self.count = shard_rows, width, count
        lib = N.lib()
        nbytes = max(shard_rows * width * 4, 16)
        self._own, self._opened, self._arrays, self._keep, self._local = [], [], [], [], []
        self._flags_local = None
        self.ok, self.error = True, ""
        # Every rank runs the same collectives whatever fails locally (no CUDA IPC between these
        # devices, out of memory, ...): the ranks then agree on ok / not ok, and the caller falls
        # back to the all-gather exchange on ALL ranks or on none.
        handles = None
        try:
            import os
            if os.environ.get("PB200_TEST_PEER_FAIL_RANK") == str(self.rank):     # test hook (tools/multi_gpu_check.py)
                raise RuntimeError("simulated peer-buffer failure")
            handles = []
            for b in range(count + 1):                       # `count` row buffers + the barrier flag array
                p = ctypes.c_void_p()
                N.check(lib.pb200_peer_alloc(nbytes if b < count else 256, ctypes.byref(p)), "peer_alloc")
                self._own.append(p)
                h = (ctypes.c_uint8 * 64)()
                N.check(lib.pb200_peer_export(p, h), "peer_export")
                handles.append(bytes(h))
        except Exception as e:                               # noqa: BLE001 -- reported through self.error
            self.ok, self.error, handles = False, str(e), None
        gathered = [None] * self.ws
        dist.all_gather_object(gathered, handles, group=group)
        if self.ok and any(g is None for g in gathered):
            self.ok, self.error = False, "a peer could not allocate / export its buffers"
        if self.ok:
            try:
                table = []                                   # table[b][r]: buffer b of rank r, mapped here
                for b in range(count + 1):
                    ptrs = []
                    for r in range(self.ws):
                        if r == self.rank:
                            ptrs.append(self._own[b].value)
                        else:
                            q = ctypes.c_void_p()
                            hb = (ctypes.c_uint8 * 64).from_buffer_copy(gathered[r][b])
                            N.check(lib.pb200_peer_open(hb, ctypes.byref(q)), "peer_open")
                            self._opened.append(q)
                            ptrs.append(q.value)
                    table.append(ptrs)
                for b in range(count):
                    self._arrays.append((ctypes.c_void_p * self.ws)(*table[b]))
                    raw = _RawCuda(self._own[b].value, (shard_rows, width))
                    self._keep.append(raw)
                    self._local.append(torch.as_tensor(raw, device=dev))
                # barrier state: one flag array per rank in peer memory + this rank's sequence counter
                flags_raw = _RawCuda(self._own[count].value, (64,))
                self._keep.append(flags_raw)
                self._flags_local = torch.as_tensor(flags_raw, device=dev)
                self._flags_local.zero_()
                self._flag_ptrs = torch.tensor(table[count], dtype=torch.int64, device=dev)
                self._seq = torch.zeros(2, dtype=torch.int32, device=dev)      # [sequence counter, error flag]
            except Exception as e:                           # noqa: BLE001
                self.ok, self.error = False, str(e)
        torch.cuda.synchronize(dev)
        agree = torch.tensor([1 if self.ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)   # also: every rank's flags are zeroed and mapped
        if int(agree.item()) == 0:
            if self.ok:
                self.ok, self.error = False, "a peer could not map the exchange buffers"
            self._release()

    def _release(self):
        from . import _native as N
        for q in self._opened:
            N.lib().pb200_peer_close(q)
        for p in self._own:
            N.lib().pb200_peer_free(p)
        self._opened, self._own, self._local, self._keep, self._flags_local, self._arrays = [], [], [], [], None, []

    def ptr_array(self, b):
        return self._arrays[b]

    def local(self, b):
        return self._local[b]

    def barrier(self):
        """All ranks' work queued before this point is complete and visible before anything queued
        after it starts on any rank: one small kernel on peer-memory flags (pb200_peer_barrier), so
        a step containing it can be captured in a CUDA graph."""
        from . import _native as N
        N.check(N.lib().pb200_peer_barrier_ex(N.ptr(self._flag_ptrs), N.ptr(self._seq[0:1]), self.rank, self.ws,
                                              N.ptr(self._seq[1:2]), self.max_spins, N.stream_ptr(self.dev)),
                "peer_barrier")
        self._unchecked = True

    max_spins = 1 << 27            # ~5 s of 40 ns polls before a barrier gives up

    def timed_out(self):
        """Bit mask of the peers a barrier gave up waiting for since the last check (device sync)."""
        return int(self._seq[1].item()) & 0xFFFFFFFF

    def check(self):
        """Raises if any barrier since the last check timed out: what was pooled after it may have read
        rows a peer had not written yet.  Synchronises the stream; clears the flag."""
        from . import _native as N
        if not getattr(self, "_unchecked", False):
            return
        missing = self.timed_out()
        self._unchecked = False
        if missing:
            self._seq[1].zero_()
            peers = [r for r in range(self.ws) if missing >> r & 1]
            raise N.NativeError(f"rank {self.rank}: peer barrier timed out waiting for rank(s) {peers}; the "
                                "embeddings of this step are not valid")

    def close(self):
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        self._release()


_PEER_CACHE = {}


def peer_buffers(count, shard_rows, width, dev, group=None):
    """The cached PeerBuffers for this shape, or None when the ranks agreed that peer memory is not
    usable here (the caller then all-gathers; the decision is collective, so all ranks take the
    same path)."""
    key = (count, shard_rows, width, str(dev), id(group))
    if key not in _PEER_CACHE:
        pb = PeerBuffers(count, shard_rows, width, dev, group)
        if not pb.ok and pb.rank == 0:
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({pb.error}); falling back to all-gather of h")
        _PEER_CACHE[key] = pb
    pb = _PEER_CACHE[key]
    return pb if pb.ok else None


def check_peer_barriers(force=False):
    """Raises NativeError if any peer barrier issued since the last check timed out (see PeerBuffers.check).
    force: also read the flag when no barrier was issued from Python (CUDA-graph replays issue them)."""
    for pb in _PEER_CACHE.values():
        if pb.ok:
            if force:
                pb._unchecked = True
            pb.check()


def release_peer_buffers():
    for pb in _PEER_CACHE.values():
        if pb.ok:
            pb.close()
    _PEER_CACHE.clear()


def _use_peer_exchange(model, dev, ws):
    import os
    from . import _native as N
    if ws == 1 or dev.type != "cuda" or ws > 16:
        return False
    if os.environ.get("PB200_SHARD_EXCHANGE", "p2p") != "p2p":
        return False
    return dist.get_backend() == "nccl" and model.precision != N.PREC_FP32 and not model.fuse_pool


_WALK_STREAMS = {}


def _fork_walks(rows):
    env = os.environ.get("PB200_FORK_WALKS")
    if env is not None:
        return env != "0"
    return rows <= 74 * 128


def _walk_stream(dev):
    key = (dev.type, dev.index)
    if key not in _WALK_STREAMS:
        _WALK_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _WALK_STREAMS[key]


def get_embeddings_sharded(model, x_local, sampler, num_items, num_neighbors=10, group=None,
                           epoch_base=None, epoch_dev=None, check_barriers=True, before_forward=None,
                           layout=None):
    """PinSage.get_embeddings with rows split across ranks.  x_local: this rank's rows of the
    feature matrix (``local_rows(x)``: EMB_LAYOUT, cyclic by default).  Returns this rank's rows of the
    embeddings (same layout; ``all_gather_rows(emb, n, layout=EMB_LAYOUT)`` reassembles them).
    epoch_base / epoch_dev: fixed host epoch + device-side counter (CUDA-graph capture, see
    graphs.GraphedEmbeddings); by default the sampler's own epoch counter advances.
    check_barriers: read the peer barriers' time-out flag after the step (one stream sync) and raise
    if a peer never arrived; pass False inside a CUDA-graph capture or a pipelined loop and call
    sharding.check_peer_barriers() before using the results.
    before_forward: optional callable run after the walk kernels are queued and before the features are
    first read (e.g. a stream wait on an upload that runs under the walks)."""
    from . import _native as N
    from . import kernels as K
    from . import neighbor_lists as NL
    rank, ws = world(group)
    layout = layout or EMB_LAYOUT
    mine_rows = local_slice(num_items, rank, ws, layout)
    dev = model._device()
    nodes = torch.arange(mine_rows.start, mine_rows.stop, mine_rows.step or 1, dtype=torch.int32, device=dev)

    def sample():                                           # same epochs on every rank; one launch for all layers
        return sampler.sample_layers(nodes, num_neighbors, model.num_layers, epoch=epoch_base, epoch_dev=epoch_dev)
    P = lambda lin: (lin.weight, lin.bias)     # Parameter objects: identity keys the TF32 weight cache
    RND = 0 if model.precision == N.PREC_FP32 else N.EPI_ROUND_TF32     # see PinSage.forward
    PRE = 0 if model.precision == N.PREC_FP32 else N.IN_A1_TF32
    pb = None
    if _use_peer_exchange(model, dev, ws):
        srows = shard_size(num_items, ws)
        pb = peer_buffers(model.num_layers, srows, model.input_proj.out_features, dev, group)
    rows = nodes.numel()
    # Order of the two independent prologue pieces.  Device-resident features: input projection FIRST, then
    # the walks -- by the time a rank reaches the first barrier every peer's h^(0) has long been written, so
    # the walk kernel absorbs the skew between ranks instead of the barrier.  Host features being uploaded
    # on a forked stream: walks first, so the upload runs under them.
    # Peer exchange, small shards: the walks run on a FORKED stream beside the input projection and the first
    # barrier -- a 61-tile GEMM (C2 on 8 GPUs) leaves most SMs idle and the barrier is a wait: 172.7 vs 183.5 us
    # per step at 7,803 rows per rank (profiles/r2_sharded_per_op_2gpus_7803rows_forked_walks.json).  A shard
    # whose GEMM fills the GPU loses from sharing it (15,606 rows per rank: 254 vs 244 us; 31,212: 0.400 vs 0.367 ms), so the fork is
    # taken only below half an SM-count of 128-row tiles.  PB200_FORK_WALKS=0 / 1 forces it off / on.
    batches = None
    forked = None
    if before_forward is not None:
        batches = sample()
        before_forward()
    elif pb is not None and dev.type == "cuda" and _fork_walks(rows):
        main = torch.cuda.current_stream(dev)
        forked = _walk_stream(dev)
        fork_point = main.record_event()
    xd = N.dev_tensor(x_local, torch.float32, dev)
    h_loc = K.gather_dense(xd, *P(model.input_proj), flags=N.EPI_RELU | RND, precision=model.precision,
                           out=pb.local(0)[:rows] if pb is not None else None)
    if forked is not None:                                      # queued after the GEMM so that its CTAs are placed first
        forked.wait_event(fork_point)
        with torch.cuda.stream(forked):
            batches = sample()
    if batches is None:
        batches = sample()
    if pb is not None:
        # neighbour rows are read from their owners' memory; no all-gather.  Every layer's output is
        # written by its GEMM straight into this rank's peer-visible shard of the next layer's input.
        for i in range(model.num_layers):
            pb.barrier()                                        # every rank's h^(i) is in place
            if forked is not None and i == 0:                   # join: the lists are needed from here on
                main.wait_stream(forked)
                if not torch.cuda.is_current_stream_capturing():   # (a capture's private pool is not shared)
                    for b in batches:
                        for t in b.as_args():
                            if t is not None:
                                t.record_stream(main)           # allocated on the forked stream, read on this one
            wf, bf = model._folded_layer(i)
            ids, wts, ll, wl = batches[i].as_args()
            h_neigh = K.pool_sharded(pb.ptr_array(i), ws, srows, num_items, h_loc.size(1), ids, wts, ll, wl,
                                     N.POOL_PINSAGE | N.POOL_ROUND_TF32, dev,
                                     layout=N.SHARD_CYCLIC if layout == "cyclic" else N.SHARD_BLOCKS)
            h_loc = K.gather_dense(h_loc, wf, bf, a2=h_neigh,
                                   flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE | N.IN_A2_TF32,
                                   precision=model.precision,
                                   out=pb.local(i + 1)[:rows] if i + 1 < model.num_layers else None)
        # Buffer i is rewritten by the next call only after the barrier of layer i+1 (which every
        # rank reaches after its pooling of layer i); a single layer has no such barrier.
        if model.num_layers == 1:
            pb.barrier()
        emb = K.gather_dense(h_loc, *P(model.output_proj), flags=N.EPI_L2NORM | PRE,
                             precision=model.precision)
        if check_barriers:
            pb.check()
        return emb
    for i in range(model.num_layers):
        h_full = all_gather_rows(h_loc, num_items, group, layout)   # the one exchange per layer
        wf, bf = model._folded_layer(i)
        if model.precision != N.PREC_FP32 and not model.fuse_pool:      # see PinSage.forward
            h_neigh = K.pool(h_full, *batches[i].as_args(), N.POOL_PINSAGE | N.POOL_ROUND_TF32)
            h_loc = K.gather_dense(h_full[mine_rows].contiguous(), wf, bf, a2=h_neigh,
                                   flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE | N.IN_A2_TF32,
                                   precision=model.precision)
            continue
        h_loc = K.gather_dense(h_full[mine_rows].contiguous(), wf, bf, pool_x=h_full, lists=batches[i].as_args(),
                               pool_mode=N.POOL_PINSAGE, flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE,
                               precision=model.precision)
    return K.gather_dense(h_loc, *P(model.output_proj), flags=N.EPI_L2NORM | PRE,
                          precision=model.precision)


def exact_search_item_sharded(queries, items_local, item_offset, k, metric, exclude_ids=None,
                              group=None):
    """Every rank scores all queries against its item shard; lists are all-gathered and merged."""
    from . import _native as N
    from . import kernels as K
    s, i = K.topk(queries, items_local, k, metric, exclude_ids=exclude_ids, id_offset=item_offset)
    rank, ws = world(group)
    if ws == 1:
        return s, i
    return K.topk_merge(all_gather_cols(s, group), all_gather_cols(i, group), k,
                        largest=metric == N.METRIC_IP)


def search_query_sharded(search_fn, queries, group=None):
    """search_fn(q_local) -> (scores [n,k], ids [n,k]) on this rank's block of queries; results
    of all ranks are all-gathered in query order."""
    rank, ws = world(group)
    n = queries.size(0)
    lo, hi = shard_range(n, rank, ws)
    s, i = search_fn(queries[lo:hi])
    return all_gather_rows(s, n, group), all_gather_rows(i, n, group)
