"""Drop-in for the reference's ``utils/nearest_neighbors.py`` without faiss, B200-native.

LSHIndex (reference :7-68) wraps ``faiss.IndexLSH(dim, num_bits, num_tables)``; faiss's third
positional parameter is ``rotate_data``, so what the reference computes is ONE num_bits-bit
sign-random-rotation code per vector and an exhaustive Hamming top-k.  ``mode="exhaustive"``
(default) reproduces that; ``mode="tables"`` is the bucketed variant the constructor's
``num_tables`` suggests (num_tables keys of num_bits/num_tables bits, bucket probe, exact
dedup, popcount or dot-product re-rank).
WeakANDIndex (:70-139) = IndexFlatL2 quantizer + IndexIVFFlat(nlist) with
nprobe = min(nlist, 20); ``candidates_factor`` is stored and unused, as in the reference.
faiss is not vendored by the reference and absent offline: rotation matrix and k-means follow
faiss's published algorithms with this package's own seeded generators (parity is conditional
on shared parameters: pass ``projection=`` / ``centroids=`` to pin them).
Outputs follow faiss: (float32 [n,k] distances, int64 [n,k] labels) as numpy on the host,
ascending distance, label -1 when fewer than k results.
"""
from __future__ import annotations

import time
import types

import numpy as np
import torch

from .. import _native as N
from .. import kernels as K


def _to_f32_tensor(a):
    """reference :35-39 / :59-63: tensors -> cpu numpy -> float32 (here: straight to float32)."""
    if isinstance(a, torch.Tensor):
        return a.detach().to(torch.float32)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32))


def _np_results(dist, ids):
    return dist.cpu().numpy().astype(np.float32), ids.cpu().numpy().astype(np.int64)


def random_rotation(dim, num_bits, seed=5):
    """Gaussian + QR, as faiss RandomRotationMatrix(d_in=dim, d_out=num_bits).init(seed):
    orthonormal rows if num_bits <= dim, else the first dim columns of a num_bits x num_bits
    orthonormal matrix (tight frame).  Returns float32 [num_bits, dim]."""
    g = torch.Generator().manual_seed(seed)
    if num_bits <= dim:
        q, _ = torch.linalg.qr(torch.randn(dim, num_bits, generator=g, dtype=torch.float64))
        return q.t().contiguous().to(torch.float32)
    q, _ = torch.linalg.qr(torch.randn(num_bits, num_bits, generator=g, dtype=torch.float64))
    return q[:, :dim].contiguous().to(torch.float32)


class LSHIndex:
    def __init__(self, dim, num_bits=256, num_tables=16, *, mode="exhaustive", rerank="hamming",
                 projection=None, device=None):
        if mode not in ("exhaustive", "tables"):
            raise ValueError("mode must be 'exhaustive' or 'tables'")
        if num_bits % 32:
            raise ValueError("num_bits must be a multiple of 32")
        self.dim, self.num_bits, self.num_tables = dim, num_bits, num_tables
        self.mode, self.rerank = mode, rerank
        self.device = N.device_of(device=device)
        proj = random_rotation(dim, num_bits) if projection is None else _to_f32_tensor(projection)
        if tuple(proj.shape) != (num_bits, dim):
            raise ValueError(f"projection must have shape ({num_bits}, {dim})")
        self.projection = proj.to(self.device).contiguous()
        self.codes = torch.empty((0, num_bits // 8), dtype=torch.uint8, device=self.device)
        self.vectors = None
        self._tables = None
        self.index = types.SimpleNamespace(ntotal=0, d=dim, nbits=num_bits, is_trained=True)

    def build(self, embeddings):
        """reference :28-45 (train is a no-op for IndexLSH; add appends)."""
        x = _to_f32_tensor(embeddings).to(self.device).contiguous()
        if x.dim() != 2 or x.size(1) != self.dim:
            raise RuntimeError(f"embeddings must be [n, {self.dim}]")
        self.codes = torch.cat([self.codes, K.lsh_encode(x, self.projection)])
        if self.mode == "tables":
            if self.rerank == "dot":
                self.vectors = x if self.vectors is None else torch.cat([self.vectors, x])
            self._tables = K.lsh_build_tables(self.codes, self.num_tables)
        self.index.ntotal = self.codes.size(0)
        print(f"Built LSH index with {x.size(0)} embeddings")

    def search(self, queries, k=10):
        """reference :47-68 -> (Hamming distances as float32 [n,k], int64 ids [n,k])."""
        q = _to_f32_tensor(queries).to(self.device).contiguous()
        if q.dim() == 1:
            q = q[None]
        if q.size(1) != self.dim:
            raise RuntimeError(f"queries must be [n, {self.dim}]")
        cq = K.lsh_encode(q, self.projection)
        if self.mode == "exhaustive":
            dist, ids = K.hamming_topk(cq, self.codes, k)
        else:
            dot = self.rerank == "dot"
            dist, ids, self.last_num_candidates = K.lsh_search_tables(
                cq, self.codes, self.num_tables, *self._tables, k,
                queries=q if dot else None, vectors=self.vectors if dot else None)
        return _np_results(dist, ids)


def train_kmeans(x, nlist, niter=20, seed=1234, max_points_per_centroid=256, split_empty=True):
    """Lloyd iterations with faiss Clustering's defaults (subsample to 256 points/centroid,
    centroids initialised from a random subset, niter=20); assignment via pb200_topk (L2, k=1),
    update via the deterministic per-list mean kernel."""
    n = x.size(0)
    g = torch.Generator().manual_seed(seed)
    if n > nlist * max_points_per_centroid:
        x = x[torch.randperm(n, generator=g)[:nlist * max_points_per_centroid].to(x.device)]
        n = x.size(0)
    if n < nlist:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as "
                           f"number of clusters ({nlist})")
    cent = x[torch.randperm(n, generator=g)[:nlist].to(x.device)].contiguous().clone()
    for _ in range(niter):
        _, a = K.topk(x, cent, 1, N.METRIC_L2)
        offsets, _ids, vecs = K.ivf_build(x, a.view(-1).contiguous(), nlist)
        K.ivf_centroid_update(vecs, offsets, cent)
        if split_empty:
            _split_empty_clusters(cent, offsets, g)
    return cent


def _split_empty_clusters(cent, offsets, gen, eps=1.0 / 1024):
    """faiss Clustering::split_clusters: an empty cluster takes a copy of a donor cluster's centroid (donor drawn
    with probability ~ size - 1), the two copies are nudged apart by (1 +- eps) on alternating coordinates and
    share the donor's points from the next assignment on.  Same rule, this package's seeded generator instead of
    faiss's (unpinned either way: faiss is absent).  Without it, near-duplicate (collapsed) embeddings leave
    most lists dead for good.  nlist x d work: host-side bookkeeping of the build."""
    sizes = (offsets[1:] - offsets[:-1]).to(torch.float64).cpu()
    empty = (sizes == 0).nonzero().view(-1).tolist()
    if not empty:
        return 0
    d = cent.size(1)
    sign = torch.where(torch.arange(d, device=cent.device) % 2 == 0, 1.0 + eps, 1.0 - eps).to(cent.dtype)
    for ci in empty:
        p = (sizes - 1).clamp_min(0)
        if float(p.sum()) <= 0:
            break
        cj = int(torch.multinomial(p, 1, generator=gen))
        cent[ci] = cent[cj] * sign
        cent[cj] = cent[cj] * (2.0 - sign)
        sizes[ci] = sizes[cj] / 2
        sizes[cj] -= sizes[ci]
    return len(empty)


class WeakANDIndex:
    def __init__(self, dim, num_partitions=100, candidates_factor=10, *, centroids=None,
                 device=None):
        self.dim, self.num_partitions, self.candidates_factor = dim, num_partitions, candidates_factor
        self.device = N.device_of(device=device)
        self.centroids = None if centroids is None else \
            _to_f32_tensor(centroids).to(self.device).contiguous()
        self.quantizer = types.SimpleNamespace(ntotal=0 if centroids is None else num_partitions, d=dim)
        self.index = types.SimpleNamespace(ntotal=0, d=dim, nlist=num_partitions, nprobe=1,
                                           is_trained=centroids is not None)
        self._x = torch.empty((0, dim), dtype=torch.float32, device=self.device)
        self._lists = None
        self._tc_layout = None
        self.precision = "auto"      # "fp32": always the list-scan kernel; "auto": tensor cores for >= 256 queries

    def build(self, embeddings):
        """reference :94-113: train (k-means) + add."""
        x = _to_f32_tensor(embeddings).to(self.device).contiguous()
        if x.dim() != 2 or x.size(1) != self.dim:
            raise RuntimeError(f"embeddings must be [n, {self.dim}]")
        if self.centroids is None:
            self.centroids = train_kmeans(x, self.num_partitions)
            self.quantizer.ntotal = self.num_partitions
            self.index.is_trained = True
        self._x = torch.cat([self._x, x])
        _, a = K.topk(self._x, self.centroids, 1, N.METRIC_L2)
        self.assign = a.view(-1).contiguous()
        self._lists = K.ivf_build(self._x, self.assign, self.num_partitions)
        # padded list-ordered TF32 copy for the tensor-core search path (None: shape not covered)
        self._tc_layout = K.ivf_tc_layout(*self._lists, self.num_partitions)
        self.index.ntotal = self._x.size(0)
        print(f"Built Weak AND index with {x.size(0)} embeddings")

    def search(self, queries, k=10):
        """reference :115-139."""
        if self._lists is None:
            raise RuntimeError("index is not trained: call build() first")
        q = _to_f32_tensor(queries).to(self.device).contiguous()
        if q.dim() == 1:
            q = q[None]
        self.index.nprobe = min(self.num_partitions, 20)                       # :134
        _, probes = K.topk(q, self.centroids, self.index.nprobe, N.METRIC_L2)
        lay = getattr(self, "_tc_layout", None)
        if (self.precision != "fp32" and lay is not None and q.size(0) >= K.TOPK_TC_MIN_QUERIES
                and K.ivf_search_tc_supported(q.size(0), lay[0].size(0), self.dim, k, self.num_partitions)):
            st = {}
            dist, ids = K.ivf_search_tc(q, probes, *self._lists, lay, self.num_partitions, k, stats=st)
            out = _np_results(dist, ids)                       # (synchronises: the counter below is ready)
            # near-tied data (e.g. collapsed embeddings) defeats the TF32 certificate and
            # every query is re-run by the list-scan kernel anyway: stop paying for the tensor-core pass
            if self.precision == "auto" and int(st["list_scan_reruns"].item()) > q.size(0) // 4:
                self.precision = "fp32"
            return out
        dist, ids = K.ivf_search(q, probes, *self._lists, k)
        return _np_results(dist, ids)


class FlatL2Index:
    """What the reference's 'exact' method uses: faiss.IndexFlatL2 (:176-181)."""

    def __init__(self, dim, device=None):
        self.dim = dim
        self.device = N.device_of(device=device)
        self._x = torch.empty((0, dim), dtype=torch.float32, device=self.device)
        self.ntotal = 0
        self.precision = "auto"      # "auto" | "fp32" | "tf32" (kernels.topk); see search()

    def add(self, embeddings):
        self._x = torch.cat([self._x, _to_f32_tensor(embeddings).to(self.device)]).contiguous()
        self.ntotal = self._x.size(0)

    def search(self, queries, k):
        q = _to_f32_tensor(queries).to(self.device).contiguous()
        st = {}
        out = _np_results(*K.topk(q, self._x, k, N.METRIC_L2, precision=self.precision, stats=st))
        # results are the fp32 kernel's either way; if the certificate sends most queries back to it
        # (near-tied data), skip the tensor-core pass from now on
        if (self.precision == "auto" and st.get("path") == "tf32"
                and int(st["fp32_reruns"].item()) > q.size(0) // 4):
            self.precision = "fp32"
        return out


def benchmark_search_methods(embeddings, queries, k=10, methods=None):
    """reference :141-254 -- same result dict; search time includes the device sync and the
    D2H copy of the results (the reference times a synchronous CPU call)."""
    emb = _to_f32_tensor(embeddings)
    qs = _to_f32_tensor(queries)
    dim = emb.shape[1]
    if methods is None:
        methods = ['exact', 'lsh', 'ivf']
    results = {}
    for method in methods:
        print(f"Benchmarking {method} search...")
        if method == 'exact':
            index = FlatL2Index(dim)
            index.add(emb)
            searcher, size, label = index, lambda: index.ntotal, 'Exact (Brute Force)'
        elif method == 'lsh':
            lsh = LSHIndex(dim)
            lsh.build(emb)
            searcher, size, label = lsh, lambda: lsh.index.ntotal, 'Locality-Sensitive Hashing'
        elif method == 'ivf':
            ivf = WeakANDIndex(dim)
            ivf.build(emb)
            searcher, size, label = ivf, lambda: ivf.index.ntotal, 'Weak AND (IVF)'
        else:
            continue
        torch.cuda.synchronize()
        start_time = time.time()
        distances, indices = searcher.search(qs, k)
        search_time = time.time() - start_time
        results[method] = {'distances': distances, 'indices': indices, 'search_time': search_time,
                           'index_size': size(), 'method': label}
    print("\nBenchmark Results:")
    print("-----------------")
    for method, data in results.items():
        print(f"{data['method']}:")
        print(f"  Search time: {data['search_time']:.6f} seconds")
        print(f"  Index size: {data['index_size']} vectors")
    if 'exact' in results:                                                     # :237-252
        exact_indices = results['exact']['indices']
        for method, data in results.items():
            if method != 'exact':
                recall = 0
                for i in range(len(qs)):
                    recall += len(set(exact_indices[i]) & set(data['indices'][i])) / k
                recall /= len(qs)
                results[method]['recall'] = recall
                print(f"  {data['method']} recall@{k}: {recall:.4f}")
    return results
