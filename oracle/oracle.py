"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain numpy restatement of the reference's PinSage inference + retrieval hot
path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the
product package must never do so
(tests/test_abi_and_host.py::test_product_never_imports_the_oracle enforces it).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Pinning status:

  S0-S3, P1-P5, G1-G4, E1 : PINNED by tests/golden/*.npz, which were produced by
      importing the unmodified reference in the build container
      (tests/golden/make_golden.py, make_golden_r2.py: incl. negative ids, the shipped
      checkpoint, the unpatched sampler's visit distribution, and the 8(f) rows N1-N4).
  E2, L1, L2, I1, I2, B1  : PARITY UNPINNED.  The arithmetic lives in
      faiss-cpu==1.7.4 (requirements.txt:19), which is not vendored, not
      installed and not fetchable offline.  These functions restate faiss's
      published algorithms and parity with the CUDA path
      is conditional on shared parameters (projection matrix, centroids).
"""
from __future__ import annotations

import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------
# Philox4x32-10 and the 53-bit uniform recipe (shared spec with the kernel)
# --------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10.  ctr: uint32[...,4], key: uint32[...,2]."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    key = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = key[..., 0].copy(), key[..., 1].copy()
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _MASK32
        n1 = p1 & _MASK32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _MASK32
        n3 = p0 & _MASK32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(_W0)) & _MASK32
        k1 = (k1 + np.uint64(_W1)) & _MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def walk_uniform53(seed, epoch, start, walk, step):
    """53-bit numerators k (u = k / 2**53) for broadcastable (start, walk, step).

    Bit recipe is numpy's legacy random_sample(): (a >> 5) * 2**26 + (b >> 6)."""
    start, walk, step = np.broadcast_arrays(
        np.asarray(start, dtype=np.uint64), np.asarray(walk, dtype=np.uint64),
        np.asarray(step, dtype=np.uint64))
    ctr = np.stack([start, walk, step >> np.uint64(1),
                    np.full(start.shape, epoch, dtype=np.uint64)], axis=-1)
    key = np.empty(start.shape + (2,), dtype=np.uint64)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    r = philox4x32_10(ctr, key).astype(np.uint64)
    odd = (step & np.uint64(1)).astype(bool)
    a = np.where(odd, r[..., 2], r[..., 0])
    b = np.where(odd, r[..., 3], r[..., 1])
    return ((a >> np.uint64(5)) << np.uint64(26)) | (b >> np.uint64(6))


# --------------------------------------------------------------------------
# S0: adjacency in edge order == stable CSR   (utils/random_walk.py:33-50)
# --------------------------------------------------------------------------
def csr_build(edge_index, edge_weights=None, num_nodes=None, quant_shift=1):
    """Returns (row_ptr int64[N+1], col int32[E], cum).  quant_shift >= 0: cum is
    uint32 row-local inclusive prefix of w * 2**quant_shift (must be integral);
    quant_shift < 0: cum is float64 sequential row-local prefix."""
    ei = np.asarray(edge_index, dtype=np.int64)
    E = ei.shape[1]
    N = int(ei.max()) + 1 if num_nodes is None else int(num_nodes)  # random_walk.py:36
    w = np.ones(E, dtype=np.float32) if edge_weights is None else \
        np.asarray(edge_weights, dtype=np.float32)                   # :45-48
    order = np.argsort(ei[0], kind="stable")                         # append order, :50
    deg = np.bincount(ei[0], minlength=N).astype(np.int64)
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    col = ei[1][order].astype(np.int32)
    ws = w[order]
    row_of = np.repeat(np.arange(N, dtype=np.int64), deg)
    if quant_shift >= 0:
        q = ws.astype(np.float64) * float(1 << quant_shift)
        if np.any(q < 0) or np.any(q != np.floor(q)):
            raise ValueError("weights are not non-negative multiples of the quantum")
        g = np.cumsum(q.astype(np.uint64))
        base = np.concatenate([[np.uint64(0)], g])[row_ptr[:-1]]
        cum = g - base[row_of] if E else g
        if E and cum.max() > 0xFFFFFFFF:
            raise OverflowError("row total overflows uint32 quanta")
        return row_ptr, col, cum.astype(np.uint32)
    cum = np.empty(E, dtype=np.float64)
    for v in range(N):  # sequential float64 prefix per row
        a, b = row_ptr[v], row_ptr[v + 1]
        if b > a:
            cum[a:b] = np.cumsum(ws[a:b].astype(np.float64))
    return row_ptr, col, cum


# --------------------------------------------------------------------------
# S1-S3: walks, visit counts, top-T              (utils/random_walk.py:52-142)
# --------------------------------------------------------------------------
def _pick(cum, r0, r1, k53):
    """First edge i in [r0, r1) with cum_i > u * total (the shared walk rule)."""
    if cum.dtype == np.uint32:
        total = int(cum[r1 - 1])
        t = (int(k53) * total) >> 53
        return r0 + int(np.searchsorted(cum[r0:r1], np.uint32(t), side="right"))
    total = float(cum[r1 - 1])
    x = (float(int(k53)) * (1.0 / 9007199254740992.0)) * total
    i = int(np.searchsorted(cum[r0:r1], x, side="right"))
    return r0 + min(i, r1 - r0 - 1)


def walk_bucket_index(row_ptr, col, cum, slots=8):
    """numpy restatement of the direct-addressed sampling index the walk kernel reads
    (csrc/walk_bucket.cu): per row a shift s and ceil(S / 2^s) 32-byte buckets.
    slots=8 (PB200_LEAF_BUCKET): {8 x u8 min(cum - j 2^s, 2^s) (unused: 128), 8 x id byte 0, 8 x id byte 1,
    8 x id byte 2}; slots=6 (PB200_LEAF_BUCKET32): {6 x u8 rel, 2 x 128, 6 x little-endian u32 id}.
    Returns (meta uint32 [N, 4] = {first bucket, degree, S, s}, leaf uint8 [buckets, 32]) or None
    when a zero-weight edge makes the format unusable.  Pure loops: small graphs only."""
    assert slots in (6, 8)
    N = len(row_ptr) - 1
    meta = np.zeros((N, 4), np.uint32)
    blocks = []
    for v in range(N):
        a, b = int(row_ptr[v]), int(row_ptr[v + 1])
        c = cum[a:b].astype(np.int64)
        deg = b - a
        S = int(c[-1]) if deg else 0
        prev = np.concatenate([[0], c[:-1]]) if deg else c
        if deg and np.any(c == prev):
            return None
        s = 7
        for i in range(deg - slots):                   # edges i..i+slots must not share a bucket
            x = int(c[i] - 1) ^ int(c[i + slots - 1])
            s = min(s, x.bit_length() - 1)
        meta[v] = (len(blocks), deg, S, s)
        w = 1 << s
        for j in range(((S - 1) >> s) + 1 if S else 0):
            lo = j << s
            blk = np.zeros(32, np.uint8); blk[:8] = 128
            first = int(np.searchsorted(c, lo, side="right"))      # first edge with cum > lo
            q = 0
            e = first
            while e < deg and (e == first or prev[e] < lo + w):
                assert q < slots, "bucket overflow: the shift rule is wrong"
                blk[q] = min(int(c[e]) - lo, w)
                nid = int(col[a + e])
                if slots == 8:
                    blk[8 + q], blk[16 + q], blk[24 + q] = nid & 255, (nid >> 8) & 255, (nid >> 16) & 255
                else:
                    blk[8 + 4 * q: 12 + 4 * q] = [(nid >> (8 * k)) & 255 for k in range(4)]
                q += 1; e += 1
            blocks.append(blk)
    leaf = np.stack(blocks) if blocks else np.zeros((0, 32), np.uint8)
    return meta, leaf


def walk_bucket_pick(meta, leaf, v, k53, slots=8):
    """The bucket walk step: neighbour chosen at node v for the 53-bit uniform numerator k53, or -1."""
    first, deg, S, s = (int(x) for x in meta[v])
    if deg == 0:
        return -1
    t = (k53 * S) >> 53
    blk = leaf[first + (t >> s)]
    tr = t & ((1 << s) - 1)
    c = int(np.sum(blk[:slots] <= tr))
    assert c < slots
    if slots == 6:
        return int.from_bytes(bytes(blk[8 + 4 * c: 12 + 4 * c]), "little")
    return int(blk[8 + c]) | (int(blk[16 + c]) << 8) | (int(blk[24 + c]) << 16)


def walk_topt(row_ptr, col, cum, starts, W, L, T, seed, epoch=0, return_trace=False):
    """Pure-python loop restatement (small cases).  Returns ids int32[n,T] (-1 pad),
    counts int32[n,T], w64 float64[n,T], nvalid int32[n] (+ trace int32[n,W,L])."""
    starts = np.asarray(starts, dtype=np.int64)
    n = len(starts)
    ids = np.full((n, T), -1, dtype=np.int32)
    counts = np.zeros((n, T), dtype=np.int32)
    w64 = np.zeros((n, T), dtype=np.float64)
    nvalid = np.zeros(n, dtype=np.int32)
    trace = np.full((n, W, L), -1, dtype=np.int32)
    for s, start in enumerate(starts.tolist()):
        ks = walk_uniform53(seed, epoch, start, np.arange(W)[:, None], np.arange(L)[None, :])
        seen = {}  # node -> [count, first]; dict keeps first-visit order (:101-104)
        for w in range(W):
            cur = start
            for l in range(L):
                r0, r1 = int(row_ptr[cur]), int(row_ptr[cur + 1])
                if r1 == r0:          # dead end, :68-69
                    break
                cur = int(col[_pick(cum, r0, r1, ks[w, l])])
                trace[s, w, l] = cur
                if cur in seen:
                    seen[cur][0] += 1
                else:
                    seen[cur] = [1, w * L + l]
        # sorted(..., key=count, reverse=True) is stable => first visit wins (:107)
        top = sorted(seen.items(), key=lambda kv: kv[1][0], reverse=True)[:T]
        tot = sum(c for _, (c, _f) in top)
        for j, (node, (c, _f)) in enumerate(top):
            ids[s, j] = node
            counts[s, j] = c
            w64[s, j] = c / tot       # python int / int, :113-115
        nvalid[s] = len(top)
    if return_trace:
        return ids, counts, w64, nvalid, trace
    return ids, counts, w64, nvalid


def count_topt_from_trace(trace, T):
    """Counting stage alone, given walk traces int32[n,W,L] (-1 = not visited).
    Mirrors utils/random_walk.py:100-117."""
    n = trace.shape[0]
    ids = np.full((n, T), -1, dtype=np.int32)
    counts = np.zeros((n, T), dtype=np.int32)
    nvalid = np.zeros(n, dtype=np.int32)
    for s in range(n):
        seen = {}
        for v in trace[s].reshape(-1).tolist():
            if v >= 0:
                seen[v] = seen.get(v, 0) + 1
        top = sorted(seen.items(), key=lambda kv: kv[1], reverse=True)[:T]
        for j, (node, c) in enumerate(top):
            ids[s, j], counts[s, j] = node, c
        nvalid[s] = len(top)
    return ids, counts, nvalid


# ---- C restatement (oracle/walk_oracle.c) loader: fast path for big cases ----
_C = None
C_BUILD = None


def c_oracle():
    global _C
    if _C is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle` or __graft_entry__.build()")
        # CPU baseline: -O3 -march=native, compiled on the machine that runs it (a .so built with
        # -march=native elsewhere may not run here); the portable build is the fallback
        global C_BUILD
        C_BUILD = "-O3 (portable build)"
        try:
            import subprocess
            import tempfile
            out = os.path.join(tempfile.gettempdir(), f"liboracle_native_{os.getuid()}.so")
            r = subprocess.run(["gcc", "-O3", "-march=native", "-fPIC", "-pthread", "-std=c11", "-shared", "-o", out,
                                os.path.join(_HERE, "walk_oracle.c")], capture_output=True, timeout=120)
            if r.returncode == 0:
                path, C_BUILD = out, "-O3 -march=native (built on this host)"
        except Exception:        # noqa: BLE001 -- no compiler here: portable build
            pass
        lib = ctypes.CDLL(path)
        lib.orc_walk_uniform53.restype = ctypes.c_uint64
        lib.orc_walk_uniform53.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_uint32, ctypes.c_uint32]
        lib.orc_max_threads.restype = ctypes.c_int
        _C = lib
    return _C


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def c_csr_build(edge_index, edge_weights, num_nodes, quant_shift=1):
    lib = c_oracle()
    ei = np.ascontiguousarray(edge_index, dtype=np.int64)
    E = ei.shape[1]
    w = None if edge_weights is None else np.ascontiguousarray(edge_weights, dtype=np.float32)
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    col = np.zeros(E, dtype=np.int32)
    kind = 0 if quant_shift >= 0 else 1
    cum = np.zeros(E, dtype=np.uint32 if kind == 0 else np.float64)
    rc = lib.orc_csr_build(_p(ei), _p(w), ctypes.c_int64(E), ctypes.c_int64(num_nodes),
                           ctypes.c_int(kind), ctypes.c_int(max(quant_shift, 0)),
                           _p(row_ptr), _p(col),
                           _p(cum) if kind == 0 else None, _p(cum) if kind == 1 else None)
    if rc != 0:
        raise ValueError(f"orc_csr_build failed rc={rc}")
    return row_ptr, col, cum


def c_walk_topt(row_ptr, col, cum, starts, W, L, T, seed, epoch=0, return_trace=False,
                num_threads=0):
    lib = c_oracle()
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    n = len(starts)
    ids = np.empty((n, T), dtype=np.int32)
    counts = np.empty((n, T), dtype=np.int32)
    w32 = np.empty((n, T), dtype=np.float32)
    w64 = np.empty((n, T), dtype=np.float64)
    nvalid = np.empty(n, dtype=np.int32)
    trace = np.empty((n, W, L), dtype=np.int32) if return_trace else None
    kind = 0 if cum.dtype == np.uint32 else 1
    rc = lib.orc_walk_topt(_p(row_ptr), _p(col), ctypes.c_int(kind),
                           _p(cum) if kind == 0 else None, _p(cum) if kind == 1 else None,
                           _p(starts), ctypes.c_int64(n), ctypes.c_int(W), ctypes.c_int(L),
                           ctypes.c_int(T), ctypes.c_uint64(seed), ctypes.c_uint32(epoch),
                           _p(ids), _p(counts), _p(w32), _p(w64), _p(nvalid), _p(trace),
                           ctypes.c_int(num_threads))
    if rc != 0:
        raise ValueError(f"orc_walk_topt failed rc={rc}")
    out = dict(ids=ids, counts=counts, w32=w32, w64=w64, nvalid=nvalid)
    if return_trace:
        out["trace"] = trace
    return out


# --------------------------------------------------------------------------
# P1-P5: pooling variants on ragged python lists
# --------------------------------------------------------------------------
def _zeros_row(x):
    return np.zeros(x.shape[1], dtype=np.float32)


def pool_pinsage(x, neighbors, weights):
    """model/pinsage.py:101-150 (ImportancePooling.forward)."""
    x = np.asarray(x, dtype=np.float32)
    max_idx = x.shape[0] - 1
    out = []
    for nb, wt in zip(neighbors, weights):
        if isinstance(nb, (int, np.integer)):          # :110-112
            nb, wt = [nb], [1.0]
        if len(nb) == 0:                               # :115-117
            out.append(_zeros_row(x)); continue
        vi, vw = [], []
        for j, idx in enumerate(nb):                   # :123-129
            if isinstance(idx, (int, np.integer)) and idx <= max_idx:
                vi.append(int(idx))
                vw.append(wt[j] if j < len(wt) else 1.0)
        if not vi:                                     # :132-134
            out.append(_zeros_row(x)); continue
        w32 = np.asarray(vw, dtype=np.float32)         # torch.tensor(list) -> fp32, :140
        s = w32.sum(dtype=np.float32)
        if s > 0:                                      # :142-143
            w32 = w32 / s
        out.append((x[vi] * w32[:, None]).sum(axis=0, dtype=np.float32))   # :146
    return np.stack(out) if out else np.zeros((0, x.shape[1]), np.float32)


def pool_layers(x, neighbors, weights, kind="importance"):
    """model/layers.py:87-133 (ImportancePoolingLayer), :143-195
    (WeightedMeanPoolingLayer; weights may be None), :205-236 (MaxPoolingLayer)."""
    x = np.asarray(x, dtype=np.float32)
    M = x.shape[0]
    out = []
    for i, nb in enumerate(neighbors):
        if len(nb) == 0:
            out.append(_zeros_row(x)); continue
        vi = [int(n) for n in nb if n < M]             # layers.py:109 / :165 / :225
        if not vi:
            out.append(_zeros_row(x)); continue
        feats = x[vi]
        if kind == "max":
            out.append(feats.max(axis=0)); continue
        if kind == "wmean" and (weights is None or len(weights) <= i):
            out.append(feats.mean(axis=0, dtype=np.float32)); continue     # :189-191
        wt = list(weights[i][:len(vi)])                # head truncation, :115 / :175
        s = sum(wt)
        if s == 0:
            if kind == "wmean":
                out.append(feats.mean(axis=0, dtype=np.float32)); continue  # :176-178
            wt = [1.0 / len(vi)] * len(vi)             # :116-117
        else:
            wt = [w / s for w in wt]                   # python float64, :120-121
        w32 = np.asarray(wt, dtype=np.float32)
        out.append((feats * w32[:, None]).sum(axis=0, dtype=np.float32))
    return np.stack(out) if out else np.zeros((0, x.shape[1]), np.float32)


def pool_aggregator(x, neighbors, weights=None):
    """model/aggregators.py:49-91 (WeightedAggregator) and :13-39 (MeanAggregator,
    weights=None).  No id filtering: out-of-range ids raise IndexError."""
    x = np.asarray(x, dtype=np.float32)
    out = []
    for i, nb in enumerate(neighbors):
        if len(nb) == 0:
            out.append(_zeros_row(x)); continue
        feats = x[[int(n) for n in nb]]
        if weights is None:
            out.append(feats.mean(axis=0, dtype=np.float32)); continue
        wt = list(weights[i][:len(nb)])                # aggregators.py:74
        s = sum(wt)
        if s == 0:                                     # :78-80
            out.append(feats.mean(axis=0, dtype=np.float32)); continue
        w32 = np.asarray([w / s for w in wt], dtype=np.float32)
        out.append((feats * w32[:, None]).sum(axis=0, dtype=np.float32))
    return np.stack(out) if out else np.zeros((0, x.shape[1]), np.float32)


def layer_norm(v, gamma, beta, eps=1e-5):
    mu = v.mean(axis=-1, keepdims=True)
    var = v.var(axis=-1, keepdims=True)
    return (v - mu) / np.sqrt(var + eps) * gamma + beta


def importance_aggregator(x, neighbors, weights, W, b, gamma, beta, eps=1e-5):
    """model/aggregators.py:233-287: Linear on gathered rows -> weighted sum (zero-sum
    -> mean) -> LayerNorm; empty list -> zeros(out) without LayerNorm."""
    x = np.asarray(x, dtype=np.float32)
    out_dim = W.shape[0]
    out = []
    for i, nb in enumerate(neighbors):
        if len(nb) == 0:
            out.append(np.zeros(out_dim, np.float32)); continue
        tr = x[[int(n) for n in nb]] @ W.T + b
        wt = list(weights[i][:len(nb)])
        s = sum(wt)
        if s == 0:
            agg = tr.mean(axis=0)
        else:
            agg = (tr * np.asarray([w / s for w in wt], np.float32)[:, None]).sum(axis=0)
        out.append(layer_norm(agg, gamma, beta, eps).astype(np.float32))
    return np.stack(out)


# --------------------------------------------------------------------------
# G1-G4: conv step and full forward
# --------------------------------------------------------------------------
def l2_normalize(v, eps=1e-12):
    """F.normalize(p=2, dim=1): v / max(||v||, eps)."""
    n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    return (v / np.maximum(n, eps)).astype(np.float32)


def linear(x, W, b, dtype=np.float64):
    return (x.astype(dtype) @ W.astype(dtype).T + b.astype(dtype))


def pinsage_forward(x, sd, num_layers, sampled_neighbors=None, importance_weights=None,
                    dtype=np.float64):
    """model/pinsage.py:186-251 with edge_index=None.  sd: state_dict of numpy arrays
    (reference key names).  Accumulates in float64 by default and returns float32: the oracle
    is the 'true' value both the torch reference and the CUDA path are compared against
    (dtype=np.float32 is the reference's own arithmetic, used for the CPU-baseline timing)."""
    import functools
    linear = functools.partial(globals()["linear"], dtype=dtype)
    h = np.maximum(linear(np.asarray(x, np.float32), sd["input_proj.weight"],
                          sd["input_proj.bias"]), 0).astype(np.float32)          # :202
    if sampled_neighbors is None or importance_weights is None:                   # :205-214
        for i in range(num_layers):
            h = np.maximum(linear(h, sd[f"convs.{i}.lin_self.weight"],
                                  sd[f"convs.{i}.lin_self.bias"]), 0).astype(np.float32)
    else:
        per_layer = isinstance(sampled_neighbors, list) and isinstance(importance_weights, list)
        for i in range(num_layers):                                               # :222-240
            if per_layer and len(sampled_neighbors) > i:
                nb, wt = sampled_neighbors[i], importance_weights[i]
            else:
                nb, wt = sampled_neighbors, importance_weights
            h_neigh = pool_pinsage(h, nb, wt)
            h_self = linear(h, sd[f"convs.{i}.lin_self.weight"],
                            sd[f"convs.{i}.lin_self.bias"]).astype(np.float32)
            cat = np.concatenate([h_self, h_neigh], axis=1)
            h = np.maximum(linear(cat, sd[f"convs.{i}.lin_update.weight"],
                                  sd[f"convs.{i}.lin_update.bias"]), 0).astype(np.float32)
            h = l2_normalize(h)
    emb = linear(h, sd["output_proj.weight"], sd["output_proj.bias"]).astype(np.float32)
    return l2_normalize(emb)                                                      # :248-249


def graph_conv_layer_eval(x, neigh_x, sd, eps=1e-5):
    """model/layers.py:44-77 in eval mode (BatchNorm1d uses running stats)."""
    s = linear(x, sd["linear_self.weight"], sd["linear_self.bias"])
    n = linear(neigh_x, sd["linear_neigh.weight"], sd["linear_neigh.bias"])
    out = linear(np.concatenate([s, n], axis=1), sd["linear_out.weight"], sd["linear_out.bias"])
    if out.shape[0] > 1:                                                          # :68-69
        out = (out - sd["bn.running_mean"]) / np.sqrt(sd["bn.running_var"] + eps) \
            * sd["bn.weight"] + sd["bn.bias"]
    return l2_normalize(np.maximum(out, 0).astype(np.float32))


# --------------------------------------------------------------------------
# E1 / E2: exact search
# --------------------------------------------------------------------------
def topk_total_order(scores, k, largest, ids=None):
    """k best per row under the total order (score, then id ascending)."""
    n, m = scores.shape
    ids = np.arange(m, dtype=np.int64) if ids is None else np.asarray(ids, np.int64)
    key = -scores if largest else scores
    order = np.lexsort((np.broadcast_to(ids, scores.shape), key), axis=1)[:, :k]
    return np.take_along_axis(scores, order, axis=1), ids[order] if ids.ndim == 1 else \
        np.take_along_axis(ids, order, axis=1)


def exact_ip(emb, queries, k, exclude=None):
    """utils/evaluation.py:119-130 batched: sim = q @ E^T, sim[q_idx] = -inf, topk."""
    sim = np.asarray(queries, np.float32) @ np.asarray(emb, np.float32).T
    if exclude is not None:
        ex = np.asarray(exclude)
        rows = np.nonzero(ex >= 0)[0]
        sim[rows, ex[rows]] = -np.inf
    return topk_total_order(sim, k, largest=True)


def rank_of_target(emb, query_ids, target_ids):
    """Rank (1-based) of the ground-truth item in the descending similarity order of
    utils/evaluation.py:24-33 (hit-rate: rank <= k) and :56-66 (MRR: np.where(sorted == gt)).
    Ties (measure zero on real embeddings): items with an equal score and a smaller index
    come first -- the order a stable descending sort gives."""
    e = np.asarray(emb, np.float32)
    q = np.asarray(query_ids, np.int64)
    g = np.asarray(target_ids, np.int64)
    sim = e[q] @ e.T
    s_gt = sim[np.arange(len(q)), g]
    above = (sim > s_gt[:, None]).sum(1)
    ties = ((sim == s_gt[:, None]) & (np.arange(e.shape[0])[None, :] < g[:, None])).sum(1)
    return (1 + above + ties).astype(np.int64)


def hit_rate(emb, query_ids, target_ids, k):
    """utils/evaluation.py:5-36."""
    return float((rank_of_target(emb, query_ids, target_ids) <= k).sum() / len(query_ids))


def mrr(emb, query_ids, target_ids, scale=100):
    """utils/evaluation.py:38-73: mean of 1 / (rank / scale)."""
    r = rank_of_target(emb, query_ids, target_ids)
    return float(np.mean([1.0 / (int(x) / scale) for x in r]))


def exact_l2(emb, queries, k):
    """faiss IndexFlatL2.search as used at utils/nearest_neighbors.py:174-181:
    squared L2 via ||x||^2 + ||y||^2 - 2<x,y>, clipped at 0, ascending.  UNPINNED."""
    q = np.asarray(queries, np.float32)
    e = np.asarray(emb, np.float32)
    d = (q * q).sum(1, dtype=np.float32)[:, None] + (e * e).sum(1, dtype=np.float32)[None, :] \
        - 2.0 * (q @ e.T)
    d = np.maximum(d, 0).astype(np.float32)
    return topk_total_order(d, k, largest=False)


# --------------------------------------------------------------------------
# L1 / L2 / L3: LSH  (faiss.IndexLSH(d, nbits, rotate_data=True)).  UNPINNED.
# --------------------------------------------------------------------------
def lsh_rotation(d, nbits, seed=5):
    """Same *distribution* as faiss RandomRotationMatrix(d, nbits).init(5): Gaussian,
    QR-orthonormalised; nbits > d keeps the first d columns of a nbits x nbits
    orthonormal matrix.  Not bit-identical to faiss (LAPACK / faiss RNG dependent):
    the same matrix is handed to oracle and GPU."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if nbits <= d:
        q, _ = np.linalg.qr(rng.standard_normal((d, nbits)))
        return np.ascontiguousarray(q.T.astype(np.float32))        # [nbits, d], orthonormal rows
    q, _ = np.linalg.qr(rng.standard_normal((nbits, nbits)))
    return np.ascontiguousarray(q[:, :d].astype(np.float32))       # [nbits, d]


def lsh_project(x, A):
    return np.asarray(x, np.float32) @ np.asarray(A, np.float32).T


def lsh_encode(x, A):
    """bit j = (A x)_j >= 0, packed LSB-first, (nbits+7)//8 bytes per vector."""
    y = lsh_project(x, A)
    return np.packbits(y >= 0, axis=1, bitorder="little"), y


def hamming_matrix(cq, cx):
    lut = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint16)
    out = np.zeros((cq.shape[0], cx.shape[0]), dtype=np.int32)
    for b in range(cq.shape[1]):
        out += lut[cq[:, b][:, None] ^ cx[:, b][None, :]]
    return out


def lsh_search_exhaustive(codes_x, codes_q, k):
    """hammings_knn over all stored codes; distances returned as float32
    (utils/nearest_neighbors.py:59-68)."""
    d = hamming_matrix(codes_q, codes_x)
    ds, ids = topk_total_order(d, k, largest=False)
    return ds.astype(np.float32), ids


def lsh_search_tables(codes_x, codes_q, k, num_tables, vectors=None, queries=None):
    """North-star bucketed mode: split the code into num_tables keys, candidates = items
    sharing >= 1 key with the query, re-rank by full-code Hamming (vectors None) or by
    inner product; (score, id) total order; missing results padded with id -1."""
    nb = codes_x.shape[1] * 8
    kb = nb // num_tables // 8                      # bytes per key
    ham = hamming_matrix(codes_q, codes_x)
    match = np.zeros(ham.shape, dtype=bool)
    for t in range(num_tables):
        a = codes_q[:, t * kb:(t + 1) * kb]
        b = codes_x[:, t * kb:(t + 1) * kb]
        match |= (a[:, None, :] == b[None, :, :]).all(axis=2)
    if vectors is None:
        score = np.where(match, ham, np.iinfo(np.int32).max).astype(np.float64)
        ds, ids = topk_total_order(score, k, largest=False)
        bad = ds >= np.iinfo(np.int32).max
    else:
        ip = np.asarray(queries, np.float32) @ np.asarray(vectors, np.float32).T
        score = np.where(match, ip, -np.inf)
        ds, ids = topk_total_order(score, k, largest=True)
        bad = ~np.isfinite(ds)
    ids = np.where(bad, -1, ids)
    return ds.astype(np.float32), ids, match.sum(axis=1)


# --------------------------------------------------------------------------
# I1 / I2: IVF-Flat ("Weak AND")  (faiss.IndexIVFFlat, L2).  UNPINNED.
# --------------------------------------------------------------------------
def l2sqr_direct(q, x):
    diff = q[:, None, :].astype(np.float32) - x[None, :, :].astype(np.float32)
    return (diff * diff).sum(axis=2, dtype=np.float32)


def kmeans(x, nlist, niter=20, seed=1234, max_points_per_centroid=256):
    """Lloyd iterations in the shape of faiss Clustering defaults (subsample, random
    init from the data, niter=20).  Not bit-identical to faiss."""
    x = np.asarray(x, np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    n = x.shape[0]
    if n > nlist * max_points_per_centroid:
        x = x[rng.permutation(n)[:nlist * max_points_per_centroid]]
        n = x.shape[0]
    cent = x[rng.permutation(n)[:nlist]].copy()
    for _ in range(niter):
        a = ivf_assign(x, cent)
        for c in range(nlist):
            m = a == c
            if m.any():
                cent[c] = x[m].mean(axis=0)
    return cent


def ivf_assign(x, centroids):
    x = np.asarray(x, np.float32)
    out = np.empty(x.shape[0], dtype=np.int64)
    for s in range(0, x.shape[0], 4096):
        _, ids = exact_l2(centroids, x[s:s + 4096], 1)
        out[s:s + 4096] = ids[:, 0]
    return out


def ivf_search(x, centroids, assign, queries, k, nprobe):
    """Quantizer top-nprobe lists, scan them, k smallest squared L2 (direct difference
    form, like faiss fvec_L2sqr), id -1 / +inf padding (utils/nearest_neighbors.py:134-139)."""
    x = np.asarray(x, np.float32)
    q = np.asarray(queries, np.float32)
    _, probe = exact_l2(centroids, q, min(nprobe, centroids.shape[0]))
    ds = np.full((q.shape[0], k), np.inf, dtype=np.float32)
    ids = np.full((q.shape[0], k), -1, dtype=np.int64)
    for i in range(q.shape[0]):
        cand = np.nonzero(np.isin(assign, probe[i]))[0]
        if cand.size == 0:
            continue
        d = l2sqr_direct(q[i:i + 1], x[cand])
        dd, ii = topk_total_order(d, min(k, cand.size), largest=False, ids=cand)
        ds[i, :dd.shape[1]] = dd[0]
        ids[i, :ii.shape[1]] = ii[0]
    return ds, ids


def recall_at_k(exact_ids, method_ids, k):
    """utils/nearest_neighbors.py:243-251."""
    r = 0.0
    for a, b in zip(exact_ids, method_ids):
        r += len(set(a.tolist()) & set(b.tolist())) / k
    return r / len(exact_ids)


def topk_merge(scores_list, ids_list, k, largest):
    """Multi-GPU merge restatement: concatenate per-shard lists, re-select under the
    (score, id) total order.  id < 0 entries are padding."""
    s = np.concatenate(scores_list, axis=1).astype(np.float64)
    i = np.concatenate(ids_list, axis=1).astype(np.int64)
    pad = i < 0
    s = np.where(pad, -np.inf if largest else np.inf, s)
    key = -s if largest else s
    big = np.where(pad, np.iinfo(np.int64).max, i)
    order = np.lexsort((big, key), axis=1)[:, :k]
    return np.take_along_axis(s, order, 1).astype(np.float32), np.take_along_axis(i, order, 1)
