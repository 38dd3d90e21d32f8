"""GPU parity: S0-S3 (CSR build, walk, count, top-T) through the C ABI vs the oracle and the
reference-generated golden fixtures.  Bit-exact bar (integer / index work)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import mre_b200  # noqa: F401
    from mre_b200 import kernels
    assert torch.cuda.is_available()
    return kernels


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("name", ["walk_quant.npz", "walk_unit.npz", "walk_float.npz"])
def test_csr_and_walk_match_reference_golden(K, name):
    c = Hh.walk_case(name)
    N = int(c["ei"].max()) + 1
    csr = K.csr_build(torch.from_numpy(c["ei"]), None if c["w"] is None else torch.from_numpy(c["w"]),
                      num_nodes=N)
    qs = Hh.quant_shift_for(c["w"])
    assert csr.quant_shift == qs and csr.cum_kind == (0 if qs >= 0 else 1)
    row_ptr, col, cum = O.csr_build(c["ei"], c["w"], N, quant_shift=qs)
    np.testing.assert_array_equal(_np(csr.row_ptr), row_ptr)
    np.testing.assert_array_equal(_np(csr.col), col)
    got_cum = _np(csr.cum).view(np.uint32) if qs >= 0 else _np(csr.cum)
    np.testing.assert_array_equal(got_cum, cum)
    assert (csr.meta is not None) == (qs >= 0)        # sampling index on quantised graphs
    flat = K.walk_topt(csr, torch.from_numpy(c["starts"]), c["W"], c["L"], c["T"], c["seed"],
                       c["epoch"], return_trace=True, use_index=False)
    ids, counts, w32, nvalid, trace = K.walk_topt(csr, torch.from_numpy(c["starts"]), c["W"], c["L"],
                                                  c["T"], c["seed"], c["epoch"], return_trace=True)
    for a, b in zip(flat, (ids, counts, w32, nvalid, trace)):   # indexed == flat search, bitwise
        np.testing.assert_array_equal(_np(a), _np(b))
    # vs the UNMODIFIED reference driven by the same uniforms
    np.testing.assert_array_equal(_np(nvalid), c["nvalid"])
    np.testing.assert_array_equal(_np(ids).astype(np.int64), c["ids"])
    np.testing.assert_array_equal(_np(w32), c["weights"].astype(np.float32))
    # vs the C oracle: counts and full traces
    o = O.c_walk_topt(row_ptr, col, cum, c["starts"], c["W"], c["L"], c["T"], c["seed"], c["epoch"],
                      return_trace=True)
    np.testing.assert_array_equal(_np(counts), o["counts"])
    np.testing.assert_array_equal(_np(trace), o["trace"])
    # counting stage alone, given the traces ("same walk traces" bar)
    i2, c2, w2, n2 = K.count_topt(trace.reshape(trace.size(0), -1), c["T"])
    np.testing.assert_array_equal(_np(i2), _np(ids)); np.testing.assert_array_equal(_np(c2), _np(counts))
    np.testing.assert_array_equal(_np(w2), _np(w32)); np.testing.assert_array_equal(_np(n2), _np(nvalid))


def test_walk_c1_full_vs_oracle(K):
    """Config C1 (2k movies / 5k users / 100k ratings), all items, both layer samples."""
    import mre_b200.synthetic as S
    M, U, R = S.CONFIGS["C1"][:3]
    ei, w = S.bipartite_graph(M, U, R, seed=0)
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=M + U)
    row_ptr, col, cum = O.c_csr_build(ei, w, M + U, 1)
    np.testing.assert_array_equal(_np(csr.row_ptr), row_ptr)
    np.testing.assert_array_equal(_np(csr.cum).view(np.uint32), cum)
    for epoch in range(2):
        ids, counts, w32, nvalid = K.walk_topt(csr, torch.arange(M), 100, 2, 10, 1234, epoch,
                                               use_index=bool(epoch))
        o = O.c_walk_topt(row_ptr, col, cum, np.arange(M), 100, 2, 10, 1234, epoch)
        np.testing.assert_array_equal(_np(ids), o["ids"])
        np.testing.assert_array_equal(_np(counts), o["counts"])
        np.testing.assert_array_equal(_np(w32), o["w32"])
        np.testing.assert_array_equal(_np(nvalid), o["nvalid"])


def test_walk_is_batching_and_order_invariant(K):
    """Counter-based RNG: a start node's result does not depend on the batch it is in (this is
    what makes 1/2/4/8-GPU sharding bitwise identical)."""
    import mre_b200.synthetic as S
    ei, w = S.bipartite_graph(500, 900, 12000, seed=2)
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=1400)
    full = K.walk_topt(csr, torch.arange(500), 100, 2, 10, 5, 0)
    perm = torch.randperm(500, generator=torch.Generator().manual_seed(0))
    part = K.walk_topt(csr, perm[:123], 100, 2, 10, 5, 0)
    for a, b in zip(full, part):
        np.testing.assert_array_equal(_np(a)[perm[:123].numpy()], _np(b))
    other = K.walk_topt(csr, torch.arange(500), 100, 2, 10, 5, 1)     # next epoch differs
    assert not np.array_equal(_np(full[0]), _np(other[0]))


@pytest.mark.parametrize("W,L,T", [(1, 1, 1), (7, 5, 3), (33, 2, 64), (100, 2, 50), (300, 4, 10)])
def test_walk_shapes_vs_oracle(K, W, L, T):
    c = Hh.walk_case("walk_quant.npz")
    N = int(c["ei"].max()) + 1
    csr = K.csr_build(torch.from_numpy(c["ei"]), torch.from_numpy(c["w"]), num_nodes=N)
    row_ptr, col, cum = O.c_csr_build(c["ei"], c["w"], N, 1)
    starts = np.arange(N)
    ids, counts, w32, nvalid = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 42, 9)
    o = O.c_walk_topt(row_ptr, col, cum, starts, W, L, T, 42, 9)
    np.testing.assert_array_equal(_np(ids), o["ids"])
    np.testing.assert_array_equal(_np(counts), o["counts"])
    np.testing.assert_array_equal(_np(w32), o["w32"])
    np.testing.assert_array_equal(_np(nvalid), o["nvalid"])


def test_sampler_dropin_api(K):
    from mre_b200.utils.random_walk import RandomWalkSampler
    c = Hh.walk_case("walk_quant.npz")
    s = RandomWalkSampler(torch.from_numpy(c["ei"]), torch.from_numpy(c["w"]), walk_length=2,
                          num_walks=100, seed=c["seed"])
    nbrs, wts = s.batch_sample_neighbors(c["starts"].tolist(), 10)      # epoch 0
    for r in range(len(nbrs)):
        nv = c["nvalid"][r]
        assert nbrs[r] == c["ids"][r, :nv].tolist()
        assert wts[r] == c["weights"][r, :nv].tolist()                   # float64, bit-exact
    assert all(isinstance(v, int) for v in nbrs[0]) and all(isinstance(v, float) for v in wts[0])
    n1, w1 = s.sample_neighbors(0, 10)                                    # epoch 1: new sample
    assert len(n1) == len(w1) <= 10 and abs(sum(w1) - 1.0) < 1e-12
    walk = s._single_walk(3)
    assert walk[0] == 3 and 1 <= len(walk) <= 3
    adj = s.adj_list
    assert len(adj) == int(c["ei"].max()) + 1
    e0 = [(int(d), float(w)) for sidx, d, w in zip(c["ei"][0], c["ei"][1], c["w"]) if sidx == 0]
    assert adj[0] == e0                                                   # edge order preserved
    with pytest.raises(IndexError):
        s.sample_neighbors(10**6)
    for lonely in (420, 422):             # sink (in-edges only) and isolated node of the fixture
        assert adj[lonely] == []
        empty_n, empty_w = s.sample_neighbors(lonely)
        assert empty_n == [] and empty_w == []


@pytest.mark.parametrize("leaf_format", ["wide", "compact"])
def test_walk_index_structure_and_deep_rows(K, leaf_format, monkeypatch):
    """The 8-ary sampling index: block layout vs a numpy restatement (64-byte leaves, and the
    compact 32-byte leaves: u8 separator-minus-cum, u16 id low halves, u8 id high bytes), and rows
    deep enough for 3-4 tree levels (degree up to 5000) searched identically to the flat binary
    search."""
    from mre_b200 import _native as NV
    monkeypatch.setenv("PB200_WALK_LEAF", leaf_format)
    rng = np.random.Generator(np.random.PCG64(3))
    degs = [0, 1, 7, 8, 9, 63, 64, 65, 511, 512, 513, 4097, 5000]
    src = np.concatenate([np.full(d, v) for v, d in enumerate(degs)]).astype(np.int64)
    N = 6000
    dst = rng.integers(0, len(degs), size=src.size).astype(np.int64)
    w = (0.5 * rng.integers(0, 11, size=src.size)).astype(np.float32)      # zero weights allowed
    w[np.cumsum(degs)[1:] - 1] = 2.5                                        # rows keep a positive total
    perm = rng.permutation(src.size)                                        # edge order != row order
    ei = np.stack([src[perm], dst[perm]])
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w[perm]), num_nodes=N)
    row_ptr, col, cum = O.csr_build(ei, w[perm], N, 1)
    meta = _np(csr.meta).view(np.uint32); leaf = _np(csr.leaf).view(np.uint32); idx = _np(csr.idx).view(np.uint32)
    assert csr.leaf_format == (NV.LEAF_WIDE if leaf_format == "wide" else NV.LEAF_COMPACT)
    assert leaf.shape[1] == (16 if leaf_format == "wide" else 8)
    lo = io = 0
    for v, d in enumerate(degs):
        a = row_ptr[v]
        nb0 = (d + 7) // 8
        assert meta[v].tolist() == [lo, d, int(cum[a + d - 1]) if d else 0, io]
        keys = np.full(nb0 * 8, 0xFFFFFFFF, np.uint32); keys[:d] = cum[a:a + d]
        cols = np.full(nb0 * 8, 0xFFFFFFFF, np.uint32); cols[:d] = col[a:a + d].view(np.uint32)
        if leaf_format == "wide":
            np.testing.assert_array_equal(leaf[lo:lo + nb0, :8].reshape(-1), keys)
            np.testing.assert_array_equal(leaf[lo:lo + nb0, 8:].reshape(-1), cols)
        elif nb0:
            blk = np.ascontiguousarray(leaf[lo:lo + nb0]).view(np.uint8).reshape(nb0, 32)
            last = np.minimum(8 * np.arange(nb0) + 7, d - 1)
            sep = cum[a + last].astype(np.int64)                               # separator of every block
            want_delta = np.where(keys.reshape(nb0, 8) != 0xFFFFFFFF, sep[:, None] - keys.reshape(nb0, 8).astype(np.int64), 0)
            assert want_delta.max() <= 255
            np.testing.assert_array_equal(blk[:, :8], want_delta.astype(np.uint8))
            ids = np.where(cols.reshape(nb0, 8) != 0xFFFFFFFF, cols.reshape(nb0, 8), 0xFFFFFF).astype(np.uint32)
            np.testing.assert_array_equal(blk[:, 8:24].copy().view(np.uint16).reshape(nb0, 8), (ids & 0xFFFF).astype(np.uint16))
            np.testing.assert_array_equal(blk[:, 24:32], (ids >> 16).astype(np.uint8))
        levels, nb = [], nb0
        while nb > 1:
            levels.append(nb); nb = (nb + 7) // 8
        off = io
        for l in range(len(levels), 0, -1):               # top level first
            nkeys, nblk = levels[l - 1], (levels[l - 1] + 7) // 8
            want = np.full(nblk * 8, 0xFFFFFFFF, np.uint32)
            last = np.minimum(8 ** l * (np.arange(nkeys) + 1), d) - 1
            want[:nkeys] = cum[a + last]
            np.testing.assert_array_equal(idx[off:off + nblk].reshape(-1), want)
            off += nblk
        lo += nb0; io = off
    assert leaf.shape[0] == lo and idx.shape[0] == io
    starts = np.arange(len(degs))
    got = K.walk_topt(csr, torch.from_numpy(starts), 200, 3, 20, 77, 0, return_trace=True)
    flat = K.walk_topt(csr, torch.from_numpy(starts), 200, 3, 20, 77, 0, return_trace=True, use_index=False)
    o = O.c_walk_topt(row_ptr, col, cum, starts, 200, 3, 20, 77, 0, return_trace=True)
    for a, b, key in zip(got, flat, ["ids", "counts", "w32", "nvalid", "trace"]):
        np.testing.assert_array_equal(_np(a), _np(b))
        np.testing.assert_array_equal(_np(a), o[key])


def test_compact_leaf_falls_back_to_wide_for_heavy_blocks(K, monkeypatch):
    """Blocks spanning more than 255 weight quanta (or >= 2^24 nodes) cannot use the 32-byte leaf:
    the builder must pick the 64-byte format and the walks must still match the flat search."""
    from mre_b200 import _native as N
    monkeypatch.setenv("PB200_WALK_LEAF", "compact")
    rng = np.random.Generator(np.random.PCG64(8))
    n_edges = 4000
    src = rng.integers(0, 50, n_edges).astype(np.int64)
    dst = rng.integers(0, 50, n_edges).astype(np.int64)
    w = (0.5 * rng.integers(1, 200, n_edges)).astype(np.float32)           # up to 199 quanta per edge
    ei = np.stack([src, dst])
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=50)
    assert csr.meta is not None and csr.leaf_format == N.LEAF_WIDE
    a = K.walk_topt(csr, torch.arange(50), 64, 3, 8, 5, 0, return_trace=True)
    b = K.walk_topt(csr, torch.arange(50), 64, 3, 8, 5, 0, return_trace=True, use_index=False)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(_np(x), _np(y))


def _deep_graph(rng, degs, wmax, N=6000):
    src = np.concatenate([np.full(d, v) for v, d in enumerate(degs)]).astype(np.int64)
    dst = rng.integers(0, len(degs), size=src.size).astype(np.int64)
    w = (0.5 * rng.integers(1, wmax + 1, size=src.size)).astype(np.float32)      # no zero weights
    perm = rng.permutation(src.size)
    return np.stack([src[perm], dst[perm]]), w[perm]


@pytest.mark.parametrize("wmax", [1, 10, 120])
def test_bucket_index_structure_and_walks(K, wmax):
    """The direct-addressed bucket index (default): meta and every 32-byte bucket equal the numpy
    restatement of the format, and walks through it -- lean kernel (W*L <= 200, with and without
    trace, L = 2 and generic L) and the generic kernel (larger W*L) -- equal the flat search and the C oracle."""
    from mre_b200 import _native as NV
    rng = np.random.Generator(np.random.PCG64(30 + wmax))
    degs = [0, 1, 7, 8, 9, 63, 64, 65, 511, 512, 513, 4097, 5000]
    ei, w = _deep_graph(rng, degs, wmax)
    N = 6000
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=N)
    assert csr.leaf_format == NV.LEAF_BUCKET and csr.idx is None
    row_ptr, col, cum = O.csr_build(ei, w, N, 1)
    meta, leaf = O.walk_bucket_index(row_ptr, col, cum)
    got_meta = _np(csr.meta).view(np.uint32)
    has = meta[:, 1] > 0
    np.testing.assert_array_equal(got_meta[:, 1:], meta[:, 1:])
    np.testing.assert_array_equal(got_meta[has, 0], meta[has, 0])
    np.testing.assert_array_equal(np.ascontiguousarray(_np(csr.leaf)).view(np.uint8).reshape(-1, 32), leaf)
    starts = np.arange(len(degs))
    for (W, L, T) in [(100, 2, 10), (64, 3, 8), (200, 3, 20), (33, 1, 32), (7, 5, 3)]:
        got = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 77, 3, return_trace=True)
        flat = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 77, 3, return_trace=True, use_index=False)
        o = O.c_walk_topt(row_ptr, col, cum, starts, W, L, T, 77, 3, return_trace=True)
        for a, b, key in zip(got, flat, ["ids", "counts", "w32", "nvalid", "trace"]):
            np.testing.assert_array_equal(_np(a), _np(b))
            np.testing.assert_array_equal(_np(a), o[key])
        notrace = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 77, 3)
        for a, b in zip(notrace, got[:4]):
            np.testing.assert_array_equal(_np(a), _np(b))


@pytest.mark.parametrize("wmax", [1, 10, 120])
def test_bucket32_index_structure_and_walks(K, wmax):
    """PB200_LEAF_BUCKET32 (six slots, 32-bit ids -- what graphs with more than 2^24 nodes get): the index
    equals the numpy restatement, and walks through it equal the flat search and the C oracle, with
    neighbour ids above 2^24 (the 24-bit forms would truncate them)."""
    from mre_b200 import _native as NV
    rng = np.random.Generator(np.random.PCG64(60 + wmax))
    degs = [0, 1, 5, 6, 7, 8, 9, 63, 64, 65, 511, 512, 513, 4097, 5000]
    ei, w = _deep_graph(rng, degs, wmax)
    N = (1 << 24) + 6000
    ei = ei.copy()
    ei[1] += np.where(ei[1] % 3 == 0, 1 << 24, 0)                  # a third of the destinations beyond 24 bits
    back = np.stack([ei[1][:4000], ei[0][:4000]])                  # some edges back so that step 2 continues
    ei2 = np.concatenate([ei, back], axis=1); w2 = np.concatenate([w, w[:4000]])
    csr = K.csr_build(torch.from_numpy(ei2), torch.from_numpy(w2), num_nodes=N)
    assert csr.leaf_format == NV.LEAF_BUCKET32 and csr.idx is None      # auto: > 2^24 nodes
    row_ptr, col, cum = O.csr_build(ei2, w2, N, 1)
    assert col.max() >= (1 << 24)
    # structure on the rows that have edges (the restatement loops over rows in Python: compare on a compacted copy)
    rows = np.flatnonzero(np.diff(row_ptr))[:200]
    sub_ptr = np.zeros(len(rows) + 1, np.int64); sub_ptr[1:] = np.cumsum(np.diff(row_ptr)[rows])
    sel = np.concatenate([np.arange(row_ptr[v], row_ptr[v + 1]) for v in rows])
    meta, leaf = O.walk_bucket_index(sub_ptr, col[sel], cum[sel], slots=6)
    got_meta = _np(csr.meta).view(np.uint32)[rows]
    np.testing.assert_array_equal(got_meta[:, 1:], meta[:, 1:])
    got_leaf = np.ascontiguousarray(_np(csr.leaf)).view(np.uint8).reshape(-1, 32)
    for i in range(len(rows)):
        nb = int(((meta[i, 2] - 1) >> meta[i, 3]) + 1)
        np.testing.assert_array_equal(got_leaf[got_meta[i, 0]:got_meta[i, 0] + nb], leaf[meta[i, 0]:meta[i, 0] + nb])
    starts = np.concatenate([np.arange(len(degs)), np.unique(ei2[0][ei2[0] >= (1 << 24)])[:50]])
    for (W, L, T) in [(100, 2, 10), (64, 3, 8), (200, 3, 20), (33, 1, 32), (7, 5, 3)]:
        got = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 77, 3, return_trace=True)
        flat = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 77, 3, return_trace=True, use_index=False)
        o = O.c_walk_topt(row_ptr, col, cum, starts, W, L, T, 77, 3, return_trace=True)
        for a, b, key in zip(got, flat, ["ids", "counts", "w32", "nvalid", "trace"]):
            np.testing.assert_array_equal(_np(a), _np(b))
            np.testing.assert_array_equal(_np(a), o[key])
        assert _np(got[4]).max() >= (1 << 24)
        notrace = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, 77, 3)
        for a, b in zip(notrace, got[:4]):
            np.testing.assert_array_equal(_np(a), _np(b))
    multi = K.walk_topt(csr, torch.from_numpy(starts), 100, 2, 10, 77, 3, num_epochs=2)
    for e in range(2):
        one = K.walk_topt(csr, torch.from_numpy(starts), 100, 2, 10, 77, 3 + e)
        for a, b in zip(multi, one):
            np.testing.assert_array_equal(_np(a[e]), _np(b))


def test_bucket_index_fallbacks(K):
    """Graphs the bucket format does not cover keep the tree index: a zero-weight edge, weights so
    heavy that buckets would dwarf the edge list.  An explicit request raises."""
    from mre_b200 import _native as NV
    rng = np.random.Generator(np.random.PCG64(9))
    ei, w = _deep_graph(rng, [5, 40, 300], 10, N=400)
    w0 = w.copy(); w0[3] = 0.0
    assert K.csr_build(torch.from_numpy(ei), torch.from_numpy(w0), num_nodes=400).leaf_format != NV.LEAF_BUCKET
    heavy = (w * 4096).astype(np.float32)
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(heavy), num_nodes=400)
    assert csr.leaf_format != NV.LEAF_BUCKET and csr.meta is not None
    with pytest.raises(NV.NativeError):
        K.csr_build(torch.from_numpy(ei), torch.from_numpy(w0), num_nodes=400, index="bucket")
    assert K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=400).leaf_format == NV.LEAF_BUCKET


def test_walk_distribution_matches_the_unpatched_reference_sampler(K):
    """the Philox kernel against the UNPATCHED reference (global MT19937, np.random.choice):
    visit frequencies of step 1 and step 2 over 20,000 walks per start node.  Two independent samples of
    the same multinomial: chi-square homogeneity statistic within 4.5 sigma of its mean (df) and total
    variation distance below the sampling-noise bound."""
    g = Hh.load("walk_distribution.npz")
    ei, w, starts, n_walks = g["edge_index"], g["edge_weights"], g["starts"], int(g["n_walks"])
    ref = g["hist"].astype(np.float64)                                  # [starts, 2, N]
    N = ref.shape[2]
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=N)
    got = np.zeros_like(ref)
    per_launch = 100
    for e in range(n_walks // per_launch):                             # 200 epochs x 100 walks
        *_x, trace = K.walk_topt(csr, torch.from_numpy(starts), per_launch, 2, 1, 777, e, return_trace=True)
        tr = trace.cpu().numpy()                                        # [starts, 100, 2]
        for si in range(len(starts)):
            for step in range(2):
                v = tr[si, :, step]
                got[si, step] += np.bincount(v[v >= 0], minlength=N)
    for si in range(len(starts)):
        for step in range(2):
            a, b = ref[si, step], got[si, step]
            assert a.sum() == n_walks and b.sum() == n_walks
            keep = (a + b) >= 10                                        # pool the rare cells
            a2 = np.append(a[keep], a[~keep].sum()); b2 = np.append(b[keep], b[~keep].sum())
            nz = (a2 + b2) > 0
            chi2 = float((((a2 - b2) ** 2)[nz] / (a2 + b2)[nz]).sum())  # two-sample chi-square, equal totals
            df = int(nz.sum()) - 1
            assert abs(chi2 - df) < 4.5 * np.sqrt(2 * df) + 5, (si, step, chi2, df)
            tv = 0.5 * np.abs(a / n_walks - b / n_walks).sum()
            assert tv < 2.5 * np.sqrt(df / (2 * np.pi * n_walks)) + 0.01, (si, step, tv)


@pytest.mark.parametrize("leaf", ["bucket", "compact"])
def test_multi_epoch_launch_equals_separate_calls(K, leaf):
    """num_epochs = E: one launch over (epoch, start) pairs (bucket index) or a loop of launches (tree
    index) == E separate calls with epochs e .. e + E - 1, bitwise, traces included; sample_layers returns
    the same NeighborBatches as consecutive batch_sample_neighbors_tensor calls."""
    import mre_b200.synthetic as S
    from mre_b200.utils.random_walk import RandomWalkSampler
    ei, w = S.bipartite_graph(400, 900, 15000, seed=4)
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=1300, index=leaf)
    starts = torch.arange(0, 400, 3)
    multi = K.walk_topt(csr, starts, 100, 2, 10, 9, 5, return_trace=True, num_epochs=3)
    for e in range(3):
        one = K.walk_topt(csr, starts, 100, 2, 10, 9, 5 + e, return_trace=True)
        for a, b in zip(multi, one):
            np.testing.assert_array_equal(_np(a[e]), _np(b))
    if leaf == "bucket":
        s1 = RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100, seed=9, num_nodes=1300)
        s2 = RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100, seed=9, num_nodes=1300)
        layers = s1.sample_layers(starts, 10, 2)
        assert s1.epoch == 2
        for l in range(2):
            ref = s2.batch_sample_neighbors_tensor(starts, 10)
            for a, b in zip(layers[l].as_args()[:3], ref.as_args()[:3]):
                np.testing.assert_array_equal(_np(a), _np(b))


def test_integration_recipe_a_module_aliasing_runs_the_reference_imports():
    """INTEGRATION.md recipe A, executed in a fresh interpreter: alias the mirror modules into sys.modules,
    then use the REFERENCE's own import lines; the sampler obtained that way reproduces the reference golden."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys
sys.path.insert(0, %r)
import mre_b200
for name in ("utils.random_walk", "utils.nearest_neighbors", "utils.evaluation", "model.pinsage", "model.layers",
             "model.aggregators", "data.negative_sampler", "data.graph_builder"):
    pkg = name.split(".")[0]
    sys.modules.setdefault(pkg, __import__("mre_b200." + pkg, fromlist=["_"]))
    sys.modules[name] = __import__("mre_b200." + name, fromlist=["_"])
from utils.random_walk import RandomWalkSampler                      # reference inference.py:10
from model.pinsage import PinSage                                     # reference inference.py:9
from utils.nearest_neighbors import LSHIndex, WeakANDIndex            # reference inference.py:11
from data.negative_sampler import NegativeSampler
import numpy as np, torch
from tests import helpers as Hh
c = Hh.walk_case("walk_quant.npz")
s = RandomWalkSampler(torch.from_numpy(c["ei"]), torch.from_numpy(c["w"]), walk_length=2, num_walks=100, seed=c["seed"])
nbrs, wts = s.batch_sample_neighbors(c["starts"].tolist(), 10)
for r in range(len(nbrs)):
    nv = c["nvalid"][r]
    assert nbrs[r] == c["ids"][r, :nv].tolist() and wts[r] == c["weights"][r, :nv].tolist()
model = PinSage(16, 32, 16, 2).cuda().eval()
emb = model.get_embeddings(torch.randn(120, 16), s, 10)
assert emb.shape == (120, 16) and emb.device.type == "cpu"
print("RECIPE_A_OK")
''' % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0 and "RECIPE_A_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_integration_recipe_b_ctypes_stub_from_the_document():
    """INTEGRATION.md recipe B: the ctypes stub is cut out of the document and executed as written (library path
    relative to the repository root); the two methods it gives the reference's RandomWalkSampler reproduce the C
    oracle's lists and float64 weights."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    sec = text[text.index("## B."):]
    code = sec[sec.index("```python") + len("```python"):]
    code = code[:code.index("```")]
    ns = {}
    cwd = os.getcwd()
    os.chdir(root)
    try:
        exec(compile(code, "INTEGRATION.md#B", "exec"), ns)
    finally:
        os.chdir(cwd)
    import mre_b200.synthetic as S
    ei, w = S.bipartite_graph(300, 700, 9000, seed=8)
    cls = ns["RandomWalkSampler"]
    smp = cls.__new__(cls)                                  # the reference's __init__ (utils/random_walk.py:11-31) sets these
    smp.edge_index, smp.edge_weights = torch.from_numpy(ei), torch.from_numpy(w)
    smp.num_walks, smp.walk_length, smp.seed, smp.epoch = 100, 2, 77, 3
    smp._prepare_adjacency_list()
    starts = list(range(0, 300, 5))
    nbrs, wts = smp.batch_sample_neighbors(starts, 10)
    assert smp.epoch == 4
    N = int(ei.max()) + 1
    row_ptr, col, cum = O.csr_build(ei, w, N, 1)
    o = O.c_walk_topt(row_ptr, col, cum, np.array(starts), 100, 2, 10, 77, 3)
    for i in range(len(starts)):
        k = int(o["nvalid"][i])
        assert nbrs[i] == o["ids"][i, :k].tolist()
        assert wts[i] == o["w64"][i, :k].tolist()
