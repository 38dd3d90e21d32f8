"""Property tests (hypothesis): ragged / degenerate graphs and neighbour lists.
CPU part: the bucket sampling index restatement against the flat inverse-CDF rule on arbitrary rows, and the
numpy oracle against its C restatement.  GPU part (marked): the kernels against the oracle on generated graphs
with dead ends, isolated tail nodes, duplicate edges, self loops, T larger than the number of visited nodes."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as O

SET = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@st.composite
def small_graphs(draw, max_nodes=40, max_edges=400, weights=True):
    n = draw(st.integers(2, max_nodes))
    m = draw(st.integers(0, max_edges))
    src = draw(st.lists(st.integers(0, n - 1), min_size=m, max_size=m))
    dst = draw(st.lists(st.integers(0, n - 1), min_size=m, max_size=m))
    if weights:
        w = draw(st.lists(st.integers(1, 40), min_size=m, max_size=m))       # half-star quanta, no zero weights
        w = (0.5 * np.array(w, dtype=np.float64)).astype(np.float32)
    else:
        w = None
    tail = draw(st.integers(0, 3))                                            # isolated nodes after the last edge id
    return n + tail, np.array([src, dst], dtype=np.int64).reshape(2, m), w


@settings(max_examples=60, **SET)
@given(small_graphs(), st.sampled_from([8, 6]))
def test_bucket_index_equals_flat_rule_on_arbitrary_rows(g, slots):
    n, ei, w = g
    row_ptr, col, cum = O.csr_build(ei, w, n, 1)
    built = O.walk_bucket_index(row_ptr, col, cum, slots)
    assert built is not None
    meta, leaf = built
    rng = np.random.Generator(np.random.PCG64(int(ei.sum()) % 1000))
    for v in range(n):
        a, b = int(row_ptr[v]), int(row_ptr[v + 1])
        if a == b:
            assert O.walk_bucket_pick(meta, leaf, v, 5, slots) == -1
            continue
        S = int(cum[b - 1])
        for t in set([0, S - 1] + rng.integers(0, S, 8).tolist()):
            k53 = -((-t << 53) // S)
            want = int(col[a + np.searchsorted(cum[a:b], np.uint32(t), side="right")])
            assert O.walk_bucket_pick(meta, leaf, v, k53, slots) == want


@settings(max_examples=25, **SET)
@given(small_graphs(), st.integers(1, 40), st.integers(1, 4), st.integers(1, 12), st.integers(0, 2**40))
def test_numpy_oracle_equals_c_oracle(g, W, L, T, seed):
    n, ei, w = g
    row_ptr, col, cum = O.csr_build(ei, w, n, 1)
    starts = np.arange(n)
    ids, counts, w64, nvalid = O.walk_topt(row_ptr, col, cum, starts, W, L, T, seed, 3)
    o = O.c_walk_topt(row_ptr, col, cum, starts, W, L, T, seed, 3)
    np.testing.assert_array_equal(o["ids"], ids); np.testing.assert_array_equal(o["counts"], counts)
    np.testing.assert_array_equal(o["w64"], w64); np.testing.assert_array_equal(o["nvalid"], nvalid)
    # structural properties of any sample
    assert ((ids >= 0).sum(1) == nvalid).all() and (nvalid <= min(T, W * L)).all()
    deg = np.diff(row_ptr)
    assert (nvalid[deg == 0] == 0).all()                                    # dead-end start: ([], [])
    assert np.all(np.diff(counts, axis=1)[(counts[:, 1:] > 0)] <= 0)        # counts non-increasing
    tot = w64.sum(1)
    assert np.allclose(tot[nvalid > 0], 1.0)


@pytest.mark.gpu
@settings(max_examples=30, **SET)
@given(small_graphs(max_nodes=60, max_edges=900), st.sampled_from([(100, 2, 10), (7, 5, 3), (33, 2, 32), (64, 3, 50), (1, 1, 1)]),
       st.integers(0, 2**40), st.sampled_from(["bucket", "bucket32", "compact", "wide"]))
def test_walk_kernels_equal_the_oracle_on_generated_graphs(g, shape, seed, leaf):
    import mre_b200  # noqa: F401
    from mre_b200 import kernels as K
    n, ei, w = g
    W, L, T = shape
    csr = K.csr_build(torch.from_numpy(ei), torch.from_numpy(w), num_nodes=n, index=leaf)
    row_ptr, col, cum = O.csr_build(ei, w, n, 1)
    np.testing.assert_array_equal(csr.row_ptr.cpu().numpy(), row_ptr)
    np.testing.assert_array_equal(csr.col.cpu().numpy(), col)
    starts = np.arange(n)
    got = K.walk_topt(csr, torch.from_numpy(starts), W, L, T, seed, 1, return_trace=True)
    o = O.c_walk_topt(row_ptr, col, cum, starts, W, L, T, seed, 1, return_trace=True)
    for a, key in zip(got, ["ids", "counts", "w32", "nvalid", "trace"]):
        np.testing.assert_array_equal(a.cpu().numpy(), o[key])


@pytest.mark.gpu
@settings(max_examples=30, **SET)
@given(st.integers(1, 50), st.integers(1, 70), st.integers(0, 12), st.integers(0, 10**6))
def test_pooling_equals_the_oracle_on_generated_lists(M, dim, T, seed):
    """Ragged lists with empty rows, out-of-range ids, negative (wrapping) ids, zero-sum weights, short weights."""
    import mre_b200  # noqa: F401
    from mre_b200.model.pinsage import ImportancePooling
    from mre_b200.model.layers import ImportancePoolingLayer, MaxPoolingLayer
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.standard_normal((M, dim)).astype(np.float32)
    n = int(rng.integers(1, 30))
    nbrs = [rng.integers(-M, M + 5, size=int(rng.integers(0, T + 1))).tolist() for _ in range(n)]
    wts = [(rng.integers(0, 5, size=len(l)) / 2.0).tolist() for l in nbrs]
    xt = torch.from_numpy(x).cuda()
    tol = dict(rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(ImportancePooling()(xt, nbrs, wts).cpu().numpy(), O.pool_pinsage(x, nbrs, wts), **tol)
    np.testing.assert_allclose(ImportancePoolingLayer()(xt, nbrs, wts).cpu().numpy(), O.pool_layers(x, nbrs, wts, "importance"), **tol)
    np.testing.assert_allclose(MaxPoolingLayer()(xt, nbrs).cpu().numpy(), O.pool_layers(x, nbrs, None, "max"), **tol)
