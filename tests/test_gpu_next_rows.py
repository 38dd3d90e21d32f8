"""GPU parity for the SURVEY 8(f) "next" rows against fixtures from the unmodified reference:
N3 hard negatives (data/negative_sampler.py), N4 PPR push + top neighbours (utils/random_walk.py:144-229)
and the item-item co-occurrence graph (data/graph_builder.py:59-116)."""
import types

import numpy as np
import pytest
import torch

from tests import helpers as Hh
from tests.test_next_rows_host import _dataset

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import mre_b200  # noqa: F401
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("tag,L", [("L2", 2), ("L3", 3)])
def test_hard_negatives_equal_the_reference(dev, tag, L):
    """Same walks (shared Philox uniforms) + same global numpy RNG state => the same negatives, for the default
    window (empty with <= 300 visited nodes: random fallback), a window inside the visit list and a short one
    that needs the random top-up."""
    from mre_b200.data.negative_sampler import NegativeSampler
    from mre_b200.utils.random_walk import RandomWalkSampler
    g = Hh.load("hard_negatives.npz")
    M = int(g["M"])
    sampler = RandomWalkSampler(torch.from_numpy(g["edge_index"]), torch.from_numpy(g["edge_weights"]), walk_length=L,
                                num_walks=100, seed=int(g["seed"]), device=dev)
    neg = NegativeSampler(types.SimpleNamespace(movie_id_to_idx={i: i for i in range(M)}), sampler, num_negative_samples=20)
    queries = torch.from_numpy(g["queries"])
    for name in ("default", "window", "short"):
        lo, hi, k = (int(v) for v in g[f"{tag}_{name}_args"])
        sampler.epoch = int(g["epoch"])
        np.random.seed(99)
        got = neg.sample_hard_negatives(queries, num_hard_samples=k, max_rank=hi, min_rank=lo)
        assert got.dtype == torch.int64
        np.testing.assert_array_equal(got.numpy(), g[f"{tag}_{name}"])
    np.random.seed(7)
    np.testing.assert_array_equal(neg.sample_random_negatives(5, "cpu").numpy(), g[f"{tag}_random"])
    sampler.epoch = int(g["epoch"])
    np.random.seed(8)
    r, h = neg.sample_batch_negatives(queries, "cpu", epoch=3)
    np.testing.assert_array_equal(r.numpy(), g[f"{tag}_batch_random"])
    np.testing.assert_array_equal(h.numpy(), g[f"{tag}_batch_hard"])
    np.random.seed(8)
    r0, h0 = neg.sample_batch_negatives(queries, "cpu", epoch=0)
    assert h0 is None
    np.testing.assert_array_equal(r0.numpy(), g[f"{tag}_batch0_random"])
    with pytest.raises(ValueError):
        NegativeSampler(neg.dataset, None).sample_hard_negatives(queries)


@pytest.mark.parametrize("tag", ["bip", "gen"])
def test_ppr_push_and_top_neighbors_equal_the_reference(dev, tag):
    from mre_b200.utils.random_walk import RandomWalkSampler
    g = Hh.load("ppr.npz")
    ei, w, nodes = g[f"{tag}_edge_index"], g[f"{tag}_edge_weights"], g[f"{tag}_nodes"].tolist()
    s = RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100, device=dev)
    dense = s.compute_ppr_dense(nodes).cpu().numpy()
    np.testing.assert_allclose(dense, g[f"{tag}_ppr"], rtol=1e-12, atol=1e-300)
    assert ((dense > 0) == (g[f"{tag}_ppr"] > 0)).all()
    ppr = s.compute_ppr_matrix(nodes)
    assert [list(k) for k in ppr.keys()] == g[f"{tag}_keys_in_order"].tolist()          # dict order = the reference's
    d2 = s.compute_ppr_dense(torch.tensor(nodes[:3]), alpha=0.3, num_iterations=4).cpu().numpy()
    np.testing.assert_allclose(d2, g[f"{tag}_ppr_a03_i4"], rtol=1e-12, atol=1e-300)
    top = s.precompute_top_neighbors(nodes, num_neighbors=6)
    for r, src in enumerate(nodes):
        nb, wt = top[src]
        k = len(nb)
        assert nb == g[f"{tag}_top_ids"][r, :k].tolist() and (g[f"{tag}_top_ids"][r, k:] == -1).all()
        np.testing.assert_allclose(wt, g[f"{tag}_top_w"][r, :k], rtol=1e-12)
    with pytest.raises(IndexError):
        s.compute_ppr_matrix([s.csr.num_nodes + 3])


@pytest.mark.parametrize("thr", [1, 2, 3])
def test_item_similarity_graph_equals_the_reference(dev, thr):
    """Edges, their ORDER (first co-occurrence, as the reference's dict) and their weights."""
    from mre_b200.data.graph_builder import GraphBuilder
    g = Hh.load("item_graph.npz")
    ei, ew = GraphBuilder(_dataset(g)).build_item_similarity_graph(threshold=thr, device=dev)
    assert ei.dtype == torch.int64 and ew.dtype == torch.float32
    np.testing.assert_array_equal(ei.numpy(), g[f"sim_ei_t{thr}"])
    np.testing.assert_array_equal(ew.numpy(), g[f"sim_w_t{thr}"])


def test_item_similarity_graph_feeds_the_sampler_with_full_fan_in(dev):
    """The item-item graph makes every sampled neighbour a valid item: walks on it through
    the default sampling index equal the C oracle."""
    from oracle import oracle as O
    from mre_b200 import kernels as K
    from mre_b200.data.graph_builder import GraphBuilder
    g = Hh.load("item_graph.npz")
    ei, ew = GraphBuilder(_dataset(g)).build_item_similarity_graph(threshold=2, device=dev)
    n = int(ei.max()) + 1
    csr = K.csr_build(ei, ew, num_nodes=n, device=dev)
    row_ptr, col, cum = O.c_csr_build(ei.numpy(), ew.numpy(), n, csr.quant_shift)
    got = K.walk_topt(csr, torch.arange(n), 100, 2, 10, 3, 0)
    o = O.c_walk_topt(row_ptr, col, cum, np.arange(n), 100, 2, 10, 3, 0)
    np.testing.assert_array_equal(got[0].cpu().numpy(), o["ids"])
    np.testing.assert_array_equal(got[1].cpu().numpy(), o["counts"])
    assert (got[0].cpu().numpy() < n).all()
