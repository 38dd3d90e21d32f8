"""Pins the oracle (numpy restatement + C restatement) against fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as Hh


@pytest.mark.parametrize("name", ["walk_quant.npz", "walk_unit.npz", "walk_float.npz"])
def test_walk_topt_matches_reference(name):
    c = Hh.walk_case(name)
    qs = Hh.quant_shift_for(c["w"])
    assert (qs >= 0) == (name != "walk_float.npz")
    N = int(c["ei"].max()) + 1
    row_ptr, col, cum = O.csr_build(c["ei"], c["w"], N, quant_shift=qs)
    ids, counts, w64, nvalid = O.walk_topt(row_ptr, col, cum, c["starts"], c["W"], c["L"],
                                           c["T"], c["seed"], c["epoch"])
    # bit-exact ids (incl. order = first-visit tie-break), lengths and float64 weights
    np.testing.assert_array_equal(nvalid, c["nvalid"])
    np.testing.assert_array_equal(ids.astype(np.int64), c["ids"])
    np.testing.assert_array_equal(w64, c["weights"])
    # the C restatement agrees with the numpy one on everything incl. traces
    r2, c2, cum2 = O.c_csr_build(c["ei"], c["w"], N, quant_shift=qs)
    np.testing.assert_array_equal(r2, row_ptr); np.testing.assert_array_equal(c2, col)
    np.testing.assert_array_equal(cum2, cum)
    out = O.c_walk_topt(row_ptr, col, cum, c["starts"], c["W"], c["L"], c["T"], c["seed"],
                        c["epoch"], return_trace=True)
    np.testing.assert_array_equal(out["ids"], ids)
    np.testing.assert_array_equal(out["counts"], counts)
    np.testing.assert_array_equal(out["w64"], w64)
    np.testing.assert_array_equal(out["w32"], w64.astype(np.float32))
    _, _, _, _, trace = O.walk_topt(row_ptr, col, cum, c["starts"][:40], c["W"], c["L"], c["T"],
                                    c["seed"], c["epoch"], return_trace=True)
    np.testing.assert_array_equal(out["trace"][:40], trace)
    i3, c3, n3 = O.count_topt_from_trace(out["trace"], c["T"])
    np.testing.assert_array_equal(i3, ids); np.testing.assert_array_equal(c3, counts)
    np.testing.assert_array_equal(n3, nvalid)


def test_walk_c_oracle_thread_invariant():
    c = Hh.walk_case("walk_quant.npz")
    N = int(c["ei"].max()) + 1
    row_ptr, col, cum = O.c_csr_build(c["ei"], c["w"], N, 1)
    a = O.c_walk_topt(row_ptr, col, cum, c["starts"], 100, 2, 10, 1, 0, num_threads=1)
    b = O.c_walk_topt(row_ptr, col, cum, c["starts"], 100, 2, 10, 1, 0, num_threads=4)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


def test_pooling_variants_match_reference():
    g = Hh.load("pooling.npz")
    x = g["x"]
    nbrs, wts, nb_ok, wt_ok = Hh.lists_from_json(g, "nbrs", "wts", "nb_ok", "wt_ok")
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.pool_pinsage(x, nbrs, wts), g["pinsage"], **tol)
    nb_int, wt_int = list(nbrs), list(wts)
    nb_int[5], wt_int[5] = 7, 0.3
    nb_int[14], wt_int[14] = np.int64(x.shape[0] + 3), 1.0
    np.testing.assert_allclose(O.pool_pinsage(x, nb_int, wt_int), g["pinsage_bareint"], **tol)
    np.testing.assert_allclose(O.pool_layers(x, nbrs, wts, "importance"), g["layers_importance"], **tol)
    np.testing.assert_allclose(O.pool_layers(x, nbrs, wts, "wmean"), g["layers_wmean"], **tol)
    np.testing.assert_allclose(O.pool_layers(x, nbrs, None, "wmean"), g["layers_wmean_none"], **tol)
    np.testing.assert_allclose(O.pool_layers(x, nbrs, None, "max"), g["layers_max"], **tol)
    np.testing.assert_allclose(O.pool_aggregator(x, nb_ok, wt_ok), g["agg_weighted"], **tol)
    np.testing.assert_allclose(O.pool_aggregator(x, nb_ok, None), g["agg_mean"], **tol)
    got = O.importance_aggregator(x, nb_ok, wt_ok, g["ia_W"], g["ia_b"], g["ia_gamma"], g["ia_beta"])
    np.testing.assert_allclose(got, g["agg_importance"], rtol=1e-4, atol=1e-5)
    # the three semantics really differ on dropped ids
    assert not np.allclose(g["pinsage"], g["layers_importance"])


def test_forward_matches_reference():
    g = Hh.load("forward.npz")
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd.")}
    F_, Hd, E_, layers = (int(v) for v in g["dims"])
    x = g["x"]
    assert Hh.rel_row_err(O.pinsage_forward(x, sd, layers), g["emb_mlp"]) < 1e-5
    nb0, wt0, nb1, wt1 = Hh.lists_from_json(g, "nb0", "wt0", "nb1", "wt1")
    got = O.pinsage_forward(x, sd, layers, [nb0, nb1], [wt0, wt1])
    assert Hh.rel_row_err(got, g["emb_lists"]) < 1e-5
    # G4: get_embeddings = per-layer resampling (epoch = layer) + forward
    N = int(g["edge_index"].max()) + 1
    row_ptr, col, cum = O.csr_build(g["edge_index"], g["edge_weights"], N, 1)
    nbrs, wts = [], []
    for layer in range(layers):
        ids, _c, w64, nv = O.walk_topt(row_ptr, col, cum, np.arange(x.shape[0]), int(g["W"]),
                                       int(g["L"]), int(g["T"]), int(g["seed"]), epoch=layer)
        nbrs.append([ids[r, :nv[r]].tolist() for r in range(len(nv))])
        wts.append([w64[r, :nv[r]].tolist() for r in range(len(nv))])
    got = O.pinsage_forward(x, sd, layers, nbrs, wts)
    assert Hh.rel_row_err(got, g["emb_full"]) < 1e-5
    gsd = {k[4:]: g[k] for k in g.files if k.startswith("gcl.")}
    assert Hh.rel_row_err(O.graph_conv_layer_eval(g["gx"], g["gn"], gsd), g["g_out"]) < 1e-5
    assert Hh.rel_row_err(O.graph_conv_layer_eval(g["gx"][:1], g["gn"][:1], gsd), g["g_out1"]) < 1e-5


def test_exact_ip_matches_reference():
    g = Hh.load("exact.npz")
    emb, qs = g["emb"], g["queries"]
    _, ids = O.exact_ip(emb, emb[qs], 10, exclude=qs)
    np.testing.assert_array_equal(ids, g["ids"])
    _, ids = O.exact_ip(emb, emb[qs], 10)
    np.testing.assert_array_equal(ids, g["ids_incl"])


def test_unpinned_restatements_self_consistent():
    """E2 / LSH / IVF have no reference-side golden (faiss absent): property checks only."""
    rng = np.random.Generator(np.random.PCG64(0))
    x = rng.standard_normal((300, 32)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    d, ids = O.exact_l2(x, x[:20], 5)
    assert (ids[:, 0] == np.arange(20)).all() and (d[:, 0] < 1e-5).all()
    # unit norm: L2^2 = 2 - 2 IP => same ranking as E1
    _, ids_ip = O.exact_ip(x, x[:20], 5)
    assert (ids == ids_ip).mean() > 0.95
    A = O.lsh_rotation(32, 64)
    np.testing.assert_allclose(A.T @ A, np.eye(32) * (A.T @ A)[0, 0], atol=1e-5)  # tight frame
    codes, y = O.lsh_encode(x, A)
    assert codes.shape == (300, 8)
    assert ((codes[:, 0] & 1) == (y[:, 0] >= 0)).all()       # LSB-first packing
    hd, hid = O.lsh_search_exhaustive(codes, codes[:10], 3)
    assert (hd[:, 0] == 0).all() and hd.dtype == np.float32
    cent = O.kmeans(x, 8, niter=5)
    a = O.ivf_assign(x, cent)
    ds, ii = O.ivf_search(x, cent, a, x[:10], 4, nprobe=8)    # nprobe = nlist => exact
    d2, i2 = O.exact_l2(x, x[:10], 4)
    np.testing.assert_array_equal(ii, i2)
    np.testing.assert_allclose(ds, d2, atol=1e-5)
    s, i = O.topk_merge([ds[:, :2], ds[:, 2:]], [ii[:, :2], ii[:, 2:]], 4, largest=False)
    np.testing.assert_array_equal(i, ii)


def test_rank_metrics_match_reference():
    """N2: the oracle's rank restatement reproduces the reference's hit-rate@k, MRR
    and evaluate_embeddings (utils/evaluation.py:5-104) on the committed fixture."""
    import json
    g = Hh.load("evaluation.npz")
    q, gt = g["pairs"][:, 0], g["pairs"][:, 1]
    np.testing.assert_array_equal(O.rank_of_target(g["emb"], q, gt), g["ranks"])
    for k, want in zip(g["ks"], g["hit_rates"]):
        assert O.hit_rate(g["emb"], q, gt, int(k)) == want
    assert abs(O.mrr(g["emb"], q, gt) - float(g["mrr"])) < 1e-12
    assert abs(O.mrr(g["emb"], q, gt, scale=7) - float(g["mrr_scale7"])) < 1e-12
    ev = json.loads(str(g["evaluate"]))
    assert ev["hit_rate@10"] == O.hit_rate(g["emb"], q, gt, 10) and abs(ev["mrr"] - float(g["mrr"])) < 1e-12
