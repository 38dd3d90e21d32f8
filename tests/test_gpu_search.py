"""GPU parity: E1/E2 exact search, L1-L3 LSH, I1/I2 IVF, top-k merge through the C ABI.
E1 is pinned by the reference golden; E2/LSH/IVF are compared with the faiss restatements in
the oracle (PARITY UNPINNED: faiss is absent offline), conditional on shared parameters."""
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


def _data(n, d, seed, normalise=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.standard_normal((n, d)).astype(np.float32)
    if normalise:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def test_exact_ip_vs_reference_golden(K):
    from mre_b200.utils.evaluation import generate_recommendations, generate_recommendations_batch
    g = Hh.load("exact.npz")
    emb = torch.from_numpy(g["emb"]).cuda()
    for q, want in zip(g["queries"], g["ids"]):
        np.testing.assert_array_equal(generate_recommendations(emb, int(q), k=10), want)
    np.testing.assert_array_equal(generate_recommendations_batch(emb, g["queries"], 10, False), g["ids_incl"])
    got = generate_recommendations(torch.from_numpy(g["emb"]), 3, k=10)        # CPU tensor input
    np.testing.assert_array_equal(got, g["ids"][1])
    assert got.dtype == np.int64


@pytest.mark.parametrize("nq,nx,d,k", [(1, 1, 1, 1), (5, 40, 7, 10), (70, 3000, 64, 10), (130, 1000, 128, 32),
                                       (3, 200, 33, 50), (64, 65, 16, 11)])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_exact_topk_vs_oracle(K, nq, nx, d, k, metric):
    from mre_b200 import _native as N
    x, q = _data(nx, d, 1), _data(nq, d, 2)
    kk = min(k, nx)
    s, i = K.topk(torch.from_numpy(q), torch.from_numpy(x), k, N.METRIC_IP if metric == "ip" else N.METRIC_L2)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    rs, ri = (O.exact_ip if metric == "ip" else O.exact_l2)(x, q, kk)
    np.testing.assert_allclose(s[:, :kk], rs, atol=2e-5)                  # compare by score
    assert (i[:, :kk] == ri).mean() > 0.98                                # ids, up to fp ties
    if kk < k:                                                            # padding convention
        assert (i[:, kk:] == -1).all()
        assert np.all(np.isinf(s[:, kk:]))
    # the returned ids really have the returned scores
    full = q @ x.T if metric == "ip" else np.maximum((q * q).sum(1)[:, None] + (x * x).sum(1)[None] - 2 * q @ x.T, 0)
    np.testing.assert_allclose(np.take_along_axis(full, i[:, :kk].astype(np.int64), 1), s[:, :kk], atol=2e-5)


def test_topk_exclude_offset_and_merge(K):
    """Item-sharded search + merge == unsharded search (the multi-GPU path on one GPU)."""
    from mre_b200 import _native as N
    x, q = _data(5000, 48, 3), _data(100, 48, 4)
    xt, qt = torch.from_numpy(x).cuda(), torch.from_numpy(q).cuda()
    excl = torch.arange(100, dtype=torch.int32)
    s_full, i_full = K.topk(qt, xt, 10, N.METRIC_IP, exclude_ids=excl)
    parts = [K.topk(qt, xt[a:b].contiguous(), 10, N.METRIC_IP, exclude_ids=excl, id_offset=a)
             for a, b in [(0, 1300), (1300, 2600), (2600, 3800), (3800, 5000)]]
    s_m, i_m = K.topk_merge(torch.cat([p[0] for p in parts], 1), torch.cat([p[1] for p in parts], 1), 10, True)
    np.testing.assert_array_equal(i_m.cpu().numpy(), i_full.cpu().numpy())      # bitwise: total order
    np.testing.assert_array_equal(s_m.cpu().numpy(), s_full.cpu().numpy())
    rs, ri = O.topk_merge([p[0].cpu().numpy() for p in parts], [p[1].cpu().numpy() for p in parts], 10, True)
    np.testing.assert_array_equal(i_m.cpu().numpy(), ri)
    assert not (i_full.cpu().numpy() == np.arange(100)[:, None]).any()          # self excluded


def test_lsh_codes_and_exhaustive_search(K):
    """L1/L2: codes bit-exact except projections within 1e-6 of zero; Hamming top-k exact."""
    from mre_b200.utils.nearest_neighbors import LSHIndex
    d, nbits, n = 128, 256, 3000
    x = _data(n, d, 5)
    A = O.lsh_rotation(d, nbits)
    ref_codes, y = O.lsh_encode(x, A)
    codes, y_gpu = K.lsh_encode(torch.from_numpy(x), torch.from_numpy(A), return_projection=True)
    codes = codes.cpu().numpy()
    diff_bits = np.unpackbits(codes ^ ref_codes, axis=1, bitorder="little").astype(bool)
    assert np.all(np.abs(y[diff_bits]) < 1e-6)
    np.testing.assert_allclose(y_gpu.cpu().numpy(), y, atol=1e-5)
    idx = LSHIndex(d, nbits, 16, projection=A)
    idx.build(x)
    assert idx.index.ntotal == n
    dist, ids = idx.search(x[:200], k=10)
    assert dist.dtype == np.float32 and ids.dtype == np.int64 and dist.shape == (200, 10)
    rd, ri = O.lsh_search_exhaustive(codes, codes[:200], 10)
    np.testing.assert_array_equal(dist, rd)                                 # Hamming ints, exact
    np.testing.assert_array_equal(ids, ri)                                  # (dist, id) total order
    assert (ids[:, 0] == np.arange(200)).all() and (dist[:, 0] == 0).all()
    d2, i2 = idx.search(x[:7], k=40)                                        # k > 32: multi-pass
    rd2, ri2 = O.lsh_search_exhaustive(codes, codes[:7], 40)
    np.testing.assert_array_equal(d2, rd2); np.testing.assert_array_equal(i2, ri2)


@pytest.mark.parametrize("rerank", ["hamming", "dot"])
def test_lsh_tables_mode(K, rerank):
    """L3 (north-star bucketed mode): candidates = items sharing >= 1 of the 16 keys; exact dedup."""
    from mre_b200.utils.nearest_neighbors import LSHIndex
    import mre_b200.synthetic as S
    d, nbits, n = 64, 256, 4000
    x = S.spread_embeddings(n, d, seed=1, clusters=64, noise=0.05).numpy()
    A = O.lsh_rotation(d, nbits)
    idx = LSHIndex(d, nbits, 16, mode="tables", rerank=rerank, projection=A)
    idx.build(x)
    q = x[:300]
    dist, ids = idx.search(q, k=10)
    codes = idx.codes.cpu().numpy()
    rd, ri, ncand = O.lsh_search_tables(codes, codes[:300], 10, 16,
                                        vectors=x if rerank == "dot" else None, queries=q)
    np.testing.assert_array_equal(idx.last_num_candidates.cpu().numpy(), ncand)   # dedup is exact
    if rerank == "hamming":
        np.testing.assert_array_equal(ids, ri)
        np.testing.assert_array_equal(dist[ri >= 0], rd[ri >= 0])
    else:
        ok = ri >= 0
        np.testing.assert_allclose(dist[ok], rd[ok], atol=2e-5)
        assert (ids == ri).mean() > 0.98
    assert ((ids == -1) == (ri == -1)).all()


def test_ivf_weak_and(K):
    """I1/I2 with shared centroids: assignments, list layout and search vs the restatement."""
    from mre_b200.utils.nearest_neighbors import WeakANDIndex, train_kmeans
    import mre_b200.synthetic as S
    n, d, nlist = 6000, 32, 100
    x = S.spread_embeddings(n, d, seed=2, clusters=80, noise=0.2).numpy()
    cent = O.kmeans(x, nlist, niter=5)
    idx = WeakANDIndex(d, nlist, 10, centroids=cent)
    idx.build(x)
    assert idx.index.ntotal == n and idx.quantizer.ntotal == nlist
    a_ref = O.ivf_assign(x, cent)
    a_gpu = idx.assign.cpu().numpy()
    assert (a_gpu == a_ref).mean() > 0.999
    offsets, list_ids, _ = idx._lists
    offsets, list_ids = offsets.cpu().numpy(), list_ids.cpu().numpy()
    for l in (0, 17, 99):                                   # ascending ids inside a list
        seg = list_ids[offsets[l]:offsets[l + 1]]
        np.testing.assert_array_equal(seg, np.nonzero(a_gpu == l)[0])
    dist, ids = idx.search(x[:250], k=10)
    assert idx.index.nprobe == 20
    rd, ri = O.ivf_search(x, cent, a_gpu, x[:250], 10, 20)
    np.testing.assert_allclose(dist, rd, atol=2e-5)
    assert (ids == ri).mean() > 0.98
    # tiny lists: fewer than k results -> -1 padding
    few = WeakANDIndex(d, 4, centroids=x[:4] * 10.0 + np.eye(4, d, dtype=np.float32) * 100)
    few.build(x[:6])
    dd, ii = few.search(x[:2], k=10)
    assert (ii[:, 6:] == -1).all() and np.isinf(dd[:, 6:]).all()
    # GPU k-means: every Lloyd iteration does not increase the quantisation error
    xt = torch.from_numpy(x).cuda()
    c5 = train_kmeans(xt, 16, niter=5)
    c1 = train_kmeans(xt, 16, niter=1)
    err = lambda c: float(K.topk(xt, c, 1, 1)[0].sum())
    assert err(c5) <= err(c1) * 1.0001


def test_benchmark_search_methods_dict(K):
    from mre_b200.utils.nearest_neighbors import benchmark_search_methods
    import mre_b200.synthetic as S
    x = S.spread_embeddings(3000, 64, seed=3, clusters=50, noise=0.2)
    res = benchmark_search_methods(x, x[:100], k=10)
    assert set(res) == {"exact", "lsh", "ivf"}
    for m in res.values():
        assert {"distances", "indices", "search_time", "index_size", "method"} <= set(m)
        assert m["indices"].shape == (100, 10) and m["index_size"] == 3000
    assert res["ivf"]["recall"] > 0.8 and 0.0 <= res["lsh"]["recall"] <= 1.0
    assert (res["exact"]["indices"][:, 0] == np.arange(100)).all()


# ---- tensor-core exact search (pb200_topk_tc): bitwise equal to the fp32 kernel -----------------
def _tc_vs_fp32(K, q, x, k, metric, exclude=None, id_offset=0):
    from mre_b200 import _native as N
    m = N.METRIC_IP if metric == "ip" else N.METRIC_L2
    qd, xd = torch.from_numpy(q).cuda(), torch.from_numpy(x).cuda()
    ex = None if exclude is None else torch.from_numpy(exclude).cuda()
    st = {}
    s1, i1 = K.topk(qd, xd, k, m, exclude_ids=ex, id_offset=id_offset, precision="tf32", stats=st)
    s0, i0 = K.topk(qd, xd, k, m, exclude_ids=ex, id_offset=id_offset, precision="fp32")
    assert st["path"] == "tf32"
    np.testing.assert_array_equal(i1.cpu().numpy(), i0.cpu().numpy())
    np.testing.assert_array_equal(s1.cpu().numpy(), s0.cpu().numpy())
    return int(st["fp32_reruns"].item())


@pytest.mark.parametrize("nq,nx,d,k", [(300, 5000, 128, 10), (257, 1300, 64, 10), (1000, 9000, 128, 20),
                                       (129, 700, 256, 10), (40, 128, 32, 5), (600, 100, 100, 12),
                                       (256, 4096, 4, 3), (1, 3000, 128, 10), (513, 20000, 96, 24)])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_exact_topk_tensor_core_bitwise_equals_fp32(K, nq, nx, d, k, metric):
    x, q = _data(nx, d, 11), _data(nq, d, 12)
    _tc_vs_fp32(K, q, x, k, metric)


def test_exact_topk_tensor_core_self_queries_exclude_offset(K):
    """All-item self queries (the E1 call pattern): exclude_ids drops the query itself, id_offset
    shifts ids (item-sharded search); unnormalised vectors exercise the |q| max|x| bound."""
    x = _data(7000, 128, 21, normalise=False) * 3.0
    excl = np.arange(7000, dtype=np.int32) + 1000
    reruns = _tc_vs_fp32(K, x, x, 10, "ip", exclude=excl, id_offset=1000)
    assert reruns < 7000 // 4
    _tc_vs_fp32(K, x, x, 10, "l2")


def test_exact_topk_tensor_core_near_ties_fall_back_to_fp32(K):
    """Collapsed embeddings (pairwise cosine ~0.985) put many scores within the
    TF32 error bound: the certificate must fail and the fp32 re-run must still give the exact
    result; duplicates exercise the id tie-break."""
    rng = np.random.Generator(np.random.PCG64(5))
    base = rng.standard_normal((1, 64)).astype(np.float32)
    x = base + 0.02 * rng.standard_normal((6000, 64)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x[100:140] = x[99]                       # exact duplicates
    reruns = _tc_vs_fp32(K, x[:700].copy(), x, 10, "ip")
    assert reruns > 0
    _tc_vs_fp32(K, x[:700].copy(), x, 10, "l2")


def test_exact_topk_tensor_core_full_size_properties(K):
    """C3-sized all-item search: too large for the numpy oracle, so check size-independent
    properties -- every returned score is the fp32 score of the returned id, scores are sorted,
    the query itself is rank 0 for L2, and a random sample of rows equals the fp32 kernel."""
    from mre_b200 import _native as N
    from mre_b200 import synthetic as S
    x = S.spread_embeddings(62423, 128, seed=1).cuda().contiguous()
    st = {}
    s, i = K.topk(x, x, 10, N.METRIC_L2, precision="tf32", stats=st)
    assert (i[:, 0] == torch.arange(62423, device="cuda", dtype=torch.int32)).float().mean() > 0.999
    assert bool((s[:, 1:] >= s[:, :-1]).all())
    rows = torch.randperm(62423, generator=torch.Generator().manual_seed(0))[:2048].cuda()
    s0, i0 = K.topk(x[rows].contiguous(), x, 10, N.METRIC_L2, precision="fp32")
    assert torch.equal(i[rows], i0) and torch.equal(s[rows], s0)
    assert int(st["fp32_reruns"].item()) < 62423 // 10


# ---- tensor-core exhaustive Hamming search (pb200_hamming_topk_tc): equal to the popcount kernel ----
@pytest.mark.parametrize("nq,nx,code_bytes,k", [(300, 5000, 32, 10), (1, 700, 32, 10), (513, 20000, 32, 32),
                                                (260, 999, 4, 5), (1000, 3000, 64, 16), (300, 40, 16, 10),
                                                (257, 4097, 8, 17)])
def test_hamming_topk_tensor_core_equals_popcount_kernel(K, nq, nx, code_bytes, k):
    rng = np.random.Generator(np.random.PCG64(nq + nx))
    cx = torch.from_numpy(rng.integers(0, 256, (nx, code_bytes), dtype=np.uint8)).cuda()
    cq = torch.from_numpy(rng.integers(0, 256, (nq, code_bytes), dtype=np.uint8)).cuda()
    cq[: min(nq, nx) // 2] = cx[: min(nq, nx) // 2]          # exact matches and many ties
    d1, i1 = K.hamming_topk(cq, cx, k, id_offset=7, precision="tc")
    d0, i0 = K.hamming_topk(cq, cx, k, id_offset=7, precision="simt")
    np.testing.assert_array_equal(d1.cpu().numpy(), d0.cpu().numpy())
    np.testing.assert_array_equal(i1.cpu().numpy(), i0.cpu().numpy())
    # self search (shared expansion of the codes)
    d1, i1 = K.hamming_topk(cx, cx, min(k, 16), precision="tc")
    d0, i0 = K.hamming_topk(cx, cx, min(k, 16), precision="simt")
    assert torch.equal(d1, d0) and torch.equal(i1, i0)


# ---- tensor-core IVF search (pb200_ivf_search_tc): equal to the list-scan kernel -----------------
@pytest.mark.parametrize("n,d,nlist,nq,k", [(6000, 64, 37, 700, 10), (20000, 128, 100, 1000, 10),
                                            (3000, 32, 128, 300, 5), (9000, 128, 100, 513, 20)])
def test_ivf_search_tensor_core_equals_list_scan(K, n, d, nlist, nq, k):
    from mre_b200 import _native as N
    rng = np.random.Generator(np.random.PCG64(n + nq))
    cen = rng.standard_normal((nlist, d)).astype(np.float32)
    x = (cen[rng.integers(0, nlist, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    q = (x[rng.integers(0, n, nq)] + 0.05 * rng.standard_normal((nq, d))).astype(np.float32)
    xd, qd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(cen).cuda()
    _, a = K.topk(xd, cd, 1, N.METRIC_L2)
    lists = K.ivf_build(xd, a.view(-1).contiguous(), nlist)
    lay = K.ivf_tc_layout(*lists, nlist)
    assert lay is not None
    _, probes = K.topk(qd, cd, min(nlist, 20), N.METRIC_L2)
    st = {}
    d1, i1 = K.ivf_search_tc(qd, probes, *lists, lay, nlist, k, stats=st)
    d0, i0 = K.ivf_search(qd, probes, *lists, k)
    np.testing.assert_array_equal(i1.cpu().numpy(), i0.cpu().numpy())
    np.testing.assert_array_equal(d1.cpu().numpy(), d0.cpu().numpy())
    assert int(st["list_scan_reruns"].item()) < nq // 2


@pytest.mark.parametrize("n,d,nbits", [(5000, 128, 256), (300, 64, 64), (1000, 32, 512), (257, 100, 96)])
def test_lsh_encode_tensor_core_bit_identical(K, n, d, nbits):
    """pb200_lsh_encode_tc == pb200_lsh_encode bit for bit (projections near zero are recomputed in
    fp32); rows of zeros and tiny vectors exercise the all-uncertain case."""
    x = _data(n, d, 31)
    x[7] = 0.0
    x[11] *= 1e-20
    proj = _data(nbits, d, 32, normalise=False)
    xd, pd = torch.from_numpy(x).cuda(), torch.from_numpy(proj).cuda()
    c1 = K.lsh_encode(xd, pd, precision="tc")
    c0 = K.lsh_encode(xd, pd, precision="fp32")
    np.testing.assert_array_equal(c1.cpu().numpy(), c0.cpu().numpy())


# ---- N2: hit-rate / MRR through the rank kernel (SURVEY 8(f) "next" row) -------------------------
def test_hit_rate_mrr_vs_reference_golden(K):
    import json
    from mre_b200.utils.evaluation import calculate_hit_rate, calculate_mrr, evaluate_embeddings
    g = Hh.load("evaluation.npz")
    emb = torch.from_numpy(g["emb"])
    q, gt = g["pairs"][:, 0], g["pairs"][:, 1]
    ranks = K.rank_of_target(emb.cuda(), torch.from_numpy(q), torch.from_numpy(gt)).cpu().numpy()
    np.testing.assert_array_equal(ranks, g["ranks"])
    for k, want in zip(g["ks"], g["hit_rates"]):
        assert calculate_hit_rate(emb, q, gt, k=int(k)) == want
    assert abs(calculate_mrr(emb, q, gt) - float(g["mrr"])) < 1e-12
    assert abs(calculate_mrr(emb.cuda(), list(q), list(gt), scale=7) - float(g["mrr_scale7"])) < 1e-12
    ev = evaluate_embeddings(emb, {"positive_pairs": torch.from_numpy(g["pairs"])})
    want = json.loads(str(g["evaluate"]))
    assert set(ev) == set(want)
    for key in want:
        assert abs(ev[key] - want[key]) < 1e-12
    with pytest.raises(RuntimeError):
        calculate_hit_rate(emb, q, gt, k=701)


def test_rank_of_target_ties_edges_and_size(K):
    """Duplicates (tie rule: equal score, smaller index first), a pair whose target is the query,
    non-multiple-of-tile sizes; and a C3-sized run checked against the exact search kernel:
    rank <= 10  <=>  the target is in the query's top-10 (query not excluded)."""
    from mre_b200 import _native as N
    x = _data(1000, 48, 41)
    x[10] = x[3]; x[500] = x[3]
    q = np.array([3, 3, 3, 10, 999, 0, 77], dtype=np.int64)
    gt = np.array([3, 10, 500, 3, 0, 999, 77], dtype=np.int64)
    got = K.rank_of_target(torch.from_numpy(x).cuda(), torch.from_numpy(q), torch.from_numpy(gt)).cpu().numpy()
    np.testing.assert_array_equal(got, O.rank_of_target(x, q, gt))
    from mre_b200 import synthetic as S
    e = S.spread_embeddings(62423, 128, seed=1).cuda().contiguous()
    gen = torch.Generator().manual_seed(3)
    qq = torch.randint(0, 62423, (4096,), generator=gen)
    _s, ids = K.topk(e[qq.cuda()].contiguous(), e, 10, N.METRIC_IP, precision="fp32")
    tgt = ids[torch.arange(4096), torch.randint(0, 10, (4096,), generator=gen)].cpu()      # a top-10 member
    far = torch.randint(0, 62423, (4096,), generator=gen)
    r_in = K.rank_of_target(e, qq, tgt).cpu()
    r_far = K.rank_of_target(e, qq, far).cpu()
    assert bool((r_in <= 10).all())
    in_top = (ids.cpu() == far[:, None].to(torch.int32)).any(1)
    assert bool(((r_far <= 10) == in_top).all())


def test_indexes_leave_the_tensor_core_path_on_near_tied_data(K):
    """Collapsed embeddings fail the TF32 certificate for most queries: results
    stay exact (fp32 re-runs) and the index classes switch themselves to the fp32 kernels."""
    from mre_b200.utils.nearest_neighbors import FlatL2Index, WeakANDIndex
    rng = np.random.Generator(np.random.PCG64(17))
    base = rng.standard_normal((1, 64)).astype(np.float32)
    x = base + 0.004 * rng.standard_normal((5000, 64)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    flat = FlatL2Index(64); flat.add(x)
    d1, i1 = flat.search(x[:600], 10)
    assert flat.precision == "fp32"
    d2, i2 = flat.search(x[:600], 10)                     # now the CUDA-core path: same answer
    np.testing.assert_array_equal(i1, i2); np.testing.assert_array_equal(d1, d2)
    ivf = WeakANDIndex(64, 16, 10)
    ivf.build(x)
    a1 = ivf.search(x[:600], 10)
    a2 = ivf.search(x[:600], 10)
    np.testing.assert_array_equal(a1[1], a2[1]); np.testing.assert_array_equal(a1[0], a2[0])


# ---- tensor-core kernels against the ORACLE directly (the tcgen05 paths were only compared
# with this repo's own fp32 / popcount / list-scan kernels; the chain to the oracle was transitive) ----
@pytest.mark.parametrize("nq,nx,d,k,metric", [(300, 5000, 128, 10, "ip"), (512, 4096, 64, 10, "l2"),
                                              (257, 3001, 256, 16, "ip"), (1000, 2500, 32, 24, "l2")])
def test_exact_topk_tensor_core_vs_oracle(K, nq, nx, d, k, metric):
    from mre_b200 import _native as N
    import mre_b200.synthetic as S
    x = S.spread_embeddings(nx, d, seed=11, clusters=64, noise=0.3).numpy()
    q = S.spread_embeddings(nq, d, seed=12, clusters=64, noise=0.3).numpy()
    st = {}
    s, i = K.topk(torch.from_numpy(q), torch.from_numpy(x), k, N.METRIC_IP if metric == "ip" else N.METRIC_L2,
                  precision="tf32", stats=st)
    assert st["path"] == "tf32"
    s, i = s.cpu().numpy(), i.cpu().numpy()
    rs, ri = (O.exact_ip if metric == "ip" else O.exact_l2)(x, q, k)
    np.testing.assert_allclose(s, rs, atol=2e-5)                           # by score
    assert (i == ri).mean() > 0.98                                         # ids, up to fp32 near-ties
    for a, b in zip(i.tolist(), ri.tolist()):                              # top-k SETS match (north-star bar)
        assert len(set(a) & set(b)) >= k - 1


@pytest.mark.parametrize("nq,nx,nbits,k", [(300, 5000, 256, 10), (256, 2049, 128, 32), (700, 3000, 512, 10)])
def test_hamming_topk_tensor_core_vs_oracle(K, nq, nx, nbits, k):
    rng = np.random.Generator(np.random.PCG64(nbits + nq))
    cx = rng.integers(0, 256, size=(nx, nbits // 8), dtype=np.uint8)
    cq = np.concatenate([cx[:nq // 2], rng.integers(0, 256, size=(nq - nq // 2, nbits // 8), dtype=np.uint8)])
    dist, ids = K.hamming_topk(torch.from_numpy(cq).cuda(), torch.from_numpy(cx).cuda(), k, precision="tc")
    rd, ri = O.lsh_search_exhaustive(cx, cq, k)
    np.testing.assert_array_equal(dist.cpu().numpy(), rd)                  # Hamming ints: exact
    np.testing.assert_array_equal(ids.cpu().numpy(), ri)                   # (distance, id) total order


@pytest.mark.parametrize("n,d,nbits", [(3000, 128, 256), (513, 64, 128)])
def test_lsh_encode_tensor_core_vs_oracle(K, n, d, nbits):
    x = _data(n, d, 31)
    A = O.lsh_rotation(d, nbits)
    ref_codes, y = O.lsh_encode(x, A)
    codes = K.lsh_encode(torch.from_numpy(x), torch.from_numpy(A), precision="tc").cpu().numpy()
    diff_bits = np.unpackbits(codes ^ ref_codes, axis=1, bitorder="little").astype(bool)
    assert diff_bits.mean() < 1e-4 and np.all(np.abs(y[diff_bits]) < 1e-6)  # north-star bar for the codes


def test_ivf_search_tensor_core_vs_oracle(K):
    from mre_b200 import _native as N
    import mre_b200.synthetic as S
    n, d, nlist, nq, k = 20000, 64, 100, 600, 10
    x = S.spread_embeddings(n, d, seed=4, clusters=200, noise=0.25).numpy()
    cent = O.kmeans(x, nlist, niter=4)
    xt, ct = torch.from_numpy(x).cuda(), torch.from_numpy(cent).cuda()
    _, a = K.topk(xt, ct, 1, N.METRIC_L2)
    assign = a.view(-1).contiguous()
    lists = K.ivf_build(xt, assign, nlist)
    lay = K.ivf_tc_layout(*lists, nlist)
    q = xt[:nq].contiguous()
    _, probes = K.topk(q, ct, 20, N.METRIC_L2)
    assert K.ivf_search_tc_supported(nq, lay[0].size(0), d, k, nlist)
    dist, ids = K.ivf_search_tc(q, probes, *lists, lay, nlist, k)
    rd, ri = O.ivf_search(x, cent, assign.cpu().numpy(), x[:nq], k, 20)
    np.testing.assert_allclose(dist.cpu().numpy(), rd, atol=2e-5)
    assert (ids.cpu().numpy() == ri).mean() > 0.98


def test_ivf_and_lsh_tables_accept_k_above_32(K):
    """faiss accepts any k; the list-scan and table-probe kernels hold 32 results per pass, larger k runs
    in passes floored by the previous pass's last result under the (score, id) order."""
    from mre_b200 import _native as N
    from mre_b200.utils.nearest_neighbors import WeakANDIndex, LSHIndex
    import mre_b200.synthetic as S
    n, d, nlist, k = 5000, 32, 16, 75
    x = S.spread_embeddings(n, d, seed=6, clusters=40, noise=0.3).numpy()
    cent = O.kmeans(x, nlist, niter=4)
    idx = WeakANDIndex(d, nlist, centroids=cent)
    idx.build(x)
    dist, ids = idx.search(x[:300], k=k)
    rd, ri = O.ivf_search(x, cent, idx.assign.cpu().numpy(), x[:300], k, min(nlist, 20))
    np.testing.assert_allclose(dist, rd, atol=2e-5)
    assert (ids == ri).mean() > 0.98
    assert np.all(np.diff(dist, axis=1)[np.isfinite(dist[:, 1:])] >= 0)
    for rerank in ("hamming", "dot"):
        lt = LSHIndex(d, 64, 8, mode="tables", rerank=rerank)
        lt.build(x)
        sc, li = lt.search(x[:200], k=k)
        codes = lt.codes.cpu().numpy()
        rs, rli, _nc = O.lsh_search_tables(codes, codes[:200], k, 8, vectors=x if rerank == "dot" else None,
                                           queries=x[:200])
        ok = rli >= 0
        np.testing.assert_allclose(sc[ok], rs[ok], atol=2e-5)
        assert (li == rli).mean() > 0.98 and ((li == -1) == (rli == -1)).all()
        assert ok.sum(1).max() > 32                                   # the passes beyond the first were exercised


def test_kmeans_resplits_empty_clusters_on_collapsed_data(K):
    """on near-duplicate embeddings plain Lloyd leaves lists empty for good (an empty list kept its
    stale centroid); with faiss's split rule empty lists are re-seeded from populated ones."""
    from mre_b200.utils.nearest_neighbors import train_kmeans
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 32, generator=g)
    x = torch.nn.functional.normalize(base + 0.01 * torch.randn(4000, 32, generator=g), dim=1)
    x[:1000] = x[0]                                               # a quarter of the points are exact duplicates
    x = x.cuda()
    init_dupes = x[:64].clone()                                   # 64 identical initial centroids: 63 lists start empty

    def used(split):
        import mre_b200.utils.nearest_neighbors as NN
        cent = init_dupes.clone()
        for _ in range(12):
            _, a = K.topk(x, cent, 1, 1)
            offsets, _ids, vecs = K.ivf_build(x, a.view(-1).contiguous(), 64)
            K.ivf_centroid_update(vecs, offsets, cent)
            if split:
                NN._split_empty_clusters(cent, offsets, torch.Generator().manual_seed(1))
        _, a = K.topk(x, cent, 1, 1)
        assert torch.isfinite(cent).all()
        return torch.unique(a).numel()
    without, with_split = used(False), used(True)
    assert without <= 8 and with_split >= 16 and with_split > without, (without, with_split)
    cent = train_kmeans(x, 64, niter=10)                          # the public entry point runs with the split
    assert torch.isfinite(cent).all() and torch.unique(K.topk(x, cent, 1, 1)[1]).numel() >= 16
