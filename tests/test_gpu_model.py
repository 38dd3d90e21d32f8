"""GPU parity: P1-P5 pooling, G1-G4 conv / forward / get_embeddings through the drop-in
classes (which call the C ABI) vs golden outputs of the unmodified reference and the oracle.
Floating-point bar: max row-relative L2 error <= 1e-3 (fp32 path)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu
TOL = 1e-3     # north-star: embeddings within 1e-3 relative (fp32)


@pytest.fixture(scope="module")
def dev():
    import mre_b200  # noqa: F401
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_pooling_variants_vs_reference_golden(dev):
    from mre_b200.model.pinsage import ImportancePooling
    from mre_b200.model.layers import ImportancePoolingLayer, WeightedMeanPoolingLayer, MaxPoolingLayer
    from mre_b200.model.aggregators import WeightedAggregator, MeanAggregator, ImportanceAggregator
    g = Hh.load("pooling.npz")
    x = torch.from_numpy(g["x"]).to(dev)
    nbrs, wts, nb_ok, wt_ok = Hh.lists_from_json(g, "nbrs", "wts", "nb_ok", "wt_ok")
    tol = dict(rtol=1e-5, atol=1e-6)
    chk = lambda got, key: np.testing.assert_allclose(got.cpu().numpy(), g[key], **tol)
    chk(ImportancePooling()(x, nbrs, wts), "pinsage")
    nb_int, wt_int = list(nbrs), list(wts)
    nb_int[5], wt_int[5] = 7, 0.3
    nb_int[14], wt_int[14] = np.int64(x.size(0) + 3), 1.0
    chk(ImportancePooling()(x, nb_int, wt_int), "pinsage_bareint")
    chk(ImportancePoolingLayer()(x, nbrs, wts), "layers_importance")
    chk(WeightedMeanPoolingLayer()(x, nbrs, wts), "layers_wmean")
    chk(WeightedMeanPoolingLayer()(x, nbrs, None), "layers_wmean_none")
    chk(MaxPoolingLayer()(x, nbrs), "layers_max")
    chk(WeightedAggregator()(x, nb_ok, wt_ok), "agg_weighted")
    chk(MeanAggregator()(x, nb_ok), "agg_mean")
    ia = ImportanceAggregator(x.size(1), 16).to(dev)
    with torch.no_grad():
        ia.transform.weight.copy_(torch.from_numpy(g["ia_W"])); ia.transform.bias.copy_(torch.from_numpy(g["ia_b"]))
        ia.norm.weight.copy_(torch.from_numpy(g["ia_gamma"])); ia.norm.bias.copy_(torch.from_numpy(g["ia_beta"]))
    np.testing.assert_allclose(ia(x, nb_ok, wt_ok).cpu().numpy(), g["agg_importance"], rtol=1e-4, atol=1e-5)
    with pytest.raises(IndexError):
        WeightedAggregator()(x, [[0, x.size(0)]], [[0.5, 0.5]])
    # CPU input -> result comes back on the CPU (computed on the GPU; there is no CPU path)
    out_cpu = ImportancePooling()(x.cpu(), nbrs, wts)
    assert out_cpu.device.type == "cpu"
    np.testing.assert_allclose(out_cpu.numpy(), g["pinsage"], **tol)


@pytest.mark.parametrize("dim,T", [(1, 1), (3, 5), (130, 40), (256, 10), (512, 64)])
def test_pool_shapes_vs_oracle(dev, dim, T):
    from mre_b200.model.pinsage import ImportancePooling
    from mre_b200.model.layers import ImportancePoolingLayer, MaxPoolingLayer
    rng = np.random.Generator(np.random.PCG64(dim * 100 + T))
    M, n = 77, 150
    x = rng.standard_normal((M, dim)).astype(np.float32)
    nbrs = [rng.integers(0, M + 20, size=int(rng.integers(0, T + 1))).tolist() for _ in range(n)]
    wts = [(rng.integers(0, 9, size=len(l)) / 4.0).tolist() for l in nbrs]
    xt = torch.from_numpy(x).to(dev)
    tol = dict(rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ImportancePooling()(xt, nbrs, wts).cpu().numpy(), O.pool_pinsage(x, nbrs, wts), **tol)
    np.testing.assert_allclose(ImportancePoolingLayer()(xt, nbrs, wts).cpu().numpy(),
                               O.pool_layers(x, nbrs, wts, "importance"), **tol)
    np.testing.assert_allclose(MaxPoolingLayer()(xt, nbrs).cpu().numpy(), O.pool_layers(x, nbrs, None, "max"), **tol)


def _load_model(g, dev):
    from mre_b200.model.pinsage import PinSage
    F_, Hd, E_, layers = (int(v) for v in g["dims"])
    model = PinSage(F_, Hd, E_, layers)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    model.load_state_dict(sd)                      # reference key names load unchanged
    return model.to(dev).eval(), layers


@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_forward_vs_reference_golden(dev, precision):
    from mre_b200 import _native as N
    g = Hh.load("forward.npz")
    model, layers = _load_model(g, dev)
    model.precision = N.PRECISIONS[precision]
    x = torch.from_numpy(g["x"]).to(dev)
    assert Hh.rel_row_err(model(x).cpu().numpy(), g["emb_mlp"]) < TOL                 # MLP branch
    nb0, wt0, nb1, wt1 = Hh.lists_from_json(g, "nb0", "wt0", "nb1", "wt1")
    for fold, fuse in ((True, False), (True, True), (False, True)):
        model.fold, model.fuse_pool = fold, fuse
        got = model(x, None, [nb0, nb1], [wt0, wt1]).cpu().numpy()
        assert Hh.rel_row_err(got, g["emb_lists"]) < TOL, f"fold={fold} fuse_pool={fuse}"
    model.fold, model.fuse_pool = True, False
    # CPU input in, CPU tensor out
    assert model(torch.from_numpy(g["x"])).device.type == "cpu"


def test_get_embeddings_vs_reference_golden(dev):
    """G4 end to end: device CSR + 2x walk kernel (epoch = layer) + fused conv layers, compared
    with the unmodified reference's get_embeddings driven by the same uniform stream."""
    from mre_b200.utils.random_walk import RandomWalkSampler
    g = Hh.load("forward.npz")
    model, _ = _load_model(g, dev)
    sampler = RandomWalkSampler(torch.from_numpy(g["edge_index"]), torch.from_numpy(g["edge_weights"]),
                                int(g["L"]), int(g["W"]), seed=int(g["seed"]), device=dev)
    x = torch.from_numpy(g["x"]).to(dev)
    got = model.get_embeddings(x, sampler, int(g["T"])).cpu().numpy()
    assert Hh.rel_row_err(got, g["emb_full"]) < TOL
    # list API of a foreign sampler object goes through the same kernels
    sampler.epoch = 0

    class ListOnly:
        def batch_sample_neighbors(self, nodes, T):
            return sampler.batch_sample_neighbors(nodes, T)
    got2 = model.get_embeddings(x, ListOnly(), int(g["T"])).cpu().numpy()
    np.testing.assert_allclose(got2, got, rtol=1e-6, atol=1e-7)


def test_graph_conv_layer_vs_reference_golden(dev):
    from mre_b200.model.layers import GraphConvLayer
    g = Hh.load("forward.npz")
    gsd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("gcl.")}
    layer = GraphConvLayer(g["gx"].shape[1], g["g_out"].shape[1])
    layer.load_state_dict(gsd, strict=False)
    layer = layer.to(dev).eval()
    gx, gn = torch.from_numpy(g["gx"]).to(dev), torch.from_numpy(g["gn"]).to(dev)
    assert Hh.rel_row_err(layer(gx, gn).cpu().numpy(), g["g_out"]) < TOL
    assert Hh.rel_row_err(layer(gx[:1], gn[:1]).cpu().numpy(), g["g_out1"]) < TOL    # BN skipped


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("F_,Hd,E_,T", [(64, 64, 64, 10), (128, 256, 128, 10), (20, 100, 36, 7), (8, 300, 5, 3),
                                         (30, 50, 7, 40)])
def test_forward_dims_vs_oracle(dev, F_, Hd, E_, T, precision):
    """C1 / C2 layer widths plus odd sizes (non multiples of 4, > 256 columns)."""
    from mre_b200.model.pinsage import PinSage
    from mre_b200 import _native as N
    torch.manual_seed(F_ + Hd)
    M = 333
    model = PinSage(F_, Hd, E_, 2).to(dev).eval()
    model.precision = N.PRECISIONS[precision]
    x = torch.randn(M, F_)
    rng = np.random.Generator(np.random.PCG64(1))
    nbrs = [[rng.integers(0, M + 50, size=int(rng.integers(0, T + 1))).tolist() for _ in range(M)] for _ in range(2)]
    wts = [[(rng.integers(1, 30, size=len(l)) / 10.0).tolist() for l in layer] for layer in nbrs]
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    ref = O.pinsage_forward(x.numpy(), sd, 2, nbrs, wts)
    got = model(x.to(dev), None, nbrs, wts).cpu().numpy()
    assert Hh.rel_row_err(got, ref) < TOL
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("n,k1,k2,nout,pooled,flags", [
    (128, 32, 0, 16, False, 0), (300, 128, 0, 256, False, 1), (1000, 256, 256, 256, True, 3),
    (777, 256, 0, 128, False, 2), (200, 64, 64, 64, True, 3), (515, 20, 12, 40, False, 1),
    (1, 8, 0, 8, False, 3), (129, 4, 4, 250, True, 0), (4000, 512, 0, 256, False, 3)])
def test_tcgen05_dense_vs_cuda_core_fp32(dev, n, k1, k2, nout, pooled, flags):
    """The tcgen05 kind::tf32 kernel against the exact-fp32 CUDA-core kernel of the same op:
    TF32 operand rounding only (<= 1e-3 row-relative, the fp32 bar of the north star)."""
    from mre_b200 import kernels as K, _native as N
    g = torch.Generator(device="cpu").manual_seed(n + k1 + nout)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
    a1, w, b = rnd(n, k1), rnd(nout, k1 + k2) / (k1 + k2) ** 0.5, rnd(nout)
    kw = {}
    if k2 and pooled:
        T = 10
        ids = torch.randint(0, n + 50, (n, T), generator=g, dtype=torch.int32).to(dev)
        kw = dict(pool_x=rnd(n, k2), lists=(ids, torch.rand(n, T, generator=g).to(dev),
                                            torch.randint(0, T + 1, (n,), generator=g, dtype=torch.int32).to(dev), None))
    elif k2:
        kw = dict(a2=rnd(n, k2))
    ref = K.gather_dense(a1, w, b, flags=flags, precision=N.PREC_FP32, **kw).cpu().numpy()
    got = K.gather_dense(a1, w, b, flags=flags, precision=N.PREC_TF32, **kw).cpu().numpy()
    assert Hh.rel_row_err(got, ref) < 1e-3
    # shapes the tensor-core path does not cover are refused loudly, and AUTO routes around them
    with pytest.raises(N.NativeError, match="PB200_PREC_TF32"):
        K.gather_dense(rnd(8, 6), rnd(300, 6), None, precision=N.PREC_TF32)
    assert K.gather_dense(rnd(8, 6), rnd(300, 6), None, precision=N.PREC_AUTO).shape == (8, 300)


@pytest.mark.parametrize("fill", ["empty", "ragged", "all_valid"])
def test_tcgen05_dense_many_tiles_per_cta(dev, fill):
    """Persistent kernel with several 128-row tiles per CTA (> 148 tiles) and extreme list
    shapes: all-empty tiles make the pool warps race ahead of the MMA ring (regression test for a
    barrier-phase bug that deadlocked), all-valid lists exercise the > 2 neighbour path."""
    from mre_b200 import kernels as K, _native as N
    M, T = 148 * 128 * 2 + 77, 10
    g = torch.Generator().manual_seed(7)
    h = torch.randn(M, 64, generator=g).to(dev)
    w, b = (torch.randn(48, 128, generator=g) / 11).to(dev), torch.randn(48, generator=g).to(dev)
    ids = torch.randint(0, 3 * M, (M, T), generator=g, dtype=torch.int32).to(dev)
    wt = torch.rand(M, T, generator=g).to(dev)
    ll = torch.randint(0, T + 1, (M,), generator=g, dtype=torch.int32).to(dev)
    if fill == "empty":
        ll.zero_()
    elif fill == "all_valid":
        ids = ids % M
    lists = (ids, wt, ll, None)
    ref = K.gather_dense(h, w, b, pool_x=h, lists=lists, flags=3, precision=N.PREC_FP32).cpu().numpy()
    for flags in (3, 3 | N.EPI_ROUND_TF32 | N.IN_A1_TF32):
        hh = h
        if flags & N.IN_A1_TF32:
            hh = h.clone(); K.lib().pb200_round_tf32(K.ptr(hh), K.ptr(hh), hh.numel(), None)
        got = K.gather_dense(hh, w, b, pool_x=hh, lists=lists, flags=flags, precision=N.PREC_TF32).cpu().numpy()
        assert Hh.rel_row_err(got, ref) < 1.5e-3


def test_graphed_embeddings_replays_equal_eager_calls(dev):
    """graphs.GraphedEmbeddings: replay k of the captured step == eager get_embeddings call k
    (the sampling epoch advances on the device inside the graph)."""
    from mre_b200 import synthetic as S
    from mre_b200.graphs import GraphedEmbeddings
    from mre_b200.model.pinsage import PinSage
    from mre_b200.utils.random_walk import RandomWalkSampler
    M, U, R = 1500, 4000, 60000
    ei, w = S.bipartite_graph(M, U, R, seed=3)
    x = S.features(M, 64).to(dev)
    torch.manual_seed(0)
    model = PinSage(64, 128, 64, 2).to(dev).eval()
    sampler = RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100, seed=7, device=dev, num_nodes=M + U)
    sampler.epoch = 40
    g = GraphedEmbeddings(model, x, sampler, 10)
    outs = [g.replay().clone() for _ in range(3)]
    assert not torch.equal(outs[0], outs[1])                 # fresh walks every replay
    for k, got in enumerate(outs):
        sampler.epoch = 40 + 2 * k
        want = model.get_embeddings(x, sampler, 10)
        assert torch.equal(got, want)


def test_pooling_negative_ids_follow_python_indexing_like_the_reference(dev):
    """negative neighbour ids index from the end in every reference class (x[list]); ids below
    -M raise IndexError.  Golden: tests/golden/pooling_negative.npz from the unmodified reference."""
    import json
    from mre_b200.model.pinsage import ImportancePooling
    from mre_b200.model.layers import ImportancePoolingLayer, WeightedMeanPoolingLayer, MaxPoolingLayer
    from mre_b200.model.aggregators import WeightedAggregator, MeanAggregator, ImportanceAggregator
    g = Hh.load("pooling_negative.npz")
    meta = json.loads(str(g["lists"]))
    x = torch.from_numpy(g["x"]).to(dev)
    M = x.size(0)
    nbrs, wts, nb_ok, wt_ok = meta["nbrs"], meta["wts"], meta["nb_ok"], meta["wt_ok"]
    assert any(v < 0 for l in nbrs for v in l)
    tol = dict(rtol=1e-5, atol=1e-6)
    chk = lambda got, key: np.testing.assert_allclose(got.cpu().numpy(), g[key], **tol)
    chk(ImportancePooling()(x, nbrs, wts), "pinsage")
    chk(ImportancePoolingLayer()(x, nbrs, wts), "layers_importance")
    chk(WeightedMeanPoolingLayer()(x, nbrs, wts), "layers_wmean")
    chk(MaxPoolingLayer()(x, nbrs), "layers_max")
    chk(WeightedAggregator()(x, nb_ok, wt_ok), "agg_weighted")
    chk(MeanAggregator()(x, nb_ok), "agg_mean")
    ia = ImportanceAggregator(x.size(1), 8).to(dev)
    with torch.no_grad():
        ia.transform.weight.copy_(torch.from_numpy(g["ia_W"])); ia.transform.bias.copy_(torch.from_numpy(g["ia_b"]))
        ia.norm.weight.copy_(torch.from_numpy(g["ia_gamma"])); ia.norm.bias.copy_(torch.from_numpy(g["ia_beta"]))
    np.testing.assert_allclose(ia(x, nb_ok, wt_ok).cpu().numpy(), g["agg_importance"], rtol=1e-4, atol=1e-5)
    assert set(meta["raises"].values()) == {"IndexError"}          # what the reference does below -M ...
    bad, bw = [[1, -M - 3]], [[0.5, 0.5]]
    for call in (lambda: ImportancePooling()(x, bad, bw), lambda: ImportancePoolingLayer()(x, bad, bw),
                 lambda: WeightedMeanPoolingLayer()(x, bad, bw), lambda: MaxPoolingLayer()(x, bad),
                 lambda: WeightedAggregator()(x, bad, bw), lambda: MeanAggregator()(x, bad), lambda: ia(x, bad, bw)):
        with pytest.raises(IndexError):                             # ... and what the drop-ins do
            call()


def test_reference_checkpoint_loads_and_matches_reference_outputs(dev):
    """checkpoints/best_model.pt (committed as a state_dict fixture): loads by the reference's key names;
    MLP branch and importance branch equal the unmodified reference's outputs (fp32 path <= 1e-3, and the
    TF32 tensor-core path within the same bar)."""
    import json
    from mre_b200 import _native as N
    g = Hh.load("checkpoint.npz")
    model, layers = _load_model(g, dev)
    assert tuple(int(v) for v in g["dims"]) == (128, 256, 128, 2)
    x = torch.from_numpy(g["x"]).to(dev)
    meta = json.loads(str(g["lists"]))
    nb = [meta[f"nb{i}"] for i in range(layers)]
    wt = [meta[f"wt{i}"] for i in range(layers)]
    for prec in (N.PREC_FP32, N.PREC_AUTO):
        model.precision = prec
        assert Hh.rel_row_err(model(x).cpu().numpy(), g["emb_mlp"]) < TOL
        assert Hh.rel_row_err(model(x, None, nb, wt).cpu().numpy(), g["emb_imp"]) < TOL


def test_graph_conv_layer_training_mode_matches_the_reference(dev):
    """A freshly constructed reference GraphConvLayer is in TRAINING mode: batch statistics, running statistics
    updated (momentum 0.1, unbiased variance), single-row batches skip BatchNorm.  Golden from the unmodified
    reference (tests/golden/graphconv_train.npz)."""
    from mre_b200.model.layers import GraphConvLayer
    g = Hh.load("graphconv_train.npz")
    layer = GraphConvLayer(20, 12)
    layer.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd0.")})
    layer = layer.to(dev)
    assert layer.training
    gx, gn = torch.from_numpy(g["gx"]).to(dev), torch.from_numpy(g["gn"]).to(dev)
    assert Hh.rel_row_err(layer(gx, gn).cpu().numpy(), g["out1"]) < 1e-4
    assert Hh.rel_row_err(layer(gx * 0.5, gn + 1.0).cpu().numpy(), g["out2"]) < 1e-4
    assert Hh.rel_row_err(layer(gx[:1], gn[:1]).cpu().numpy(), g["out_single"]) < 1e-4
    sd = layer.state_dict()
    for k in ("bn.running_mean", "bn.running_var"):
        np.testing.assert_allclose(sd[k].cpu().numpy(), g["sd1." + k], rtol=1e-4, atol=1e-6)
    assert int(sd["bn.num_batches_tracked"]) == int(g["sd1.bn.num_batches_tracked"]) == 2
    layer.eval()                                                   # eval afterwards uses the moved running statistics
    out_eval = layer(gx, gn)
    assert out_eval.shape == (33, 12)
