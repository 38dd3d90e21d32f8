"""GPU parity at BASELINE.json's full size (config C2: 62,423 items / 162,541 users /
25,000,095 ratings; F=128, H=256, E=128, 2 layers): every start item's ids / counts / weights
bit-exact against the C oracle, embeddings of ALL items within 1e-3 relative of the float64
oracle forward, plus size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import mre_b200  # noqa: F401
    from mre_b200 import synthetic as S
    from mre_b200.utils.random_walk import RandomWalkSampler
    M, U, R, F_, Hd, E_, layers = S.CONFIGS["C2"]
    ei, w = S.bipartite_graph(M, U, R, seed=0)
    assert ei.shape == (2, 2 * R)
    sampler = RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100, seed=1234,
                                num_nodes=M + U)
    return dict(M=M, U=U, ei=ei, w=w, sampler=sampler, dims=(F_, Hd, E_, layers))


def test_c2_csr_and_walks_bit_exact(c2):
    from mre_b200 import kernels as K
    M, U = c2["M"], c2["U"]
    csr = c2["sampler"].csr
    assert csr.cum_kind == 0 and csr.quant_shift == 1            # ratings are multiples of 0.5
    row_ptr, col, cum = O.c_csr_build(c2["ei"], c2["w"], M + U, 1)
    np.testing.assert_array_equal(csr.row_ptr.cpu().numpy(), row_ptr)
    np.testing.assert_array_equal(csr.col.cpu().numpy(), col)
    np.testing.assert_array_equal(csr.cum.cpu().numpy().view(np.uint32), cum)
    c2["oracle_csr"] = (row_ptr, col, cum)
    c2["lists"] = []
    for epoch in range(2):
        ids, counts, w32, nvalid = K.walk_topt(csr, torch.arange(M), 100, 2, 10, 1234, epoch)
        o = O.c_walk_topt(row_ptr, col, cum, np.arange(M), 100, 2, 10, 1234, epoch)
        np.testing.assert_array_equal(ids.cpu().numpy(), o["ids"])
        np.testing.assert_array_equal(counts.cpu().numpy(), o["counts"])
        np.testing.assert_array_equal(w32.cpu().numpy(), o["w32"])
        np.testing.assert_array_equal(nvalid.cpu().numpy(), o["nvalid"])
        c2["lists"].append(o)
    # size-independent properties: weights sum to 1, counts sorted, counts <= W*L
    w = w32.cpu().numpy(); c = counts.cpu().numpy(); nv = nvalid.cpu().numpy()
    np.testing.assert_allclose(w.sum(1)[nv > 0], 1.0, atol=1e-6)
    assert (np.diff(c, axis=1) <= 0).all() and c.sum(1).max() <= 200
    # user ids dominate the bipartite neighbourhoods
    assert (o["ids"][o["ids"] >= 0] >= M).mean() > 0.5


def test_c2_embeddings_all_items(c2):
    from mre_b200.model.pinsage import PinSage
    from mre_b200 import synthetic as S, neighbor_lists as NL
    if "lists" not in c2:
        pytest.skip("walk test did not run")
    M = c2["M"]
    F_, Hd, E_, layers = c2["dims"]
    torch.manual_seed(0)
    model = PinSage(F_, Hd, E_, layers).cuda().eval()
    x = S.features(M, F_)
    c2["sampler"].epoch = 0
    emb = model.get_embeddings(x.cuda(), c2["sampler"], 10).cpu().numpy()
    nbrs = [[o["ids"][r, :o["nvalid"][r]].tolist() for r in range(M)] for o in c2["lists"]]
    wts = [[o["w64"][r, :o["nvalid"][r]].tolist() for r in range(M)] for o in c2["lists"]]
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    ref = O.pinsage_forward(x.numpy(), sd, layers, nbrs, wts)
    err = Hh.rel_row_err(emb, ref)
    assert err < 1e-3, err
    np.testing.assert_allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)
    # pinned-host in / pinned-host out public path gives the same bits as the device path
    c2["sampler"].epoch = 0
    out = torch.empty(M, E_).pin_memory()
    model.get_embeddings(x.pin_memory(), c2["sampler"], 10, out=out)
    np.testing.assert_array_equal(out.numpy(), emb)
