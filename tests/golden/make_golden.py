"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed .npz files.

Fixtures (all small, seeded):
  walk_quant.npz   S0-S3, ratings-weighted bipartite graph + dead ends / duplicates / self loop
  walk_unit.npz    S0-S3, edge_weights=None, L=3, T > #distinct
  walk_float.npz   S0-S3, arbitrary float32 weights (float64 prefix rule)
  pooling.npz      P1-P5 on ragged lists (empty, out-of-range, zero-sum, bare int)
  forward.npz      G1, G2, G4 (importance + MLP branches, get_embeddings), G3
  exact.npz        E1 generate_recommendations
  evaluation.npz   N2 calculate_hit_rate / calculate_mrr / evaluate_embeddings (8(f) "next" row)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                      # noqa: E402
from oracle import ref_harness as H                 # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rw, ps, ly, ag, ev = H.import_reference()


def small_graph(M, U, R, seed, extras=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    pairs = np.unique(rng.integers(0, U, size=2 * R) * M + (rng.zipf(1.6, size=2 * R) % M))
    rng.shuffle(pairs)
    pairs = pairs[:R]
    u, i = pairs // M + M, pairs % M
    r = (0.5 * rng.integers(1, 11, size=len(pairs))).astype(np.float32)
    src, dst, w = [u, i], [i, u], [r, r]
    N = M + U
    if extras:
        sink, loop, iso = N, N + 1, N + 2        # dead end, self loop only, isolated (max id)
        src.append(np.array([0, 1, 2, 0, loop, 5, 5, 5, N + 3], dtype=np.int64))
        dst.append(np.array([sink, sink, sink, sink, loop, 7, 7, 7, iso + 1], dtype=np.int64))
        w.append(np.array([5.0, 0.5, 2.5, 5.0, 1.0, 3.0, 3.0, 0.5, 1.0], dtype=np.float32))
    ei = np.stack([np.concatenate(src), np.concatenate(dst)]).astype(np.int64)
    return ei, np.concatenate(w).astype(np.float32)


def pad(lists, T, fill, dtype):
    out = np.full((len(lists), T), fill, dtype=dtype)
    for r, l in enumerate(lists):
        out[r, :len(l)] = l
    return out


def run_walk_case(name, ei, w, W, L, T, seed, epoch, starts):
    t_ei = torch.from_numpy(ei)
    t_w = None if w is None else torch.from_numpy(w)
    sampler = rw.RandomWalkSampler(t_ei, t_w, walk_length=L, num_walks=W)

    def u_fn(start, walk, step):
        return int(O.walk_uniform53(seed, epoch, start, walk, step)) / 9007199254740992.0

    inj = H.UniformInjector(sampler, u_fn)
    nbrs, wts = inj.batch_sample(starts, T)
    for n_, w_ in zip(nbrs, wts):
        assert all(isinstance(x, (int, np.integer)) for x in n_) and len(n_) == len(w_)
    np.savez_compressed(
        os.path.join(OUT, name), edge_index=ei, edge_weights=np.zeros(0, np.float32) if w is None else w,
        has_weights=np.array(w is not None), starts=np.asarray(starts, np.int64),
        W=W, L=L, T=T, seed=seed, epoch=epoch,
        ids=pad(nbrs, T, -1, np.int64), weights=pad(wts, T, 0.0, np.float64),
        nvalid=np.array([len(n_) for n_ in nbrs], np.int32))
    print(f"{name}: {len(starts)} starts, mean nvalid {np.mean([len(n_) for n_ in nbrs]):.2f}")
    return sampler


def make_walks():
    assert H.check_choice_rule(), "installed numpy's choice() does not follow the documented rule"
    ei, w = small_graph(120, 300, 3000, seed=11)
    N = int(ei.max()) + 1
    run_walk_case("walk_quant.npz", ei, w, W=100, L=2, T=10, seed=1234, epoch=0,
                  starts=list(range(N)))
    run_walk_case("walk_unit.npz", ei, None, W=20, L=3, T=50, seed=99, epoch=3,
                  starts=list(range(0, N, 3)))
    rng = np.random.Generator(np.random.PCG64(5))
    wf = rng.random(ei.shape[1]).astype(np.float32) * 3.0 + 0.01
    run_walk_case("walk_float.npz", ei, wf, W=50, L=2, T=10, seed=2**40 + 17, epoch=1,
                  starts=list(range(0, N, 2)))


def ragged_lists(n, M, T, rng, out_of_range=True):
    nbrs, wts = [], []
    for r in range(n):
        kind = r % 9
        k = int(rng.integers(1, T + 1))
        ids = rng.integers(0, M, size=k).tolist()
        ws = (rng.integers(1, 20, size=k) / 7.0).tolist()
        if kind == 0:
            ids, ws = [], []
        elif kind == 1 and out_of_range:
            ids[0] = M + int(rng.integers(0, 50))           # dropped id (first position)
        elif kind == 2 and out_of_range:
            ids = [M + j for j in range(k)]                 # nothing valid
        elif kind == 3:
            ws = [0.0] * k                                  # zero-sum weights
        elif kind == 4 and out_of_range:
            ids[-1] = M                                     # boundary id == x.size(0)
        elif kind == 5:
            ids = [np.int64(v) for v in ids]                # numpy ints, as the sampler emits
        elif kind == 6:
            ids[0] = M - 1                                  # boundary id == max valid
        nbrs.append(ids)
        wts.append(ws)
    return nbrs, wts


def make_pooling():
    rng = np.random.Generator(np.random.PCG64(21))
    M, Hd, T, n = 50, 24, 10, 90
    x = torch.randn(M, Hd, generator=torch.Generator().manual_seed(3))
    nbrs, wts = ragged_lists(n, M, T, rng)
    out = {}
    out["pinsage"] = ps.ImportancePooling()(x, nbrs, wts).numpy()
    # bare-int entries (pinsage.py:110-112)
    nb_int = list(nbrs); wt_int = list(wts)
    nb_int[5], wt_int[5] = 7, 0.3
    nb_int[14], wt_int[14] = np.int64(M + 3), 1.0
    out["pinsage_bareint"] = ps.ImportancePooling()(x, nb_int, wt_int).numpy()
    out["layers_importance"] = ly.ImportancePoolingLayer()(x, nbrs, wts).numpy()
    out["layers_wmean"] = ly.WeightedMeanPoolingLayer()(x, nbrs, wts).numpy()
    out["layers_wmean_none"] = ly.WeightedMeanPoolingLayer()(x, nbrs, None).numpy()
    out["layers_max"] = ly.MaxPoolingLayer()(x, nbrs).numpy()
    nb_ok, wt_ok = ragged_lists(n, M, T, rng, out_of_range=False)
    out["agg_weighted"] = ag.WeightedAggregator()(x, nb_ok, wt_ok).numpy()
    out["agg_mean"] = ag.MeanAggregator()(x, nb_ok).numpy()
    torch.manual_seed(5)
    ia = ag.ImportanceAggregator(Hd, 16)
    with torch.no_grad():
        ia.norm.weight.uniform_(0.5, 1.5); ia.norm.bias.uniform_(-0.2, 0.2)
        out["agg_importance"] = ia(x, nb_ok, wt_ok).numpy()
    meta = dict(nbrs=[[int(v) for v in l] for l in nbrs], wts=wts,
                nb_ok=[[int(v) for v in l] for l in nb_ok], wt_ok=wt_ok)
    np.savez_compressed(os.path.join(OUT, "pooling.npz"), x=x.numpy(), lists=json.dumps(meta),
                        ia_W=ia.transform.weight.detach().numpy(), ia_b=ia.transform.bias.detach().numpy(),
                        ia_gamma=ia.norm.weight.detach().numpy(), ia_beta=ia.norm.bias.detach().numpy(),
                        **out)
    print("pooling.npz:", {k: v.shape for k, v in out.items()})


def make_forward():
    F_, Hd, E_, layers, M, U = 16, 32, 16, 2, 120, 300
    torch.manual_seed(0)
    model = ps.PinSage(F_, Hd, E_, num_layers=layers).eval()
    x = torch.randn(M, F_, generator=torch.Generator().manual_seed(1))
    ei, w = small_graph(M, U, 3000, seed=11, extras=False)
    sampler = rw.RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100)
    seed = 4321
    call = {"epoch": 0}

    def u_fn(start, walk, step):
        return int(O.walk_uniform53(seed, call["epoch"], start, walk, step)) / 9007199254740992.0

    inj = H.UniformInjector(sampler, u_fn)

    class SamplerProxy:               # get_embeddings only calls batch_sample_neighbors
        def batch_sample_neighbors(self, nodes, T):
            r = inj.batch_sample(nodes, T)
            call["epoch"] += 1        # one Philox epoch per sampling call (per layer)
            return r

    with torch.no_grad():
        emb_full = model.get_embeddings(x, SamplerProxy(), num_neighbors=10).numpy()
        emb_mlp = model(x).numpy()
        rng = np.random.Generator(np.random.PCG64(8))
        nb0, wt0 = ragged_lists(M, M, 10, rng)
        nb1, wt1 = ragged_lists(M, M, 10, rng)
        emb_lists = model(x, None, [nb0, nb1], [wt0, wt1]).numpy()
        emb_shared = model(x, None, [nb0], [wt0]).numpy() if False else None
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    torch.manual_seed(2)
    gcl = ly.GraphConvLayer(Hd, 24).eval()
    with torch.no_grad():
        gcl.bn.running_mean.uniform_(-0.1, 0.1); gcl.bn.running_var.uniform_(0.5, 1.5)
        gcl.bn.weight.uniform_(0.5, 1.5); gcl.bn.bias.uniform_(-0.1, 0.1)
        gx = torch.randn(40, Hd, generator=torch.Generator().manual_seed(4))
        gn = torch.randn(40, Hd, generator=torch.Generator().manual_seed(5))
        g_out = gcl(gx, gn).numpy()
        g_out1 = gcl(gx[:1], gn[:1]).numpy()        # single row: BatchNorm skipped (layers.py:68)
    gsd = {"gcl." + k: v.detach().numpy() for k, v in gcl.state_dict().items()
           if k != "bn.num_batches_tracked"}
    meta = dict(nb0=[[int(v) for v in l] for l in nb0], wt0=wt0,
                nb1=[[int(v) for v in l] for l in nb1], wt1=wt1)
    np.savez_compressed(os.path.join(OUT, "forward.npz"), x=x.numpy(), edge_index=ei, edge_weights=w,
                        seed=seed, W=100, L=2, T=10, dims=np.array([F_, Hd, E_, layers]),
                        emb_full=emb_full, emb_mlp=emb_mlp, emb_lists=emb_lists,
                        lists=json.dumps(meta), gx=gx.numpy(), gn=gn.numpy(), g_out=g_out,
                        g_out1=g_out1, **{"sd." + k: v for k, v in sd.items()}, **gsd)
    print("forward.npz: emb_full", emb_full.shape, "row norms", np.linalg.norm(emb_full, axis=1)[:3])


def make_exact():
    g = torch.Generator().manual_seed(9)
    emb = torch.nn.functional.normalize(torch.randn(500, 32, generator=g), dim=1)
    qs = [0, 3, 17, 250, 499]
    ids = np.stack([ev.generate_recommendations(emb, q, k=10) for q in qs])
    ids_incl = np.stack([ev.generate_recommendations(emb, q, k=10, exclude_query=False) for q in qs])
    np.savez_compressed(os.path.join(OUT, "exact.npz"), emb=emb.numpy(), queries=np.array(qs),
                        ids=ids, ids_incl=ids_incl)
    print("exact.npz", ids.shape)


def make_evaluation():
    g = torch.Generator().manual_seed(13)
    emb = torch.nn.functional.normalize(torch.randn(700, 32, generator=g), dim=1)
    pairs = torch.stack([torch.randint(0, 700, (300,), generator=g), torch.randint(0, 700, (300,), generator=g)], 1)
    q, gt = pairs[:, 0].numpy(), pairs[:, 1].numpy()
    ks = [1, 10, 50, 100, 500]
    hr = np.array([ev.calculate_hit_rate(emb, q, gt, k=k) for k in ks])
    m = ev.calculate_mrr(emb, q, gt)
    m7 = ev.calculate_mrr(emb, q, gt, scale=7)
    res = ev.evaluate_embeddings(emb, {"positive_pairs": pairs})
    # ranks as the reference's MRR loop sees them (torch.sort descending, first match)
    ranks = np.array([int(np.where(torch.sort(emb[a] @ emb.t(), descending=True)[1].numpy() == b)[0][0]) + 1
                      for a, b in zip(q, gt)])
    np.savez_compressed(os.path.join(OUT, "evaluation.npz"), emb=emb.numpy(), pairs=pairs.numpy(), ks=np.array(ks),
                        hit_rates=hr, mrr=m, mrr_scale7=m7, ranks=ranks,
                        evaluate=json.dumps({k: float(v) for k, v in res.items()}))
    print("evaluation.npz", hr, m)


if __name__ == "__main__":
    make_walks()
    make_pooling()
    make_forward()
    make_exact()
    make_evaluation()
