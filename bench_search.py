#!/usr/bin/env python
"""bench_search.py -- retrieval half of the hot path (BASELINE.json configs[2] and [3]).

All-item top-10 queries over ML-25M-shaped embeddings (N = 62,423, d = 128):
  exact inner product / exact L2 (pb200_topk), LSH 256 bits exhaustive Hamming (what the
  reference's faiss.IndexLSH computes), LSH 256 bits x 16 tables (bucket probe + dedup +
  popcount re-rank), IVF "Weak AND" (nlist = 100, nprobe = 20).
Embedding set B of SURVEY.md 8(d) (1,024 clusters + 0.3 noise, L2-normalised) by default;
`--set A` uses collapsed embeddings like the reference checkpoint produces.
Prints one JSON line per method: queries/s (CUDA events, inputs resident in HBM), e2e
queries/s through the drop-in class with host numpy in/out, recall@10 vs exact
(utils/nearest_neighbors.py:243-251), and the numpy restatement timed on a bounded sample of
queries ("restatement, not faiss": faiss is absent offline).
"""
import argparse
import contextlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3, out


def recall(exact_ids, ids, k):
    return float(np.mean([len(set(a) & set(b)) / k for a, b in zip(exact_ids.tolist(), ids.tolist())]))


def main():
    with contextlib.redirect_stdout(sys.stderr):
        lines = _run()
    for l in lines:
        print(json.dumps(l))


def _run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=62423)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--set", default="B", choices=["A", "B"])
    ap.add_argument("--cpu-queries", type=int, default=256)
    args = ap.parse_args()
    import mre_b200  # noqa: F401
    from mre_b200 import synthetic as S, kernels as K, _native as N
    from mre_b200.utils.nearest_neighbors import LSHIndex, WeakANDIndex, FlatL2Index
    from oracle import oracle as O

    n, d, k = args.n, args.d, args.k
    if args.set == "B":
        x = S.spread_embeddings(n, d, seed=1)
    else:   # collapsed: a common direction plus small noise (pairwise cosine ~ 0.985)
        g = torch.Generator().manual_seed(1)
        x = torch.nn.functional.normalize(torch.randn(1, d, generator=g) + 0.125 * torch.randn(n, d, generator=g), dim=1)
    xd = x.cuda().contiguous()
    x_np = x.numpy()
    base = dict(unit="queries/s", n_items=n, dim=d, k=k, queries=n, embedding_set=args.set, data="synthetic")
    lines = []

    # ---- exact: tensor-core shortlist + fp32 re-rank + certificate (bitwise = the fp32 kernel) ----
    for metric, name in ((N.METRIC_IP, "exact_ip (E1)"), (N.METRIC_L2, "exact_l2 (E2, IndexFlatL2)")):
        st = {}
        t, (s_tc, i_tc) = timed(lambda: K.topk(xd, xd, k, metric, precision="tf32", stats=st))
        t32, (s_32, i_32) = timed(lambda: K.topk(xd, xd, k, metric, precision="fp32"), iters=1)
        lines.append(dict(base, method=name + " tcgen05 tf32 shortlist + fp32 re-rank", value=n / t, ms=t * 1e3,
                          tflops=2 * n * n * d / t / 1e12, fp32_reruns=int(st["fp32_reruns"].item()),
                          bitwise_equal_to_fp32_kernel=bool(torch.equal(i_tc, i_32) and torch.equal(s_tc, s_32)),
                          fp32_kernel_queries_per_s=n / t32, fp32_kernel_ms=t32 * 1e3))
    i_l2 = i_tc
    exact_ids = i_l2.cpu().numpy()
    flat = FlatL2Index(d); flat.add(x_np)
    t0 = time.perf_counter(); flat.search(x_np, k); lines[-1]["e2e"] = n / (time.perf_counter() - t0)

    # ---- LSH exhaustive (reference behaviour) ----
    lsh = LSHIndex(d, 256, 16)
    lsh.build(x_np)
    cq = lsh.codes
    t_enc, _ = timed(lambda: K.lsh_encode(xd, lsh.projection))
    t, (hd, hi) = timed(lambda: K.hamming_topk(cq, cq, k, precision="tc"))
    t_pop, (hd0, hi0) = timed(lambda: K.hamming_topk(cq, cq, k, precision="simt"), iters=1)
    t0 = time.perf_counter(); _d, ids = lsh.search(x_np, k); e2e = n / (time.perf_counter() - t0)
    lines.append(dict(base, method="lsh_exhaustive 256 bits (L1/L2, faiss.IndexLSH behaviour)", value=n / (t + t_enc),
                      ms=(t + t_enc) * 1e3, encode_ms=t_enc * 1e3, e2e=e2e, recall_at_10=recall(exact_ids, ids, k),
                      hamming_ms=t * 1e3, kernel="tcgen05 +-1 bf16 GEMM + fused shortlist (exact)",
                      equal_to_popcount_kernel=bool(torch.equal(hd, hd0) and torch.equal(hi, hi0)),
                      popcount_kernel_ms=t_pop * 1e3, popcount_kernel_queries_per_s=n / (t_pop + t_enc)))
    # CPU restatement on a bounded sample of queries
    nq = min(args.cpu_queries, n)
    codes_np = cq.cpu().numpy()
    t0 = time.perf_counter(); O.lsh_search_exhaustive(codes_np, codes_np[:nq], k); tc = time.perf_counter() - t0
    lines[-1]["cpu_baseline"] = dict(value=nq / tc, unit="queries/s", kind="port (numpy restatement, not faiss)",
                                     cores=torch.get_num_threads(), sample=f"{nq} queries vs all {n} codes")

    # ---- LSH 16 tables ----
    for rerank in ("hamming", "dot"):
        lt = LSHIndex(d, 256, 16, mode="tables", rerank=rerank, projection=lsh.projection)
        lt.build(x_np)
        vec = lt.vectors if rerank == "dot" else None
        t, (ts, ti, nc) = timed(lambda: K.lsh_search_tables(cq, cq, 16, *lt._tables, k,
                                                            queries=xd if rerank == "dot" else None, vectors=vec))
        t0 = time.perf_counter(); _d, ids = lt.search(x_np, k); e2e = n / (time.perf_counter() - t0)
        lines.append(dict(base, method=f"lsh_tables 256 bits x 16 tables, {rerank} re-rank (L3)", value=n / (t + t_enc),
                          ms=(t + t_enc) * 1e3, e2e=e2e, recall_at_10=recall(exact_ids, ids, k),
                          mean_unique_candidates=float(nc.float().mean())))

    # ---- IVF ----
    ivf = WeakANDIndex(d, 100, 10)
    t0 = time.perf_counter(); ivf.build(x_np); torch.cuda.synchronize(); build_s = time.perf_counter() - t0

    def ivf_search_simt():
        _, probes = K.topk(xd, ivf.centroids, 20, N.METRIC_L2)
        return K.ivf_search(xd, probes, *ivf._lists, k)

    ivf_stats = {}

    def ivf_search_tc():
        _, probes = K.topk(xd, ivf.centroids, 20, N.METRIC_L2)
        return K.ivf_search_tc(xd, probes, *ivf._lists, ivf._tc_layout, 100, k, stats=ivf_stats)
    t_simt, (vd0, vi0) = timed(ivf_search_simt, iters=1)
    t, (vd, vi) = timed(ivf_search_tc)
    t0 = time.perf_counter(); _d, ids = ivf.search(x_np, k); e2e = n / (time.perf_counter() - t0)
    lines.append(dict(base, method="ivf_weak_and nlist=100 nprobe=20 (I1/I2) tcgen05 probe-masked shortlist + fp32 re-rank",
                      value=n / t, ms=t * 1e3, e2e=e2e, recall_at_10=recall(exact_ids, ids, k), build_s=build_s,
                      list_scan_reruns=int(ivf_stats["list_scan_reruns"].item()),
                      bitwise_equal_to_list_scan_kernel=bool(torch.equal(vi, vi0) and torch.equal(vd, vd0)),
                      list_scan_kernel_queries_per_s=n / t_simt, list_scan_kernel_ms=t_simt * 1e3))
    t0 = time.perf_counter()
    O.ivf_search(x_np, ivf.centroids.cpu().numpy(), ivf.assign.cpu().numpy(), x_np[:nq], k, 20)
    tc = time.perf_counter() - t0
    lines[-1]["cpu_baseline"] = dict(value=nq / tc, unit="queries/s", kind="port (numpy restatement, not faiss)",
                                     cores=torch.get_num_threads(), sample=f"{nq} queries")
    for l in lines:
        l["metric"] = "top-10 queries/sec"
    return lines


if __name__ == "__main__":
    main()
