#!/usr/bin/env python
"""bench_search.py -- retrieval half of the hot path (BASELINE.json configs[2] and [3]), as a library
for bench.py (`run_retrieval`, every rank calls it) and as a stand-alone CLI.

All-item top-10 queries over ML-25M-shaped embeddings (N = 62,423, d = 128):
  exact_ip / exact_l2  tcgen05 TF32 GEMM fused with a shortlist + fp32 re-rank + certificate (E1 / E2)
  lsh_exhaustive       256-bit codes, exhaustive Hamming top-k -- what the reference's faiss.IndexLSH
                       computes (L1/L2); +-1 bf16 GEMM on the tensor cores, exact
  lsh_tables_hamming / lsh_tables_dot   256 bits x 16 tables: bucket probe + dedup + re-rank (L3)
  ivf                  "Weak AND" = IVF nlist 100, nprobe 20 (I1/I2), probe-masked tensor-core scoring
Embedding set B of SURVEY.md 8(d) (1,024 clusters + 0.3 noise, L2-normalised); set A = collapsed
embeddings like the reference checkpoint produces (bench.py passes the ones it just computed).

Per method: value = queries/s by CUDA events with queries and index resident in HBM (max over
ranks); e2e = queries/s through the drop-in class (`.search(numpy) -> numpy`), host wall clock;
recall@10 vs exact (utils/nearest_neighbors.py:243-251); roofline (tensor pipe or HBM); cpu_baseline =
the numpy restatement on a bounded sample of queries ("port, not faiss": faiss is absent offline).
N > 1 (torch.distributed initialised): queries are split 1/N per rank, the index is replicated, the
result lists are all-gathered (sharding.search_query_sharded); exact_ip additionally runs item-sharded
(every rank scores all queries against its item block, all-gather + pb200_topk_merge) and is checked
bitwise against the query-sharded result.
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


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], source="MEASURED_PEAKS.json")
    except Exception:                       # noqa: BLE001
        return dict(hbm=6650.0, bf16=1600.0, source="fallback (B200_PROFILING.md)")


def recall(exact_ids, ids, k):
    return float(np.mean([len(set(a) & set(b)) / k for a, b in zip(exact_ids.tolist(), ids.tolist())]))


def run_retrieval(dev, emb=None, set_name="B", n=62423, d=128, k=10, iters=3, cpu=True, cpu_queries=256,
                  methods=("exact_ip", "exact_l2", "lsh_exhaustive", "lsh_tables_hamming", "lsh_tables_dot", "ivf"),
                  item_sharded=True):
    """Returns the `retrieval` dict (identical on every rank).  emb: [n, d] float32 tensor (any device)
    or None for synthetic set B."""
    import torch.distributed as dist
    import mre_b200  # noqa: F401
    from mre_b200 import synthetic as S, kernels as K, _native as N, sharding as SH
    from mre_b200.utils.nearest_neighbors import LSHIndex, WeakANDIndex, FlatL2Index

    rank, ws = SH.world()
    if emb is None:
        emb = S.spread_embeddings(n, d, seed=1)
    x_np = np.ascontiguousarray(emb.detach().cpu().numpy(), dtype=np.float32)
    n, d = x_np.shape
    xd = torch.from_numpy(x_np).to(dev).contiguous()
    lo, hi = SH.shard_range(n, rank, ws)
    q_dev, q_np = xd[lo:hi].contiguous(), x_np[lo:hi]
    peaks = _peaks()
    tf32_peak = peaks["bf16"] / 2.0

    def sync_all():
        torch.cuda.synchronize(dev)
        if ws > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        """fn() -> per-rank result; returns (seconds per call, last result): CUDA events, max over ranks."""
        fn(); sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            out = fn()
        b.record(); torch.cuda.synchronize(dev)
        return max_over_ranks(a.elapsed_time(b) / iters * 1e-3), out

    def timed_host(fn):
        fn(); sync_all()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(dev)
        return max_over_ranks(time.perf_counter() - t0), out

    def gather(s, i):
        return SH.all_gather_rows(s, n), SH.all_gather_rows(i, n)

    res = {}
    exact_ids = None
    launches0 = N.launch_count()

    # ---- exact (E1 inner product, E2 squared L2): tcgen05 TF32 shortlist + fp32 re-rank + certificate ----
    flat = FlatL2Index(d, device=dev); flat.add(x_np)
    for metric, name in ((N.METRIC_IP, "exact_ip"), (N.METRIC_L2, "exact_l2")):
        if name not in methods:
            continue
        st = {}
        t, (s_, i_) = timed(lambda: gather(*K.topk(q_dev, xd, k, metric, stats=st)))
        reruns = int(st["fp32_reruns"].item()) if st.get("fp32_reruns") is not None else None
        flops = 2.0 * n * n * d
        res[name] = dict(value=n / t, ms=t * 1e3, kernel_path=st.get("path"), fp32_reruns_this_rank=reruns,
                         roofline=dict(bound="tensor", achieved=flops / t / 1e12, peak=tf32_peak, unit="TFLOP/s",
                                       frac=flops / t / 1e12 / tf32_peak,
                                       peak_source=f"{peaks['source']} bf16_tflops / 2 (kind::tf32 runs at half the bf16 rate; "
                                                   "no TF32 entry in MEASURED_PEAKS.json)"))
        if name == "exact_l2":
            exact_ids = i_.cpu().numpy()
            te, _ = timed_host(lambda: flat.search(q_np, k))
            res[name]["e2e"] = n / te
        else:
            ip_ids = i_
    if exact_ids is None:
        _s, i_ = gather(*K.topk(q_dev, xd, k, N.METRIC_L2))
        exact_ids = i_.cpu().numpy()
    if "exact_ip" in methods:
        te, _ = timed_host(lambda: (lambda s_, i_: (s_.cpu().numpy(), i_.cpu().numpy()))(
            *K.topk(torch.from_numpy(q_np).to(dev), xd, k, N.METRIC_IP)))
        res["exact_ip"]["e2e"] = n / te
        if ws > 1 and item_sharded:
            # item-sharded: every rank scores ALL queries against its item block, lists are all-gathered
            # and merged under the (score, id) total order -> bitwise the unsharded / query-sharded result
            items_local = xd[lo:hi].contiguous()
            t, (s2, i2) = timed(lambda: SH.exact_search_item_sharded(xd, items_local, lo, k, N.METRIC_IP))
            res["exact_ip_item_sharded"] = dict(value=n / t, ms=t * 1e3,
                                                equals_query_sharded=bool(torch.equal(i2, ip_ids)),
                                                exchange=f"all-gather of [{n}, {k}] (score, id) lists from {ws} ranks + pb200_topk_merge")

    # ---- LSH: one 256-bit sign code per vector ----
    lsh = None
    if any(m.startswith("lsh") for m in methods):
        lsh = LSHIndex(d, 256, 16, device=dev)
        with contextlib.redirect_stdout(sys.stderr):
            lsh.build(x_np)
    if "lsh_exhaustive" in methods:
        def lsh_search():
            cq = K.lsh_encode(q_dev, lsh.projection)
            return gather(*K.hamming_topk(cq, lsh.codes, k))
        t, (hd, hi_) = timed(lsh_search)
        te, _ = timed_host(lambda: lsh.search(q_np, k))
        flops = 2.0 * n * n * 256 + 2.0 * n * d * 256
        res["lsh_exhaustive"] = dict(value=n / t, ms=t * 1e3, e2e=n / te, recall_at_10=recall(exact_ids, hi_.cpu().numpy(), k),
                                     note="hash of the queries included; +-1 bf16 GEMM, exact Hamming distances",
                                     roofline=dict(bound="tensor", achieved=flops / t / 1e12, peak=peaks["bf16"], unit="TFLOP/s",
                                                   frac=flops / t / 1e12 / peaks["bf16"], peak_source=peaks["source"] + " bf16_tflops (burst)"))
    for rerank in ("hamming", "dot"):
        name = "lsh_tables_" + rerank
        if name not in methods:
            continue
        lt = LSHIndex(d, 256, 16, mode="tables", rerank=rerank, projection=lsh.projection, device=dev)
        with contextlib.redirect_stdout(sys.stderr):
            lt.build(x_np)

        def tables_search():
            cq = K.lsh_encode(q_dev, lt.projection)
            s_, i_, nc = K.lsh_search_tables(cq, lt.codes, 16, *lt._tables, k, queries=q_dev if rerank == "dot" else None,
                                             vectors=lt.vectors if rerank == "dot" else None)
            return gather(s_, i_) + (nc,)
        t, (ts, ti, nc) = timed(tables_search)
        te, _ = timed_host(lambda: lt.search(q_np, k))
        cu = float(nc.float().mean())
        per_q = 16 * 8 + 4 * cu + (4 * d if rerank == "dot" else 32) * cu + 12 * k + 4 * d + 32
        res[name] = dict(value=n / t, ms=t * 1e3, e2e=n / te, recall_at_10=recall(exact_ids, ti.cpu().numpy(), k),
                         mean_unique_candidates=cu,
                         roofline=dict(bound="hbm", achieved=per_q * n / t / 1e9, peak=peaks["hbm"], unit="GB/s",
                                       frac=per_q * n / t / 1e9 / peaks["hbm"], peak_source=peaks["source"],
                                       note="bytes per query = bucket headers + ids and codes (or vectors) of the UNIQUE candidates "
                                            "(raw candidate count not returned: lower bound); index is L2 resident at this size"))

    # ---- IVF "Weak AND" ----
    if "ivf" in methods:
        ivf = WeakANDIndex(d, 100, 10, device=dev)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            ivf.build(x_np)
        torch.cuda.synchronize(dev); build_s = time.perf_counter() - t0
        st = {}
        lay = ivf._tc_layout
        use_tc = lay is not None and q_dev.size(0) >= K.TOPK_TC_MIN_QUERIES and \
            K.ivf_search_tc_supported(q_dev.size(0), lay[0].size(0), d, k, 100)

        def ivf_search():
            _, probes = K.topk(q_dev, ivf.centroids, 20, N.METRIC_L2)
            if use_tc:
                return gather(*K.ivf_search_tc(q_dev, probes, *ivf._lists, lay, 100, k, stats=st))
            return gather(*K.ivf_search(q_dev, probes, *ivf._lists, k))
        t, (vd, vi) = timed(ivf_search)
        te, _ = timed_host(lambda: ivf.search(q_np, k))
        flops = 2.0 * n * 100 * d + 2.0 * d * n * (n * 20 / 100.0)
        res["ivf"] = dict(value=n / t, ms=t * 1e3, e2e=n / te, recall_at_10=recall(exact_ids, vi.cpu().numpy(), k),
                          build_s=build_s, kernel_path="tf32" if use_tc else "fp32 list scan",
                          list_scan_reruns_this_rank=int(st["list_scan_reruns"].item()) if st.get("list_scan_reruns") is not None else None,
                          roofline=dict(bound="tensor", achieved=flops / t / 1e12, peak=tf32_peak, unit="TFLOP/s",
                                        frac=flops / t / 1e12 / tf32_peak,
                                        note="algorithmic flops = coarse quantiser + probed lists (20 of 100); the probe-masked "
                                             "kernel scores every tile", peak_source=f"{peaks['source']} bf16_tflops / 2"))

    launches = N.launch_count() - launches0

    # ---- CPU restatements on a bounded sample of queries (rank 0, N = 1) ----
    if cpu and ws == 1:
        from oracle import oracle as O
        nq = min(cpu_queries, n)
        threads = torch.get_num_threads()

        def cpu_time(fn):
            t0 = time.perf_counter(); fn(); return time.perf_counter() - t0
        if "exact_ip" in res:
            tc = cpu_time(lambda: O.exact_ip(x_np, x_np[:4 * nq], k))
            res["exact_ip"]["cpu_baseline"] = dict(value=4 * nq / tc, unit="queries/s", kind="port", cores=threads,
                                                   sample=f"{4 * nq} queries vs all {n} items, numpy matmul + top-k")
        if "exact_l2" in res:
            tc = cpu_time(lambda: O.exact_l2(x_np, x_np[:4 * nq], k))
            res["exact_l2"]["cpu_baseline"] = dict(value=4 * nq / tc, unit="queries/s", kind="port (numpy restatement, not faiss)",
                                                   cores=threads, sample=f"{4 * nq} queries vs all {n} items")
        if lsh is not None:
            codes_np = lsh.codes.cpu().numpy()
            if "lsh_exhaustive" in res:
                tc = cpu_time(lambda: O.lsh_search_exhaustive(codes_np, codes_np[:nq], k))
                res["lsh_exhaustive"]["cpu_baseline"] = dict(value=nq / tc, unit="queries/s", kind="port (numpy restatement, not faiss)",
                                                             cores=threads, sample=f"{nq} queries vs all {n} codes")
            for rerank in ("hamming", "dot"):
                if "lsh_tables_" + rerank in res:
                    m = min(64, nq)
                    tc = cpu_time(lambda: O.lsh_search_tables(codes_np, codes_np[:m], k, 16, vectors=x_np if rerank == "dot" else None,
                                                              queries=x_np[:m] if rerank == "dot" else None))
                    res["lsh_tables_" + rerank]["cpu_baseline"] = dict(value=m / tc, unit="queries/s", kind="port (numpy restatement)",
                                                                       cores=threads, sample=f"{m} queries")
        if "ivf" in res:
            cent, assign = ivf.centroids.cpu().numpy(), ivf.assign.cpu().numpy()
            tc = cpu_time(lambda: O.ivf_search(x_np, cent, assign, x_np[:nq], k, 20))
            res["ivf"]["cpu_baseline"] = dict(value=nq / tc, unit="queries/s", kind="port (numpy restatement, not faiss)",
                                              cores=threads, sample=f"{nq} queries")
    return dict(metric="top-10 queries/sec", unit="queries/s", embedding_set=set_name, n_items=n, dim=d, k=k, queries=n,
                n_gpus=ws, iters=iters, gpu_launches=launches,
                sharding="none" if ws == 1 else f"queries split 1/{ws} per rank, index replicated, result lists all-gathered",
                methods=res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=62423)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--set", default="B", choices=["A", "B"])
    ap.add_argument("--cpu-queries", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    emb = None
    if args.set == "A":     # collapsed: a common direction plus small noise (pairwise cosine ~ 0.985)
        g = torch.Generator().manual_seed(1)
        emb = torch.nn.functional.normalize(torch.randn(1, args.d, generator=g) + 0.125 * torch.randn(args.n, args.d, generator=g), dim=1)
    with contextlib.redirect_stdout(sys.stderr):
        r = run_retrieval(dev, emb, args.set, args.n, args.d, args.k, cpu_queries=args.cpu_queries)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
