#!/usr/bin/env python
"""bench_c5.py -- BASELINE.json configs[4] ("C5"): the scaled synthetic graph -- random-walk sampling + 3-layer
embedding with peer-memory pooling, then item-sharded exact top-10 merged over NVLink -- as a library for
bench.py (`run_c5`, every rank calls it) and a stand-alone CLI (`torchrun ... bench_c5.py --scale 0.0625`).

C5 proper is 10 M items / 50 M users / 2 G ratings.  What runs here is C5 x `scale` (default 1/16: 625 k items,
3.125 M users, 125 M ratings = 250 M directed edges, mean degrees unchanged), generated ON THE DEVICE
(synthetic.bipartite_graph_device, Philox), replicated per rank.  The largest scale that was run is 1/2 (5 M items,
30 M nodes, 2 G directed edges; the sampling index switches to 32-bit ids beyond 2^24 nodes).  scale = 1 is NOT
run: the int64 edge list (64 GB) + weights + pb200_csr_build's sort workspace for 4 G directed edges exceed
180 GB on one GPU -- it needs the graph itself sharded, see DESIGN.md section 6.
Steps, each timed by CUDA events (max over ranks):
  graph generation, CSR + sampling index build (one-off)
  embeddings: 3 x sampling (one launch) + input projection + 3 conv layers + output projection, rows dealt
              round-robin to the ranks, neighbour rows read from peer memory
  search:     exact inner-product top-10 of EVERY item against the catalogue, item-sharded: each rank scores all
              queries (in blocks) against its item block, lists all-gathered and merged (pb200_topk_merge)
Parity spot checks on the generated graph (rank 0): CSR rows of sampled nodes == a stable filter of the edge
list; walk / count / top-T of sampled start nodes == the C oracle run on the rows those walks can touch.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def _ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def run_c5(dev, scale=1.0 / 16, steps=3, check=True, query_block=65536, search_queries=None):
    """search_queries: score only the first so-many items as queries (None = every item) -- bounds the all-pairs
    search of the larger scales; the throughput figures are per query either way."""
    import torch.distributed as dist
    import mre_b200  # noqa: F401
    from mre_b200 import synthetic as S, kernels as K, sharding as SH, _native as N
    from mre_b200.utils.random_walk import RandomWalkSampler
    from mre_b200.model.pinsage import PinSage
    from mre_b200.graphs import GraphedEmbeddings

    rank, ws = SH.world()
    M, U, R, F_, Hd, E_, layers = S.c5_config(scale)
    T = 10

    def sync_max(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    ei, w = S.bipartite_graph_device(M, U, R, seed=0, device=dev)
    torch.cuda.synchronize(dev)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    sampler = RandomWalkSampler(ei, w, 2, 100, seed=1234, device=dev, num_nodes=M + U)
    torch.cuda.synchronize(dev)
    csr_s = time.perf_counter() - t0
    csr = sampler.csr

    # ---- parity spot checks on THIS graph (rank 0; the other ranks hold the same graph: same seed) ----
    checks = None
    if check and rank == 0:
        from oracle import oracle as O
        rng = np.random.Generator(np.random.PCG64(5))
        starts = np.unique(np.concatenate([[0, 1, M - 1], rng.integers(0, M, 29)])).astype(np.int64)
        row_ptr_d, col_d, cum_d = csr.row_ptr, csr.col, csr.cum
        # (a) CSR rows == stable filter of the edge list (edge order defines the CDF order)
        rows_ok = True
        for v in starts[:6].tolist() + [int(M + rng.integers(0, U))]:
            a, b = int(row_ptr_d[v]), int(row_ptr_d[v + 1])
            sel = (ei[0] == v).nonzero().view(-1)
            want_col = ei[1][sel].to(torch.int32)
            want_cum = torch.cumsum((w[sel] * (1 << csr.quant_shift)).to(torch.int64), 0)
            rows_ok = rows_ok and bool(torch.equal(col_d[a:b], want_col)) and \
                bool(torch.equal(cum_d[a:b].to(torch.int64) & 0xFFFFFFFF, want_cum))
        # (b) walks from sampled starts == the C oracle on the rows these walks can reach (start rows + their users' rows)
        need = set(starts.tolist())
        for v in starts.tolist():
            need.update(col_d[int(row_ptr_d[v]):int(row_ptr_d[v + 1])].tolist())
        need = np.array(sorted(need), dtype=np.int64)
        need_t = torch.from_numpy(need).to(dev)
        lens = (row_ptr_d[need_t + 1] - row_ptr_d[need_t]).cpu().numpy()
        sub_ptr = np.zeros(M + U + 1, dtype=np.int64)
        sub_ptr[need + 1] = lens
        sub_ptr = np.cumsum(sub_ptr)
        seg = torch.cat([torch.arange(int(row_ptr_d[v]), int(row_ptr_d[v + 1]), device=dev) for v in need.tolist()])
        sub_col = col_d[seg].cpu().numpy()
        sub_cum = (cum_d[seg].cpu().numpy()).view(np.uint32)
        o = O.c_walk_topt(sub_ptr, sub_col, sub_cum, starts, 100, 2, T, 1234, 0)
        g_ids, g_cnt, g_w, g_nv = K.walk_topt(csr, torch.from_numpy(starts), 100, 2, T, 1234, 0)
        walks_ok = bool(np.array_equal(g_ids.cpu().numpy(), o["ids"]) and np.array_equal(g_cnt.cpu().numpy(), o["counts"])
                        and np.array_equal(g_w.cpu().numpy(), o["w32"]) and np.array_equal(g_nv.cpu().numpy(), o["nvalid"]))
        checks = dict(csr_rows_equal_stable_filter_of_edge_list=rows_ok, walks_equal_c_oracle=walks_ok,
                      sampled_start_nodes=int(starts.size), rows_handed_to_the_oracle=int(need.size))
        del seg, need_t
    graph_bytes = ei.numel() * 8 + w.numel() * 4
    del ei, w
    torch.cuda.empty_cache()

    # ---- embeddings: 3 samples + forward, rows dealt round-robin, CUDA-graph replay per rank ----
    torch.manual_seed(0)
    model = PinSage(F_, Hd, E_, layers).to(dev).eval()
    mine = SH.local_slice(M, rank, ws) if ws > 1 else slice(0, M, 1)
    g = torch.Generator(device=dev).manual_seed(0)
    x_full = torch.randn(M, F_, generator=g, device=dev)
    x_dev = x_full[mine].contiguous()
    del x_full
    nodes = torch.arange(mine.start, mine.stop, mine.step or 1, dtype=torch.int32, device=dev)
    for _ in range(2):
        emb = SH.get_embeddings_sharded(model, x_dev, sampler, M, T) if ws > 1 else model.get_embeddings(x_dev, sampler, T)
    graphed = GraphedEmbeddings(model, x_dev, sampler, T, num_items=M)
    for _ in range(2):
        graphed.replay()
    torch.cuda.synchronize(dev)
    if ws > 1:
        dist.barrier()
    a = _ev()
    for _ in range(steps):
        emb = graphed.replay(check=False)
    b = _ev(); torch.cuda.synchronize(dev)
    if ws > 1:
        SH.check_peer_barriers(force=True)
    emb_ms = sync_max(a.elapsed_time(b) / steps)
    # walk kernel alone (this rank's starts, all layers in one launch) for its roofline
    ts = []
    for _ in range(3):
        e0 = _ev(); sampler.sample_layers(nodes, T, layers); e1 = _ev(); torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1))
    walk_ms = float(np.median(ts))
    _i, _c, _w2, _nv, trace = K.walk_topt(csr, nodes[:65536], 100, 2, T, 1234, 0, return_trace=True, num_epochs=layers)
    deg = csr.row_ptr[1:] - csr.row_ptr[:-1]
    cur = torch.cat([nodes[:65536].view(1, -1, 1, 1).expand(layers, -1, 100, 1), trace[:, :, :, :-1]], dim=3).long()
    ex = trace >= 0
    per_step = 16 + 4 * torch.ceil(torch.log2(deg[cur.clamp_min(0)].double() + 1)) + 4
    sample_n = min(65536, nodes.numel())
    algo = (float((per_step * ex).sum()) + layers * sample_n * (4 + 12 * T)) * (nodes.numel() / sample_n)
    del trace, cur, ex, per_step
    peak = 6548.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:        # noqa: BLE001
        pass

    # ---- item-sharded exact top-10 of every item: all-gather of per-shard lists + merge ----
    emb_full = SH.all_gather_rows(emb, M, layout=SH.EMB_LAYOUT) if ws > 1 else emb
    lo, hi = SH.shard_range(M, rank, ws)
    items_local = emb_full[lo:hi].contiguous()
    torch.cuda.synchronize(dev)
    if ws > 1:
        dist.barrier()
    a = _ev()
    Q = M if not search_queries else min(M, int(search_queries))
    ids_out = torch.empty((Q, T), dtype=torch.int32, device=dev)
    for q0 in range(0, Q, query_block):
        q = emb_full[q0:min(q0 + query_block, Q)]
        _s, i_ = SH.exact_search_item_sharded(q, items_local, lo, T, N.METRIC_IP)
        ids_out[q0:q0 + q.size(0)] = i_
    b = _ev(); torch.cuda.synchronize(dev)
    search_ms = sync_max(a.elapsed_time(b))
    search_ok = None
    if check:                                   # sharded result == unsharded on a block of queries (rank-count invariant)
        q = emb_full[:4096].contiguous()
        _s, i_ref = K.topk(q, emb_full, T, N.METRIC_IP)
        search_ok = bool(torch.equal(i_ref, ids_out[:4096]))
    flops = 2.0 * Q * M * E_
    reruns = {}
    _s, _i = K.topk(emb_full[:8192].contiguous(), items_local, T, N.METRIC_IP, stats=reruns)
    rr = reruns.get("fp32_reruns")
    rerun_frac = None if rr is None else float(rr.item()) / 8192
    # the same search over a SPREAD catalogue of the same size (set B of SURVEY 8(d)): random-init PinSage
    # embeddings are near-collinear, the TF32 certificate sends (almost) every query to the fp32 kernel
    del emb_full, items_local, ids_out
    spread = S.spread_embeddings(M, E_, seed=1).to(dev)
    items_b = spread[lo:hi].contiguous()
    torch.cuda.synchronize(dev)
    if ws > 1:
        dist.barrier()
    a = _ev()
    for q0 in range(0, Q, query_block):
        SH.exact_search_item_sharded(spread[q0:min(q0 + query_block, Q)], items_b, lo, T, N.METRIC_IP)
    b = _ev(); torch.cuda.synchronize(dev)
    spread_ms = sync_max(a.elapsed_time(b))
    return dict(workload=f"C5 x {scale:g}: {M:,} items / {U:,} users / {R:,} ratings ({2 * R:,} directed edges), {layers} layers, "
                         "generated on the device (Philox)",
                scale=scale, n_gpus=ws, items=M, users=U, ratings=R, layers=layers, data="synthetic (device generator)",
                graph_gen_s=round(gen_s, 2), csr_and_index_build_s=round(csr_s, 2),
                graph_bytes=graph_bytes, csr_bytes=csr.nbytes(), walk_index_bytes=csr.index_nbytes(),
                walk_index={N.LEAF_BUCKET: "bucket (8 slots, 24-bit ids)", N.LEAF_BUCKET32: "bucket32 (6 slots, 32-bit ids)"}.get(csr.leaf_format, "tree"),
                embeddings=dict(value=M / (emb_ms * 1e-3), unit="items/s", ms_per_step=emb_ms, steps=steps,
                                step_launch="cuda graph replay per rank",
                                exchange="neighbour rows of h read from peer memory" if ws > 1 else "none"),
                roofline=dict(kernel="walk_bucket_batched_kernel", bound="hbm", achieved=algo / (walk_ms * 1e-3) / 1e9, peak=peak,
                              unit="GB/s", frac=algo / (walk_ms * 1e-3) / 1e9 / peak, avg_launch_ms=walk_ms,
                              algorithmic_bytes_per_launch=algo, start_nodes_this_rank=int(nodes.numel()),
                              samples_per_launch=layers, traffic=None,
                              note="algorithmic bytes extrapolated from the traced first 65,536 start nodes of this rank"),
                search=dict(method="exact inner product, item-sharded: every rank scores all queries against its item block; "
                                   "all-gather of per-shard (score, id) lists + pb200_topk_merge",
                            value=Q / (search_ms * 1e-3), unit="queries/s", ms=search_ms, queries=Q, catalogue=M,
                            tflops=flops / (search_ms * 1e-3) / 1e12, equals_unsharded_on_first_4096_queries=search_ok,
                            fp32_rerun_fraction_first_8192_queries=rerun_frac,
                            note="embeddings of a randomly initialised model are near-collinear: the TF32 shortlist cannot be "
                                 "certified and the exact fp32 kernel re-runs those queries (results are the fp32 kernel's either way)"),
                search_spread_catalogue=dict(value=Q / (spread_ms * 1e-3), unit="queries/s", ms=spread_ms, queries=Q, catalogue=M,
                                             tflops=flops / (spread_ms * 1e-3) / 1e12,
                                             data="set B of SURVEY 8(d) at C5's catalogue size (1,024 clusters + 0.3 noise)"),
                parity_spot_checks=checks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0 / 16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--search-queries", type=int, default=0, help="score only the first N items as queries (0 = all)")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    r = run_c5(dev, args.scale, args.steps, search_queries=args.search_queries or None)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
