#!/usr/bin/env python
"""bench.py -- item embeddings/s of the full 2-layer PinSage get_embeddings hot path.

Workload (BASELINE.json configs[1], "C2"): ML-25M-shaped synthetic graph (62,423 items,
162,541 users, 25,000,095 ratings), F=128 -> H=256 -> E=128, 2 layers, W=100 walks of length 2,
T=10 neighbours.  One step = 2 x [walk/count/top-T kernel over all items] + fused forward
(input projection, 2 fused gather+dense conv layers, output projection + L2 norm).

  value  items/s with the feature matrix already resident in HBM (CUDA events, per step)
  e2e    items/s through the public API (PinSage.get_embeddings) with HOST buffers: pinned
         features H2D and embeddings D2H inside the timed region, host wall clock
  roofline   the walk kernel (dominant): algorithmic bytes (summed exactly
             over the executed steps of one launch from its trace) / CUDA-event duration
  cpu_baseline   the oracle port (C walk sampler on all host threads + numpy forward) on a
                 bounded sample, rank 0 at N=1

  c5         BASELINE configs[4] at --c5-scale (default 1/16 on 1-2 GPUs: 625 k items, 250 M directed edges, 3 layers;
             1/8 on 4-8 GPUs: 1.25 M items, 500 M directed edges), graph
             generated on the device: embeddings/s, walk roofline at DRAM scale, item-sharded exact top-10 with
             the all-gather + merge, parity spot checks against the C oracle (bench_c5.run_c5)
  retrieval  BASELINE metric (2): all-item top-10 queries/s for exact / LSH / IVF search over C3-shaped
             embeddings (bench_search.run_retrieval): set B (spread) at every N, set A (the embeddings
             this run just computed from the reference checkpoint: collapsed) at N = 1

N > 1 (torchrun): rows are split across ranks (strong scaling over the fixed catalogue); the
pooling kernel reads neighbour rows of h from peer memory over NVLink, one flag barrier on peer
memory per layer; the per-rank step is replayed as a CUDA graph (sampling epochs advance on the
device; the e2e graph also holds the pinned-host upload, forked under the walks, and the download);
time = max over ranks.  One untimed step is compared bitwise with the unsharded result on rank 0
("sharded_equals_single"); retrieval runs query-sharded (and exact search item-sharded + merge).
`--impl reference` times the CPU port alone (rank 0) and prints the same JSON shape.
"""
import argparse
import contextlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "item embeddings/sec (2-layer PinSage)"
UNIT = "items/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2"])
    ap.add_argument("--cpu-sample", type=int, default=8192, help="start items per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="auto", choices=["fp32", "tf32", "auto"])
    ap.add_argument("--no-graph", action="store_true", help="N > 1: launch the sharded step eagerly")
    ap.add_argument("--no-retrieval", action="store_true", help="skip the retrieval (search) section")
    ap.add_argument("--no-c5", action="store_true", help="skip the scaled C5 workload")
    ap.add_argument("--c5-scale", type=float, default=None,
                    help="C5 = 10 M items / 50 M users / 2 G ratings times this (default: 1/16 on 1-2 GPUs, 1/8 on 4-8)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass
        self.active = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                        if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_inputs(workload):
    import mre_b200  # noqa: F401
    from mre_b200 import synthetic as S
    M, U, R, F_, Hd, E_, layers = S.CONFIGS[workload]
    t0 = time.time()
    ei, w = S.bipartite_graph(M, U, R, seed=0)
    x = S.features(M, F_, seed=0)
    return dict(M=M, U=U, R=R, F=F_, H=Hd, E=E_, layers=layers, ei=ei, w=w, x=x,
                gen_s=time.time() - t0)


# ----------------------------------------------------------------------------- CPU port
def model_weights(inp):
    """C2 / C3 use the reference's shipped checkpoint (checkpoints/best_model.pt, committed as
    tests/golden/checkpoint.npz by tests/golden/make_golden_r2.py); other shapes a seeded default init."""
    from mre_b200.model.pinsage import PinSage
    path = os.path.join(ROOT, "tests", "golden", "checkpoint.npz")
    model = PinSage(inp["F"], inp["H"], inp["E"], inp["layers"])
    if os.path.exists(path):
        g = np.load(path)
        if tuple(int(v) for v in g["dims"]) == (inp["F"], inp["H"], inp["E"], inp["layers"]):
            model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
            return model, "reference checkpoint best_model.pt (tests/golden/checkpoint.npz)"
    torch.manual_seed(0)
    return PinSage(inp["F"], inp["H"], inp["E"], inp["layers"]), "seeded default init (torch.manual_seed(0))"


def cpu_port_setup(inp):
    from oracle import oracle as O
    row_ptr, col, cum = O.c_csr_build(inp["ei"], inp["w"], inp["M"] + inp["U"], 1)
    model, _src = model_weights(inp)
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    return dict(O=O, csr=(row_ptr, col, cum), sd=sd, threads=O.c_oracle().orc_max_threads())


def cpu_port_step(inp, cp, sample, step):
    """Embeds the first `sample` items as a self-contained catalogue: same per-item work as the
    full job (W*L walk steps on the full graph, the same dense FLOPs per row)."""
    O = cp["O"]
    row_ptr, col, cum = cp["csr"]
    nbrs, wts = [], []
    for layer in range(inp["layers"]):
        o = O.c_walk_topt(row_ptr, col, cum, np.arange(sample), 100, 2, 10, 1234,
                          epoch=step * inp["layers"] + layer, num_threads=0)
        nv = o["nvalid"]
        nbrs.append([o["ids"][r, :nv[r]].tolist() for r in range(sample)])
        wts.append([o["w64"][r, :nv[r]].tolist() for r in range(sample)])
    return O.pinsage_forward(inp["x"][:sample].numpy(), cp["sd"], inp["layers"], nbrs, wts,
                             dtype=np.float32)


def run_cpu_port(inp, sample, steps, warmup):
    cp = cpu_port_setup(inp)
    sample = min(sample, inp["M"])
    for s in range(warmup):
        cpu_port_step(inp, cp, sample, s)
    t0 = time.perf_counter()
    for s in range(steps):
        cpu_port_step(inp, cp, sample, warmup + s)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=sample / dt, unit=UNIT, cores=cp["threads"], kind="port", extrapolated=sample < inp["M"],
                warmup_steps=warmup,
                sample=f"{sample} of {inp['M']} start items per step (walks on the full graph, "
                       f"W=100 L=2 T=10, {inp['layers']} layers + forward on those rows); items/s "
                       f"extrapolates linearly; C walk port ({cp['O'].C_BUILD}) on all host threads + numpy fp32 forward",
                ms_per_step=dt * 1e3)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp = build_inputs(args.workload)
    warm = min(args.warmup, 1)        # one warm-up step of ~0.3 s is enough for a CPU loop; reported as run
    r = run_cpu_port(inp, args.cpu_sample, args.steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "warmup_requested": args.warmup,
            "same_config": False, "extrapolated": r["extrapolated"],
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, inp, 1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, inp, n_gpus):
    return {"workload": f"{args.workload}: ML-25M-shaped synthetic graph" if args.workload == "C2"
            else f"{args.workload}: tiny synthetic graph",
            "items": inp["M"], "users": inp["U"], "ratings": inp["R"],
            "dims": [inp["F"], inp["H"], inp["E"]], "layers": inp["layers"], "num_walks": 100,
            "walk_length": 2, "num_neighbors": 10, "parallelism": f"rows/{n_gpus}" + (" (dealt round-robin)" if n_gpus > 1 else "")}


# ----------------------------------------------------------------------------- B200 arm
def walk_roofline(K, sampler, nodes, T, layers, walk_ms, dev_ms, peaks_path):
    """Roofline of the dominant kernel (walk / count / top-T) for THIS rank's start nodes: algorithmic bytes
    (16 + 4 ceil(log2(deg+1)) + 4 per executed step, 4 + 12 T per start node) summed
    exactly over the executed steps of one traced launch / mean CUDA-event duration of the launches."""
    peaks = {}
    try:
        peaks = json.load(open(peaks_path))
    except Exception:      # noqa: BLE001
        pass
    peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else \
        (6650.0, "fallback (B200_PROFILING.md)")
    _ids, _c, _w, _nv, trace = K.walk_topt(sampler.csr, nodes, 100, 2, T, 1234, 0, return_trace=True,
                                           num_epochs=layers)          # [layers, n, W, L]: what one launch executes
    deg = (sampler.csr.row_ptr[1:] - sampler.csr.row_ptr[:-1])
    cur = torch.cat([nodes.view(1, -1, 1, 1).expand(layers, -1, 100, 1), trace[:, :, :, :-1]], dim=3).long()
    executed = trace >= 0
    d = deg[cur.clamp_min(0)].double()
    per_step = 16 + 4 * torch.ceil(torch.log2(d + 1)) + 4
    algo_bytes = float((per_step * executed).sum()) + layers * nodes.numel() * (4 + 12 * T)
    avg_ms = float(np.mean(walk_ms))
    achieved = algo_bytes / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "walk_traffic.json")))
        if nodes.numel() == tj.get("start_nodes") and layers == tj.get("samples_per_launch", 1):
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:      # noqa: BLE001
        pass
    return {"kernel": "walk_bucket_batched_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src or "not measured in-run (needs ncu); see profiles/",
            "peak_source": which, "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_ms,
            "start_nodes_this_rank": int(nodes.numel()), "samples_per_launch": layers,
            "launch": f"one launch = {layers} independent samples (one per conv layer) of every start node",
            "executed_steps": int(executed.sum()),
            "share_of_step": sum(walk_ms) / dev_ms,
            "limiter": "ncu (profiles/r2_walk_bucket_batched_ncu_details.txt): LSU data pipe 72 % of peak wavefronts "
                       "(one wavefront per lane for the divergent 16 B meta and 32 B bucket loads + shared-memory atomics "
                       "of the visit table), issue slots 65 %, DRAM 55 %: two dependent loads per step, latency / LSU bound"}


def main_b200(args):
    import torch.distributed as dist
    import mre_b200  # noqa: F401
    from mre_b200 import _native as N, kernels as K, neighbor_lists as NL, sharding as SH
    from mre_b200.utils.random_walk import RandomWalkSampler
    from mre_b200.graphs import GraphedEmbeddings
    import bench_search as BS

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)

    inp = build_inputs(args.workload)
    M, T, layers = inp["M"], 10, inp["layers"]
    t0 = time.perf_counter()
    sampler = RandomWalkSampler(torch.from_numpy(inp["ei"]), torch.from_numpy(inp["w"]), 2, 100,
                                seed=1234, device=dev, num_nodes=M + inp["U"])
    torch.cuda.synchronize()
    csr_s = time.perf_counter() - t0
    model, weights_src = model_weights(inp)
    model = model.to(dev).eval()
    model.precision = N.PRECISIONS[args.precision]
    mine = SH.local_slice(M, rank, ws) if ws > 1 else slice(0, M, 1)     # rows dealt round-robin to the ranks
    x_host = inp["x"][mine].contiguous().pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty((x_host.size(0), inp["E"]), dtype=torch.float32).pin_memory()
    nodes = torch.arange(mine.start, mine.stop, mine.step or 1, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = graphed_e2e = None   # N > 1: the sharded step is launch bound -> one CUDA graph per rank (graphs.py)

    def step_device(walk_events=None):
        if ws > 1:
            if graphed is not None:
                return graphed.replay(check=False)
            return SH.get_embeddings_sharded(model, x_dev, sampler, M, T, check_barriers=False)
        if walk_events is not None:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        batches = sampler.sample_layers(nodes, T, layers)           # ONE launch: both layers' samples
        if walk_events is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
            walk_events.append((e0, e1))
        return model.forward(x_dev, None, batches, None)

    def step_e2e():
        if ws > 1:
            if graphed_e2e is not None:          # upload (forked under the walks) + step + download: one graph
                graphed_e2e.replay(check=False)
                torch.cuda.current_stream(dev).synchronize()
                return
            emb = SH.get_embeddings_sharded(model, x_host, sampler, M, T, check_barriers=False)
            out_host.copy_(emb, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        else:
            model.get_embeddings(x_host, sampler, T, out=out_host)

    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_device(); step_e2e()
    barrier()
    if ws > 1:
        SH.check_peer_barriers(force=True)

    # ---- N > 1: one untimed sharded step vs the unsharded step on rank 0, bitwise ----
    sharded_equals_single = None
    if ws > 1:
        e0 = sampler.epoch
        emb_mine = SH.get_embeddings_sharded(model, x_dev, sampler, M, T)
        full = SH.all_gather_rows(emb_mine, M, layout=SH.EMB_LAYOUT)
        if rank == 0:
            sampler.epoch = e0
            single = model.get_embeddings(inp["x"].to(dev), sampler, T)
            sharded_equals_single = bool(torch.equal(full, single))
            del single
        sampler.epoch = e0 + layers
        del full, emb_mine
        barrier()

    if ws > 1 and not args.no_graph:
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            graphed = GraphedEmbeddings(model, x_dev, sampler, T, num_items=M)
            graphed_e2e = GraphedEmbeddings(model, x_host, sampler, T, num_items=M, out=out_host)
        except Exception as e:          # noqa: BLE001 -- capture unsupported here: keep the eager step
            print(f"[bench] rank {rank}: CUDA-graph capture failed ({e}); eager launches", file=sys.stderr)
            graphed, graphed_e2e, ok[0] = None, None, 0
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # all ranks replay, or none
        if int(ok.item()) == 0:
            graphed = graphed_e2e = None
        for _ in range(3):
            step_device(); step_e2e()
        barrier()

    # ---- timed: device-resident inputs, per-step CUDA events, L2 flushed between steps ----
    walk_events, step_ms = [], []
    launches0 = N.launch_count()
    clocks.active = True
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        step_device(walk_events)
        b.record()
        step_ms.append((a, b))
    barrier()
    wall_dev = time.perf_counter() - wall0
    launches = N.launch_count() - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps   # replays bypass the ABI's launch counter
    dev_ms = sum(a.elapsed_time(b) for a, b in step_ms)
    walk_ms = [a.elapsed_time(b) for a, b in walk_events]

    # ---- timed: end to end through the public API with host buffers ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks.active = False
    clocks.stop_flag = True
    if ws > 1:
        SH.check_peer_barriers(force=True)          # a barrier that gave up invalidates the run: raise

    # N > 1: the walk kernel of this rank's shard, timed eagerly (the graph replays are timed as a whole)
    if ws > 1:
        for _ in range(4):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); sampler.sample_layers(nodes, T, layers); e1.record()
            walk_events.append((e0, e1))
        torch.cuda.synchronize()
        walk_ms = [a.elapsed_time(b) for a, b in walk_events]

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = t.tolist()

    # ---- retrieval (every rank takes part: query-sharded / item-sharded at N > 1) ----
    retrieval = None
    if not args.no_retrieval:
        with contextlib.redirect_stdout(sys.stderr):
            retrieval = BS.run_retrieval(dev, None, "B (1,024 clusters + 0.3 noise, SURVEY 8(d))", cpu=not args.no_cpu_baseline)
            if ws == 1:
                sampler.epoch = 0
                emb_a = model.get_embeddings(x_dev, sampler, T)
                ra = BS.run_retrieval(dev, emb_a, "A (this run's C2 embeddings from the reference checkpoint: collapsed, "
                                      "SURVEY fact 9)", cpu=False, iters=1, methods=("exact_ip", "exact_l2", "lsh_exhaustive", "ivf"))
                retrieval["set_A"] = {k: ra[k] for k in ("embedding_set", "methods")}
                retrieval["set_A"]["mean_pairwise_cosine_sample"] = float((emb_a[:2048] @ emb_a[:2048].t()).mean())
    # ---- C5 (BASELINE.json configs[4]) at --c5-scale: device-generated graph, 3 layers, item-sharded search ----
    c5 = None
    if not args.no_c5:
        import bench_c5 as BC
        del flush
        torch.cuda.empty_cache()
        with contextlib.redirect_stdout(sys.stderr):
            c5 = BC.run_c5(dev, scale=args.c5_scale if args.c5_scale else (1.0 / 16 if ws <= 2 else 1.0 / 8))
    if rank != 0:
        if ws > 1:
            dist.destroy_process_group()
        return

    value = M * args.steps / (dev_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "tf32 (fp32 accumulate; walks: u32/u64 integer)", "data": "synthetic",
            "config": dict(workload_config(args, inp, ws), l2="flushed between steps (256 MB "
                           "write); sampling index 0.6 GB exceeds the 126 MB L2",
                           weights=weights_src,
                           step_launch="cuda graph replay per rank" if graphed is not None else "eager",
                           exchange=os.environ.get("PB200_SHARD_EXCHANGE", "p2p") + " (neighbour rows of h read from peer memory)" if ws > 1 else "none",
                           csr_build_s=round(csr_s, 3), graph_gen_s=round(inp["gen_s"], 1),
                           csr_bytes=sampler.csr.nbytes(), walk_index_bytes=sampler.csr.index_nbytes(),
                           walk_index={N.LEAF_BUCKET: "bucket", N.LEAF_BUCKET32: "bucket32"}.get(sampler.csr.leaf_format, "tree"),
                           wall_s_timed_region=round(wall_dev, 4)),
            "e2e": {"value": M * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": inp["x"].numel() * 4, "d2h_bytes_per_step": M * inp["E"] * 4,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "path": "PinSage.get_embeddings(pinned x, sampler, T, out=pinned)" if ws == 1 else
                            ("GraphedEmbeddings(pinned x shard, out=pinned).replay() per rank" if graphed_e2e is not None
                             else "get_embeddings_sharded(pinned x shard) + D2H per rank")},
            "gpu_launches": launches, "clocks": clocks.summary()}
    if sharded_equals_single is not None:
        line["sharded_equals_single"] = sharded_equals_single
    if walk_ms:
        line["roofline"] = walk_roofline(K, sampler, nodes, T, layers, walk_ms, dev_ms if ws == 1 else float("nan"),
                                         os.path.join(ROOT, "MEASURED_PEAKS.json"))
        if ws > 1:
            line["roofline"]["share_of_step"] = None
            line["roofline"]["note"] = "rank 0's shard of the start nodes, kernel timed eagerly after the graph-replay region"
    if retrieval is not None:
        line["retrieval"] = retrieval
    if c5 is not None:
        line["c5"] = c5
    if ws == 1 and not args.no_cpu_baseline:
        r = run_cpu_port(inp, args.cpu_sample, 3, 1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")}
    print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
