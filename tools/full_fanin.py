"""K2 at FULL fan-in ("the item-item graph makes all T neighbours valid items, which exercises
K2's gather at full fan-in").  On the bipartite C2 graph a sampled node has ~1.25 valid item neighbours; here the
catalogue of C2 (62,423 items, C2's features and checkpoint weights) gets a synthetic item-item graph -- 50 random
item neighbours per item, half-star weights -- so that every one of the T = 10 sampled neighbours is a valid row.
Times, with CUDA events and an L2 flush between iterations: the whole get_embeddings step (CUDA-graph replay),
the pool kernel alone and the conv GEMM alone, and prints the pool kernel's algorithmic bytes
(4H (1 + T_valid) + 4H + 8T per node) against the measured HBM peak.
Usage: python tools/full_fanin.py > profiles/r2_k2_full_fanin.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mre_b200  # noqa: F401
from mre_b200 import kernels as K, _native as N
from mre_b200.utils.random_walk import RandomWalkSampler
from mre_b200.graphs import GraphedEmbeddings
import bench

dev = torch.device("cuda", 0)
inp = bench.build_inputs("C2")
M, T, DEG = inp["M"], 10, 50
g = torch.Generator().manual_seed(0)
src = torch.arange(M).repeat_interleave(DEG)
dst = torch.randint(0, M, (M * DEG,), generator=g)
w = 0.5 * torch.randint(1, 11, (M * DEG,), generator=g).float()
sampler = RandomWalkSampler(torch.stack([src, dst]), w, 2, 100, seed=1234, device=dev, num_nodes=M)
model, weights = bench.model_weights(inp)
model = model.to(dev).eval()
x = inp["x"].to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
peak = 6548.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:        # noqa: BLE001
    pass


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = dict(workload="C2 catalogue, synthetic item-item graph (50 random item neighbours per item)", items=M, T=T,
           weights=weights, hbm_peak_gbs=peak)
try:
    graphed = GraphedEmbeddings(model, x, sampler, T, num_items=M)
    out["step_ms"] = timed(lambda: graphed.replay())
    out["items_per_s"] = M / (out["step_ms"] * 1e-3)
except Exception as e:        # noqa: BLE001
    out["step_error"] = repr(e)
try:
    nodes = torch.arange(M, dtype=torch.int32, device=dev)
    batches = sampler.sample_layers(nodes, T, model.num_layers, epoch=0)
    ids, wts, ll, wl = batches[0].as_args()
    H = model.input_proj.out_features
    h = torch.nn.functional.normalize(torch.randn(M, H, device=dev), dim=1)
    h = K.round_tf32(h) if hasattr(K, "round_tf32") else h
    t_valid = float(((ids >= 0) & (ids < M) & (torch.arange(T, device=dev)[None, :] < ll[:, None])).sum()) / M
    out["valid_item_neighbours_per_node"] = t_valid
    mode = N.POOL_PINSAGE | N.POOL_ROUND_TF32
    out["pool_ms"] = timed(lambda: K.pool(h, ids, wts, ll, wl, mode))
    pool_bytes = M * (4 * H * t_valid + 4 * H + 8 * T)          # neighbour rows read + pooled row written + lists
    out["pool_algorithmic_bytes"] = pool_bytes
    out["pool_gbs"] = pool_bytes / (out["pool_ms"] * 1e-3) / 1e9
    out["pool_frac_of_hbm_peak"] = out["pool_gbs"] / peak
    out["pool_note"] = ("h is 64 MB: L2-resident on a B200 (126 MB), so the algorithmic fraction may exceed what DRAM "
                        "delivers; it is the SURVEY 8(d) figure")
    h_neigh = K.pool(h, ids, wts, ll, wl, mode)
    wf, bf = model._folded_layer(0)
    flags = N.EPI_RELU | N.EPI_L2NORM | N.EPI_ROUND_TF32 | N.IN_A1_TF32 | N.IN_A2_TF32
    out["conv_gemm_ms"] = timed(lambda: K.gather_dense(h, wf, bf, a2=h_neigh, flags=flags, precision=model.precision))
except Exception as e:        # noqa: BLE001
    out["op_error"] = repr(e)
print(json.dumps(out))
