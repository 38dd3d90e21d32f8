"""N2 timing: hit-rate@k / MRR for P (query, ground truth) pairs over C3-sized embeddings through
pb200_rank_of_target, next to the reference's own loop (utils/evaluation.py:5-73 restated in the
oracle: one matmul + sort per pair) on a bounded sample of pairs."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mre_b200  # noqa: F401
from mre_b200 import kernels as K, synthetic as S
from mre_b200.utils.evaluation import evaluate_embeddings
from oracle import oracle as O

n, d, P = 62423, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 62423
e = S.spread_embeddings(n, d, seed=1)
ed = e.cuda().contiguous()
g = torch.Generator().manual_seed(0)
pairs = torch.stack([torch.randint(0, n, (P,), generator=g), torch.randint(0, n, (P,), generator=g)], 1)
q, t = pairs[:, 0].cuda().int(), pairs[:, 1].cuda().int()
K.rank_of_target(ed, q, t); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    r = K.rank_of_target(ed, q, t)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
t0 = time.perf_counter(); res = evaluate_embeddings(ed, {"positive_pairs": pairs}); e2e = time.perf_counter() - t0
m = 64
t0 = time.perf_counter(); ro = O.rank_of_target(e.numpy(), pairs[:m, 0].numpy(), pairs[:m, 1].numpy()); cpu = time.perf_counter() - t0
print(json.dumps({"metric": "evaluation pairs/sec (hit-rate@k + MRR)", "value": P / (ms * 1e-3), "unit": "pairs/s", "ms": ms,
                  "pairs": P, "n_items": n, "dim": d, "tflops_fp32": 2 * P * n * d / (ms * 1e-3) / 1e12,
                  "e2e_evaluate_embeddings_pairs_per_s": P / e2e, "results": {k: float(v) for k, v in res.items()},
                  "ranks_equal_oracle_sample": bool((r[:m].cpu().numpy() == ro).all()),
                  "cpu_baseline": {"value": m / cpu, "unit": "pairs/s", "kind": "port (numpy restatement of the per-pair loop)",
                                   "cores": torch.get_num_threads(), "sample": f"{m} pairs"}}))
