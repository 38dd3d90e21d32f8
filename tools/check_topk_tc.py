"""Quick GPU check + timing of the tensor-core exact search against the fp32 kernel.
Usage: python tools/check_topk_tc.py [n] [d]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mre_b200  # noqa: F401
from mre_b200 import kernels as K, _native as N, synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 62423
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
x = S.spread_embeddings(n, d, seed=1).cuda().contiguous()


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


for metric, name in ((N.METRIC_IP, "ip"), (N.METRIC_L2, "l2")):
    st = {}
    t1, (s1, i1) = timed(lambda: K.topk(x, x, 10, metric, precision="tf32", stats=st))
    t0, (s0, i0) = timed(lambda: K.topk(x, x, 10, metric, precision="fp32"), iters=1)
    print(f"{name}: tf32 {t1:.3f} ms ({n / t1 * 1e3 / 1e6:.2f} M q/s, {2 * n * n * d / t1 / 1e9:.1f} TFLOP/s)  "
          f"fp32 {t0:.3f} ms  ids equal {torch.equal(i1, i0)}  scores equal {torch.equal(s1, s0)}  "
          f"fp32 reruns {int(st['fp32_reruns'].item())}", flush=True)
