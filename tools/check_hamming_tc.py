"""GPU check + timing of the tensor-core exhaustive Hamming search against the popcount kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mre_b200  # noqa: F401
from mre_b200 import kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 62423
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g).cuda()


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


t1, (d1, i1) = timed(lambda: K.hamming_topk(codes, codes, 10, precision="tc"))
t0, (d0, i0) = timed(lambda: K.hamming_topk(codes, codes, 10, precision="simt"), iters=1)
print(f"hamming 256 bits: tc {t1:.3f} ms ({n / t1 * 1e3 / 1e6:.2f} M q/s)  simt {t0:.3f} ms  "
      f"dist equal {torch.equal(d1, d0)}  ids equal {torch.equal(i1, i0)}", flush=True)
