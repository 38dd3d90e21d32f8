"""Per-role cycle accounting of dense_tc_kernel (needs a PB200_TC_PROFILE build of the library)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mre_b200
from mre_b200 import kernels as K, _native as N
lib = ctypes.CDLL(N.LIB_PATH)
buf = (ctypes.c_ulonglong * (148 * 12))()
torch.manual_seed(0)
dev = "cuda"; M = 62423; T = 10
h = torch.randn(M, 256, device=dev); x = torch.randn(M, 128, device=dev)
ids = torch.randint(0, 8 * M, (M, T), dtype=torch.int32, device=dev)     # ~12% valid, like C2
wt = torch.rand(M, T, device=dev); ll = torch.full((M,), T, dtype=torch.int32, device=dev)
w_in = torch.randn(256, 128, device=dev) / 11; w_cv = torch.randn(256, 512, device=dev) / 22
w_out = torch.randn(128, 256, device=dev) / 16; b256 = torch.randn(256, device=dev); b128 = torch.randn(128, device=dev)
P = N.PREC_TF32
cases = dict(inp=lambda: K.gather_dense(x, w_in, b256, flags=1, precision=P),
             conv=lambda: K.gather_dense(h, w_cv, b256, pool_x=h, lists=(ids, wt, ll, None), flags=3 | 8 | 16, precision=P),
             conv_dense=lambda: K.gather_dense(h, w_cv, b256, a2=h, flags=3, precision=P),
             out=lambda: K.gather_dense(h, w_out, b128, flags=2 | 16, precision=P))
names = ["P:lists", "P:issue", "P:data+reduce", "P:wait empty", "P:pass", "C:wait empty", "C:issue", "M:wait copy", "M:wait pool", "M:issue", "E:wait acc", "E:work"]
for name, fn in cases.items():
    for _ in range(2): fn()
    lib.pb200_debug_tc_profile(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    lib.pb200_debug_tc_profile(buf, 0)
    arr = np.array(buf[:], dtype=np.float64).reshape(148, 12).mean(0) / 1.965e3    # us per SM (one warp/lane each)
    print(f"{name:11s} (us per SM, whole launch) " + " | ".join(f"{n} {v:5.1f}" for n, v in zip(names, arr)))
