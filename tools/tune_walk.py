"""Times the walk kernel variants (PB200_WALK_VARIANT) on config C2 in separate processes:
bit1 = binary in-node search, bit2 = register top-T selection,
bits 4.. = min resident blocks per SM (register cap).  Usage: python tools/tune_walk.py [variants...]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
import mre_b200
from mre_b200 import kernels as K
d = np.load("/tmp/c2_graph.npz")
csr = K.csr_build(torch.from_numpy(d["ei"]), torch.from_numpy(d["w"]), num_nodes=int(d["N"]))
nodes = torch.arange(62423, dtype=torch.int32, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for _ in range(3): K.walk_topt(csr, nodes, 100, 2, 10, 1234, 0)
ts = []
for e in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = K.walk_topt(csr, nodes, 100, 2, 10, 1234, e); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
chk = int(out[0].long().sum()) ^ int(out[1].long().sum())
print("variant", %s, "walk_ms mean %%.4f min %%.4f chk %%d" %% (np.mean(ts), np.min(ts), chk), flush=True)
'''
if not os.path.exists("/tmp/c2_graph.npz"):
    sys.path.insert(0, ROOT)
    import numpy as np, mre_b200
    from mre_b200 import synthetic as S
    M, U, R = S.CONFIGS["C2"][:3]
    ei, w = S.bipartite_graph(M, U, R, seed=0)
    np.savez("/tmp/c2_graph.npz", ei=ei, w=w, N=M + U)
variants = sys.argv[1:] or ["0", "1", "2", "4", "6", "7", "80", "86", "96", "102"]
for v in variants:
    env = dict(os.environ, PB200_WALK_VARIANT=v)
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, v)], env=env, capture_output=True, text=True, timeout=300)
    print((r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
