"""Times the walk kernel on config C2 in separate processes, one per variant.
A variant is LEAF[:MINBLOCKS[:TABLE[:L2FETCH[:SMEMPAD[:W[:L[:CARVEOUT%]]]]]]] -- LEAF in bucket / compact / wide (PB200_WALK_LEAF),
MINBLOCKS = resident blocks per SM the lean bucket kernel is compiled for (PB200_WALK_MINBLOCKS),
TABLE = visit-table version of the lean kernel (PB200_WALK_TABLE: 0 atomics, 1 / 2 match.any + plain
stores with 256 / 512 slots), OLDVARIANT = PB200_WALK_VARIANT of the tree-index kernel.  Prints mean / min ms per launch and a
checksum of the outputs (must be identical across variants: all of them are bit-exact).
Usage: python tools/tune_walk.py [variants...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
import mre_b200
from mre_b200 import kernels as K
import os
from mre_b200 import _native as NV
if os.environ.get("PB200_L2_FETCH"):
    NV.check(NV.lib().pb200_set_l2_fetch_granularity(int(os.environ["PB200_L2_FETCH"])))
d = np.load("/tmp/c2_graph.npz")
t0 = time.time()
csr = K.csr_build(torch.from_numpy(d["ei"]), torch.from_numpy(d["w"]), num_nodes=int(d["N"]))
torch.cuda.synchronize(); t_build = time.time() - t0
nodes = torch.arange(0, 62423, int(os.environ.get("TUNE_STRIDE", "1")), dtype=torch.int32, device="cuda")
E = int(os.environ.get("TUNE_E", "0")) or None
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
W = int(os.environ.get("TUNE_W", "100")); L = int(os.environ.get("TUNE_L", "2"))
for _ in range(3): K.walk_topt(csr, nodes, W, L, 10, 1234, 0, num_epochs=E)
ts = []
for e in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = K.walk_topt(csr, nodes, W, L, 10, 1234, e, num_epochs=E); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
chk = int(out[0].long().sum()) ^ (int(out[1].long().sum()) << 20) ^ int((out[2].double() * 1e6).sum().item())
print("variant", %r, "l2fetch", NV.lib().pb200_get_l2_fetch_granularity(), "walk_ms mean %%.4f min %%.4f chk %%d index_MB %%.0f fmt %%d build_s %%.2f" %% (
    np.mean(ts), np.min(ts), chk, csr.index_nbytes() / 1e6, csr.leaf_format, t_build), flush=True)
'''
if not os.path.exists("/tmp/c2_graph.npz"):
    sys.path.insert(0, ROOT)
    import numpy as np, mre_b200
    from mre_b200 import synthetic as S
    M, U, R = S.CONFIGS["C2"][:3]
    ei, w = S.bipartite_graph(M, U, R, seed=0)
    np.savez("/tmp/c2_graph.npz", ei=ei, w=w, N=M + U)
variants = sys.argv[1:] or ["bucket:6:0", "bucket:6:0:32", "bucket:6:3:32", "bucket:6:4:32", "bucket:6:10:32", "bucket:6:13:32", "bucket:6:14:32", "bucket:8:14:32", "compact:::32"]
for v in variants:
    v, _, extra = v.partition("@")          # "...@KEY=VAL,KEY=VAL": extra environment for the child
    parts = v.split(":")
    env = dict(os.environ, PB200_WALK_LEAF=parts[0])
    if len(parts) > 1 and parts[1]:
        env["PB200_WALK_MINBLOCKS"] = parts[1]
    if len(parts) > 2 and parts[2]:
        env["PB200_WALK_TABLE"] = parts[2]
    if len(parts) > 3 and parts[3]:
        env["PB200_L2_FETCH"] = parts[3]
    if len(parts) > 4 and parts[4]:
        env["PB200_WALK_PAD"] = parts[4]
    if len(parts) > 5 and parts[5]:
        env["TUNE_W"] = parts[5]
    if len(parts) > 6 and parts[6]:
        env["TUNE_L"] = parts[6]
    if len(parts) > 7 and parts[7]:
        env["PB200_WALK_CARVE"] = parts[7]
    for kv in filter(None, extra.split(",")):
        k_, _, v_ = kv.partition("=")
        env[k_] = v_
    v = v + ("@" + extra if extra else "")
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, v)], env=env, capture_output=True, text=True, timeout=300)
    print((r.stdout.strip() or r.stderr.strip()[-600:]), flush=True)
