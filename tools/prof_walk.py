"""One C2-sized walk launch for ncu:  ncu --set full --import-source on -k regex:walk_ -s 2 -c 1 python tools/prof_walk.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mre_b200  # noqa: F401
from mre_b200 import kernels as K, synthetic as S
if not os.path.exists("/tmp/c2_graph.npz"):
    M, U, R = S.CONFIGS["C2"][:3]
    ei, w = S.bipartite_graph(M, U, R, seed=0)
    np.savez("/tmp/c2_graph.npz", ei=ei, w=w, N=M + U)
d = np.load("/tmp/c2_graph.npz")
csr = K.csr_build(torch.from_numpy(d["ei"]), torch.from_numpy(d["w"]), num_nodes=int(d["N"]))
nodes = torch.arange(62423, dtype=torch.int32, device="cuda")
for e in range(4):
    out = K.walk_topt(csr, nodes, 100, 2, 10, 1234, 2 * e, num_epochs=2)      # one launch = both layers' samples (bench.py)
torch.cuda.synchronize()
print("ok", int(out[0].long().sum()))
