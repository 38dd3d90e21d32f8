"""Run under torchrun with N >= 2 ranks: sharded results == single-GPU results (bitwise)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import mre_b200
from mre_b200 import synthetic as S, kernels as K, sharding as SH, _native as N
from mre_b200.utils.random_walk import RandomWalkSampler
from mre_b200.model.pinsage import PinSage

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
M, U, R, F_, Hd, E_ = 3001, 7000, 120000, 64, 128, 64
ei, w = S.bipartite_graph(M, U, R, seed=5)
x = S.features(M, F_)
torch.manual_seed(0)
model = PinSage(F_, Hd, E_, 2).to(dev).eval()
for prec in (N.PREC_FP32, N.PREC_AUTO):
    model.precision = prec
    sampler = RandomWalkSampler(torch.from_numpy(ei), torch.from_numpy(w), 2, 100, seed=11, device=dev, num_nodes=M + U)
    full = model.get_embeddings(x.to(dev), sampler, 10)                 # every rank: unsharded
    sampler.epoch = 0
    lo, hi = SH.shard_range(M, rank, ws)
    for layout, fork in (("cyclic", None), ("blocks", None), ("cyclic", "0"), ("cyclic", "1")):
        # fork: walks on a second stream beside the input projection (default: small shards only) / serialised
        os.environ.pop("PB200_FORK_WALKS", None)
        if fork is not None:
            os.environ["PB200_FORK_WALKS"] = fork
        sampler.epoch = 0
        rows = SH.local_slice(M, rank, ws, layout)
        mine = SH.get_embeddings_sharded(model, x[rows], sampler, M, 10, layout=layout)
        os.environ.pop("PB200_FORK_WALKS", None)
        assert torch.equal(mine, full[rows]), f"rank {rank}: sharded embeddings differ (precision {prec}, {layout}, fork {fork})"
        back = SH.all_gather_rows(mine, M, layout=layout)
        assert torch.equal(back, full), f"rank {rank}: all_gather_rows({layout}) does not reassemble the matrix"
rows = SH.local_slice(M, rank, ws)
# CUDA-graph replays of the sharded step == eager sharded calls with the same epochs
from mre_b200.graphs import GraphedEmbeddings
model.precision = N.PREC_AUTO
sampler.epoch = 100
g = GraphedEmbeddings(model, x[rows].to(dev), sampler, 10, num_items=M)
for k in range(3):
    got = g.replay().clone()
    sampler.epoch = 100 + 2 * k
    want = SH.get_embeddings_sharded(model, x[rows], sampler, M, 10)
    assert torch.equal(got, want), f"rank {rank}: graph replay {k} differs from the eager step"
# pinned host features in, pinned host embeddings out (upload forked under the walks, download at the end)
xh = x[rows].contiguous().pin_memory()
oh = torch.empty((xh.size(0), E_), dtype=torch.float32).pin_memory()
sampler.epoch = 200
gh = GraphedEmbeddings(model, xh, sampler, 10, num_items=M, out=oh)
for k in range(2):
    got = gh.replay()
    torch.cuda.synchronize()
    sampler.epoch = 200 + 2 * k
    want = SH.get_embeddings_sharded(model, x[rows], sampler, M, 10)
    assert got is oh and torch.equal(got, want.cpu()), f"rank {rank}: host-buffer graph replay {k} differs"
# a rank that cannot set up peer memory: ALL ranks must fall back to the all-gather exchange
os.environ["PB200_TEST_PEER_FAIL_RANK"] = "1"
SH._PEER_CACHE.clear()                       # (leaks the earlier buffers: test only)
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    sampler.epoch = 0
    mine_fb = SH.get_embeddings_sharded(model, x[rows], sampler, M, 10)
assert all(not pb.ok for pb in SH._PEER_CACHE.values()), f"rank {rank}: fallback was not collective"
assert torch.equal(mine_fb, full[rows]), f"rank {rank}: all-gather fallback differs"
del os.environ["PB200_TEST_PEER_FAIL_RANK"]
SH._PEER_CACHE.clear()
emb = full
# item-sharded exact search + merge == unsharded
q = emb[:257].contiguous()
excl = torch.arange(257, dtype=torch.int32, device=dev)
s_ref, i_ref = K.topk(q, emb, 10, N.METRIC_IP, exclude_ids=excl)
s_sh, i_sh = SH.exact_search_item_sharded(q, emb[lo:hi].contiguous(), lo, 10, N.METRIC_IP, exclude_ids=excl)
assert torch.equal(i_sh, i_ref) and torch.equal(s_sh, s_ref), f"rank {rank}: item-sharded search differs"
# query-sharded search: results gathered in query order
s_q, i_q = SH.search_query_sharded(lambda ql: K.topk(ql.contiguous(), emb, 10, N.METRIC_L2), emb)
s_r, i_r = K.topk(emb, emb, 10, N.METRIC_L2)
assert torch.equal(i_q, i_r) and torch.equal(s_q, s_r), f"rank {rank}: query-sharded search differs"
dist.barrier()
if rank == 0:
    print("MULTI_GPU_OK", ws, "ranks")
dist.destroy_process_group()
