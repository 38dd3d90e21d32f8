"""Per-kernel CUDA-event times of ONE rank's sharded embedding step (config C2), under torchrun:
    torchrun --nproc-per-node N tools/profile_sharded.py > profiles/r2_sharded_step_breakdown_N.json
The step is issued eagerly, op by op, exactly as sharding.get_embeddings_sharded issues it (same
wrappers, same order), with an event between ops; 20 repetitions, medians per rank.  Also times the
whole step as a CUDA-graph replay for comparison (what bench.py measures)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import mre_b200  # noqa: F401
from mre_b200 import synthetic as S, kernels as K, sharding as SH, _native as N, neighbor_lists as NL
from mre_b200.utils.random_walk import RandomWalkSampler
from mre_b200.graphs import GraphedEmbeddings
import bench

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
inp = bench.build_inputs("C2")
M, T = inp["M"], 10
sampler = RandomWalkSampler(torch.from_numpy(inp["ei"]), torch.from_numpy(inp["w"]), 2, 100, seed=1234, device=dev,
                            num_nodes=M + inp["U"])
model, _ = bench.model_weights(inp)
model = model.to(dev).eval()
mine = SH.local_slice(M, rank, ws)
xd = inp["x"][mine].to(dev)
rows = xd.size(0)
nodes = torch.arange(mine.start, mine.stop, mine.step or 1, dtype=torch.int32, device=dev)
CYC = N.SHARD_CYCLIC if SH.EMB_LAYOUT == "cyclic" else N.SHARD_BLOCKS
for _ in range(3):
    SH.get_embeddings_sharded(model, xd, sampler, M, T)
srows = SH.shard_size(M, ws)
pb = SH.peer_buffers(model.num_layers, srows, model.input_proj.out_features, dev)
assert pb is not None, "peer exchange unavailable"
P = lambda lin: (lin.weight, lin.bias)
RND, PRE = N.EPI_ROUND_TF32, N.IN_A1_TF32
names, reps = [], []
for rep in range(20):
    ev = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); ev.append((name, e))
    dist.barrier(); torch.cuda.synchronize()
    mark("start")
    h = K.gather_dense(xd, *P(model.input_proj), flags=N.EPI_RELU | RND, precision=model.precision, out=pb.local(0)[:rows])
    mark("input_proj")
    batches = sampler.sample_layers(nodes, T, model.num_layers); mark("walks_all_layers")
    for i in range(model.num_layers):
        pb.barrier(); mark(f"barrier{i}")
        wf, bf = model._folded_layer(i)
        ids, wts, ll, wl = batches[i].as_args()
        hn = K.pool_sharded(pb.ptr_array(i), ws, srows, M, h.size(1), ids, wts, ll, wl, N.POOL_PINSAGE | N.POOL_ROUND_TF32, dev, layout=CYC)
        mark(f"pool{i}")
        h = K.gather_dense(h, wf, bf, a2=hn, flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE | N.IN_A2_TF32,
                           precision=model.precision, out=pb.local(i + 1)[:rows] if i + 1 < model.num_layers else None)
        mark(f"conv{i}")
    emb = K.gather_dense(h, *P(model.output_proj), flags=N.EPI_L2NORM | PRE, precision=model.precision)
    mark("output_proj")
    torch.cuda.synchronize()
    pb.check()
    names = [n for n, _ in ev[1:]]
    reps.append([ev[i][1].elapsed_time(ev[i + 1][1]) * 1e3 for i in range(len(ev) - 1)])
med = np.median(np.array(reps), axis=0)
g = GraphedEmbeddings(model, xd, sampler, T, num_items=M)
for _ in range(3):
    g.replay()
ts = []
for _ in range(20):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(check=False); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
SH.check_peer_barriers(force=True)
mine = dict(rank=rank, rows=rows, eager_us={n: round(float(v), 1) for n, v in zip(names, med)},
            eager_sum_us=round(float(med.sum()), 1), graph_replay_us=round(float(np.median(ts)), 1),
            graph_nodes=g.launches_per_replay)
allr = [None] * ws
dist.all_gather_object(allr, mine)
if rank == 0:
    print(json.dumps(dict(n_gpus=ws, workload="C2", note="eager per-op CUDA-event medians (us) include the launch gaps between "
                          "consecutive eager launches; graph_replay_us is the whole step replayed as one CUDA graph", ranks=allr), indent=1))
dist.destroy_process_group()
