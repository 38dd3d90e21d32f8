"""GPU time of every kernel of ONE rank's sharded embedding step (config C2), under torchrun:
    torchrun --nproc-per-node N tools/profile_sharded.py > profiles/r2_sharded_step_breakdown_N.json
Each op of sharding.get_embeddings_sharded (same wrappers, same order, same buffers) is captured in its
OWN CUDA graph and replayed 20 times back to back between two events, so the figures are device time
per launch without the Python / launch overhead an eager per-op timing would include (at 8 ranks the
kernels take 5-40 us, less than the host needs to issue them).  The whole step as one graph replay
(what bench.py measures) is timed beside them.  PROFILE_ROWS_PER_RANK=k restricts the catalogue to the
first k * N items, to look at the per-rank sizes of a larger world on fewer GPUs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import mre_b200  # noqa: F401
from mre_b200 import kernels as K, sharding as SH, _native as N
from mre_b200.utils.random_walk import RandomWalkSampler
from mre_b200.graphs import GraphedEmbeddings
import bench

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
inp = bench.build_inputs("C2")
M, T = inp["M"], 10
if os.environ.get("PROFILE_ROWS_PER_RANK"):
    M = min(M, int(os.environ["PROFILE_ROWS_PER_RANK"]) * ws)
sampler = RandomWalkSampler(torch.from_numpy(inp["ei"]), torch.from_numpy(inp["w"]), 2, 100, seed=1234, device=dev,
                            num_nodes=inp["M"] + inp["U"])
model, _ = bench.model_weights(inp)
model = model.to(dev).eval()
mine = SH.local_slice(M, rank, ws)
xd = inp["x"][:M][mine].to(dev)
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
L = model.num_layers
state = {}


def op_input_proj():
    state["h0"] = K.gather_dense(xd, *P(model.input_proj), flags=N.EPI_RELU | RND, precision=model.precision, out=pb.local(0)[:rows])


def op_walks():
    state["batches"] = sampler.sample_layers(nodes, T, L, epoch=0)


def op_barrier():
    pb.barrier()


def make_pool(i):
    def f():
        ids, wts, ll, wl = state["batches"][i].as_args()
        state[f"hn{i}"] = K.pool_sharded(pb.ptr_array(i), ws, srows, M, 256, ids, wts, ll, wl, N.POOL_PINSAGE | N.POOL_ROUND_TF32, dev, layout=CYC)
    return f


def make_conv(i):
    def f():
        wf, bf = model._folded_layer(i)
        state[f"h{i + 1}"] = K.gather_dense(state[f"h{i}"], wf, bf, a2=state[f"hn{i}"],
                                            flags=N.EPI_RELU | N.EPI_L2NORM | RND | PRE | N.IN_A2_TF32, precision=model.precision,
                                            out=pb.local(i + 1)[:rows] if i + 1 < L else None)
    return f


def op_output():
    state["emb"] = K.gather_dense(state[f"h{L}"], *P(model.output_proj), flags=N.EPI_L2NORM | PRE, precision=model.precision)


def op_pool_local_alias():
    """pool0 with every 'peer' pointer aliased to this rank's own shard: the same kernel, the same address
    arithmetic, no NVLink -- separates the cost of the remote loads from the cost of the kernel."""
    import ctypes
    ids, wts, ll, wl = state["batches"][0].as_args()
    arr = (ctypes.c_void_p * ws)(*[pb._own[0].value] * ws)
    state["_keep_alias"] = arr
    K.pool_sharded(arr, ws, srows, M, 256, ids, wts, ll, wl, N.POOL_PINSAGE | N.POOL_ROUND_TF32, dev, layout=CYC)


def op_pool_unsharded():
    ids, wts, ll, wl = state["batches"][0].as_args()
    K.pool(state["h0"], ids, wts, ll, wl, N.POOL_PINSAGE | N.POOL_ROUND_TF32)


ops = [("input_proj", op_input_proj), ("walks_all_layers", op_walks)]
for i in range(L):
    ops += [(f"barrier{i}", op_barrier), (f"pool{i}", make_pool(i)), (f"conv{i}", make_conv(i))]
ops.append(("output_proj", op_output))
ops += [("x_pool0_all_pointers_local", op_pool_local_alias), ("x_pool0_unsharded_kernel_local_rows_only", op_pool_unsharded)]
times = {}
REPS = 20
for name, fn in ops:
    fn(); torch.cuda.synchronize(); dist.barrier()            # warm (also produces the inputs of the next op)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        g.replay()
    b.record(); torch.cuda.synchronize()
    times[name] = a.elapsed_time(b) * 1e3 / REPS
    dist.barrier()
pb.check()
sampler.epoch = 0
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
res = dict(rank=rank, rows=rows, per_op_us={n: round(float(v), 1) for n, v in times.items()},
           per_op_sum_us=round(float(sum(v for n, v in times.items() if not n.startswith("x_"))), 1), step_graph_replay_us=round(float(np.median(ts)), 1),
           graph_nodes=g.launches_per_replay)
allr = [None] * ws
dist.all_gather_object(allr, res)
if rank == 0:
    sys.stdout.write(json.dumps(dict(n_gpus=ws, workload="C2", items=M,
                                     note="per_op_us: device time per launch, each op replayed 20x back to back from its own CUDA "
                                          "graph (warm L2); step_graph_replay_us: the whole step as one graph, median of 20",
                                     ranks=allr), indent=1) + "\n")
dist.destroy_process_group()
