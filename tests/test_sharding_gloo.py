"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard layout, padded row
all-gather, per-shard candidate-list gather + total-order merge == unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mre_b200  # noqa: F401
    from mre_b200 import sharding as SH
    from oracle import oracle as O
    try:
        # --- row shards and the per-layer all-gather of h (uneven tail shard) ---
        lo, hi = SH.shard_range(n_items, rank, world)
        full = torch.arange(n_items * 3, dtype=torch.float32).reshape(n_items, 3)
        got = SH.all_gather_rows(full[lo:hi].clone(), n_items)
        assert torch.equal(got, full)
        # --- cyclic layout (rows dealt round-robin: the embedding step's default) ---
        assert SH.EMB_LAYOUT == "cyclic"
        mine = SH.local_rows(full)
        assert torch.equal(mine, full[rank::world]) and mine.size(0) <= SH.shard_size(n_items, world)
        got = SH.all_gather_rows(mine.clone(), n_items, layout="cyclic")
        assert torch.equal(got, full)
        # --- item-sharded exact search: local top-k with global ids, gather, merge ---
        rng = np.random.Generator(np.random.PCG64(0))
        x = rng.standard_normal((n_items, 16)).astype(np.float32)
        q = rng.standard_normal((9, 16)).astype(np.float32)
        s_loc, i_loc = O.exact_ip(x[lo:hi], q, 5)
        i_loc = i_loc + lo
        s_all = SH.all_gather_cols(torch.from_numpy(s_loc))
        i_all = SH.all_gather_cols(torch.from_numpy(i_loc.astype(np.int32)))
        assert s_all.shape == (9, 5 * world)
        ms, mi = O.topk_merge([s_all.numpy()], [i_all.numpy()], 5, largest=True)
        rs, ri = O.exact_ip(x, q, 5)
        np.testing.assert_array_equal(mi, ri)
        np.testing.assert_array_equal(ms, rs)
        # --- query-sharded search: results come back in query order on every rank ---
        def search(q_local):
            s, i = O.exact_ip(x, q_local.numpy(), 5)
            return torch.from_numpy(s), torch.from_numpy(i.astype(np.int32))
        s2, i2 = SH.search_query_sharded(search, torch.from_numpy(q))
        np.testing.assert_array_equal(i2.numpy(), ri)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [10, 11])
def test_sharding_world2_gloo(n_items):
    mp.spawn(_worker, args=(2, _free_port(), n_items), nprocs=2, join=True)


def test_shard_ranges_cover_everything():
    import mre_b200  # noqa: F401
    from mre_b200 import sharding as SH
    for n in (0, 1, 7, 8, 62423):
        for ws in (1, 2, 4, 8):
            spans = [SH.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= SH.shard_size(n, ws)
