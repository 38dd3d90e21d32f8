"""Single-GPU checks of the multi-GPU plumbing's failure path: a peer barrier that gives up must be
reported (error bits through the C ABI, NativeError from the Python mirror) -- a sharded step that
pooled rows a peer had not written yet must never be returned silently."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import mre_b200  # noqa: F401
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _barrier(N, flag_ptrs, seq, rank, world, spins, dev):
    N.check(N.lib().pb200_peer_barrier_ex(N.ptr(flag_ptrs), N.ptr(seq[0:1]), rank, world, N.ptr(seq[1:2]),
                                          spins, N.stream_ptr(dev)), "peer_barrier_ex")


def test_peer_barrier_reports_the_peer_that_never_arrived(dev):
    from mre_b200 import _native as N
    world = 3
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    ptrs = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=dev)
    seqs = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
    # ranks 0 and 1 arrive (same process, same stream: rank 0 gives up waiting for 1 and 2, then rank 1
    # finds rank 0's flag already there and only misses rank 2)
    _barrier(N, ptrs, seqs[0], 0, world, 2000, dev)
    _barrier(N, ptrs, seqs[1], 1, world, 2000, dev)
    torch.cuda.synchronize()
    assert int(seqs[0][1]) == 0b110 and int(seqs[1][1]) == 0b100
    assert int(seqs[0][0]) == 1 and int(seqs[1][0]) == 1                  # sequence numbers advanced
    # everybody arrives: rank 2 first publishes, then 0 and 1 pass without a time-out (sequence 2 for
    # 0 and 1 needs rank 2 at >= 2: bring rank 2 to sequence 2 first)
    for s in seqs:
        s[1] = 0
    _barrier(N, ptrs, seqs[2], 2, world, 10, dev)      # seq 1: 0 and 1 are already at 1 -> passes
    _barrier(N, ptrs, seqs[2], 2, world, 10, dev)      # seq 2: gives up on 0 and 1 (still at 1) but publishes 2
    _barrier(N, ptrs, seqs[0], 0, world, 2000, dev)    # seq 2: 2 is there, 1 is not
    _barrier(N, ptrs, seqs[1], 1, world, 2000, dev)    # seq 2: everybody is there
    torch.cuda.synchronize()
    assert int(seqs[2][1]) == 0b011 and int(seqs[0][1]) == 0b010 and int(seqs[1][1]) == 0
    with pytest.raises(N.NativeError):
        N.check(N.lib().pb200_peer_barrier_ex(N.ptr(ptrs), N.ptr(seqs[0][0:1]), 0, world, N.ptr(seqs[0][1:2]), 0,
                                              N.stream_ptr(dev)), "peer_barrier_ex")


def test_peer_buffers_check_raises_after_a_timed_out_barrier(dev):
    """PeerBuffers.check() (called after every eager sharded step and by GraphedEmbeddings.replay) turns the
    device-side error flag into a NativeError naming the missing rank, and clears it."""
    from mre_b200 import _native as N, sharding as SH
    pb = SH.PeerBuffers.__new__(SH.PeerBuffers)            # a 2-rank view whose second rank never shows up
    pb.rank, pb.ws, pb.dev, pb.ok = 0, 2, dev, True
    pb._keep = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(2)]
    pb._flag_ptrs = torch.tensor([f.data_ptr() for f in pb._keep], dtype=torch.int64, device=dev)
    pb._seq = torch.zeros(2, dtype=torch.int32, device=dev)
    pb.max_spins = 1000
    pb.check()                                             # nothing issued yet: no sync, no error
    pb.barrier()
    with pytest.raises(N.NativeError, match=r"rank\(s\) \[1\]"):
        pb.check()
    assert pb.timed_out() == 0                             # cleared
    pb._keep[0][1] = 2                                     # the peer arrives for the next barrier
    pb.barrier()
    pb.check()
