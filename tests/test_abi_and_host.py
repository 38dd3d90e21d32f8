"""CPU-only checks of the drop-in boundary and the host logic: the C-ABI library loads and
exports every symbol include/pinsage_b200.h declares, argument validation works without a
GPU, list padding follows the reference's semantics, and the product never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "movie-recommendation-engine_b200")


@pytest.fixture(scope="module")
def native():
    import mre_b200  # noqa: F401
    from mre_b200 import _native
    _native.build()
    return _native


def test_library_exports_every_declared_symbol(native):
    header = open(os.path.join(ROOT, "include", "pinsage_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    handle = ctypes.CDLL(native.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} not exported"
    assert native.lib().pb200_abi_version() == 1


def test_argument_validation_without_gpu(native):
    lib = native.lib()
    # invalid sizes are rejected before any CUDA call, with a message
    rc = lib.pb200_walk_topt(None, None, None, 0, 10, None, 5, 0, 2, 10, 1, 0, None, None, None,
                             None, None, None)
    assert rc == -1 and b"positive" in lib.pb200_last_error()
    rc = lib.pb200_walk_topt(None, None, None, 0, 10, None, 5, 40000, 2, 10, 1, 0, None, None, None,
                             None, None, None)
    assert rc == -1 and b"65535" in lib.pb200_last_error()
    rc = lib.pb200_lsh_encode(None, 4, 16, None, 100, None, None, None)
    assert rc == -1 and b"multiple of 32" in lib.pb200_last_error()
    rc = lib.pb200_topk(None, 4, None, 10, 8, 5000, 0, None, 0, None, None, None, 0, None)
    assert rc == -1
    assert lib.pb200_lsh_tables_workspace_bytes(1000, 32, 16) == 16 * 65536 * 4
    assert lib.pb200_lsh_tables_workspace_bytes(1000, 32, 5) == 0          # unsupported key width
    assert lib.pb200_csr_build_workspace_bytes(1000, 100) > 1000 * 24
    assert lib.pb200_topk_workspace_bytes(10, 1000, 16, 10) > 0
    # n == 0 is a no-op that needs neither pointers nor a device
    assert lib.pb200_walk_topt(None, None, None, 0, 10, None, 0, 100, 2, 10, 1, 0, None, None, None,
                               None, None, None) == 0
    assert lib.pb200_pool(None, 0, 8, None, None, None, None, 0, 4, 0, None, None) == 0


def test_no_cpu_fallback(native):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mre_b200.utils.random_walk import RandomWalkSampler
    from mre_b200.model.pinsage import PinSage, ImportancePooling
    from mre_b200.utils.nearest_neighbors import LSHIndex
    with pytest.raises(native.NativeError, match="no CPU fallback"):
        RandomWalkSampler(torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(native.NativeError, match="no CPU fallback"):
        ImportancePooling()(torch.zeros(3, 4), [[0]], [[1.0]])
    with pytest.raises(native.NativeError, match="no CPU fallback"):
        PinSage(4, 8, 4)(torch.zeros(3, 4))
    with pytest.raises(native.NativeError):
        LSHIndex(16)


def test_missing_library_fails_loudly(native, monkeypatch):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(native.NativeError, match="no CPU or PyTorch fallback"):
        native.lib()


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", re.M)
    for dirpath, _d, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} references the oracle"
    # and the reference tree is never read at run time by the product, bench or GPU tests
    for f in ["bench.py", "__graft_entry__.py", "tests/test_gpu_walk.py", "tests/test_gpu_model.py",
              "tests/test_gpu_search.py", "tests/test_gpu_c2.py"]:
        assert "/root/reference" not in open(os.path.join(ROOT, f)).read(), f


def test_state_dict_matches_reference_keys():
    import mre_b200  # noqa: F401
    from mre_b200.model.pinsage import PinSage
    keys = set(PinSage(128, 256, 128, 2).state_dict())
    want = {"input_proj.weight", "input_proj.bias", "output_proj.weight", "output_proj.bias"}
    for i in range(2):
        for l in ("lin_self", "lin_neigh", "lin_update"):
            want |= {f"convs.{i}.{l}.weight", f"convs.{i}.{l}.bias"}
    assert keys == want              # checkpoints/best_model.pt key set
    sd = PinSage(128, 256, 128, 2).state_dict()
    assert tuple(sd["convs.0.lin_update.weight"].shape) == (256, 512)
    assert tuple(sd["output_proj.weight"].shape) == (128, 256)


def test_pad_lists_semantics():
    import mre_b200  # noqa: F401
    from mre_b200 import neighbor_lists as NL
    nb = NL.pad_lists([[3, 1], [], [7, 8, 9], 5, [np.int64(2)]], [[0.5, 0.5], [], [1.0], 0.3, [2.0]],
                      "cpu", bare_int=True)
    assert nb.ids.tolist() == [[3, 1, -1], [-1, -1, -1], [7, 8, 9], [5, -1, -1], [2, -1, -1]]
    assert nb.list_len.tolist() == [2, 0, 3, 1, 1]
    assert nb.weight_len.tolist() == [2, 0, 1, 1, 1]          # short weight list kept short
    assert nb.weights[3].tolist() == [1.0, 0.0, 0.0]          # bare int -> weight 1 (pinsage.py:112)
    nb2 = NL.pad_lists([[1], [2]], [[1.0]], "cpu")            # zip semantics: min length
    assert len(nb2) == 1
    nb3 = NL.pad_lists([[2**40, 1]], None, "cpu")             # beyond int32: stays out of range
    assert nb3.ids[0, 0].item() == 2**31 - 1 and nb3.weights is None
    ids = torch.tensor([[4, 2, -1], [9, -1, -1]], dtype=torch.int32)
    cnt = torch.tensor([[3, 1, 0], [7, 0, 0]], dtype=torch.int32)
    n, w = NL.to_lists(ids, cnt, torch.tensor([2, 1], dtype=torch.int32))
    assert n == [[4, 2], [9]] and w == [[3 / 4, 1 / 4], [1.0]]


def test_synthetic_graph_layout():
    import mre_b200  # noqa: F401
    from mre_b200 import synthetic as S
    M, U, R = 200, 500, 5000
    ei, w = S.bipartite_graph(M, U, R, seed=1)
    assert ei.shape == (2, 2 * R) and ei.dtype == np.int64 and w.dtype == np.float32
    assert (ei[0, :R] >= M).all() and (ei[1, :R] < M).all()           # users -> items
    np.testing.assert_array_equal(ei[0, :R], ei[1, R:])               # reverse edges mirror
    np.testing.assert_array_equal(w[:R], w[R:])
    assert set(np.unique(w * 2)) <= set(range(1, 11))                  # half-star ratings
    assert len(np.unique(ei[0, :R] * M + ei[1, :R])) == R             # no duplicate pairs
    ei2, w2 = S.bipartite_graph(M, U, R, seed=1)
    np.testing.assert_array_equal(ei, ei2)


def test_walk_fork_rule_of_the_sharded_step(monkeypatch):
    """sharding._fork_walks: the walks run beside the input projection only for shards of at most 74 row tiles
    (measured: gain at 7,803 rows per rank, loss at 15,606 and 31,212); PB200_FORK_WALKS forces either way."""
    from mre_b200 import sharding as SH
    monkeypatch.delenv("PB200_FORK_WALKS", raising=False)
    assert SH._fork_walks(7803) and SH._fork_walks(74 * 128)
    assert not SH._fork_walks(74 * 128 + 1) and not SH._fork_walks(15606) and not SH._fork_walks(31212)
    monkeypatch.setenv("PB200_FORK_WALKS", "0")
    assert not SH._fork_walks(100)
    monkeypatch.setenv("PB200_FORK_WALKS", "1")
    assert SH._fork_walks(10 ** 6)
