"""ctypes binding of libpinsage_b200.so (include/pinsage_b200.h).

PyTorch is used for device memory and streams only: every entry point receives raw device
pointers (``tensor.data_ptr()``), sizes and the current CUDA stream handle.  There is no CPU
or PyTorch fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libpinsage_b200.so")
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

c_i32, c_i64, c_u32, c_u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
c_int, c_size, c_ptr, c_float = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float

# name -> (restype, argtypes); every symbol declared in include/pinsage_b200.h
SIGNATURES = {
    "pb200_abi_version": (c_int, []),
    "pb200_last_error": (ctypes.c_char_p, []),
    "pb200_launch_count": (c_i64, []),
    "pb200_set_l2_fetch_granularity": (c_int, [c_int]),
    "pb200_get_l2_fetch_granularity": (c_int, []),
    "pb200_edge_weight_probe": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "pb200_csr_build_workspace_bytes": (c_size, [c_i64, c_i64]),
    "pb200_csr_build": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                c_ptr, c_size, c_ptr]),
    "pb200_walk_topt": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr, c_i64, c_int, c_int,
                                c_int, c_u64, c_u32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pb200_walk_index_workspace_bytes": (c_size, [c_i64]),
    "pb200_walk_index_sizes": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_size, c_ptr]),
    "pb200_walk_index_build": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                       c_ptr]),
    "pb200_walk_topt_indexed": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_int, c_int,
                                        c_int, c_u64, c_u32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                        c_ptr]),
    "pb200_walk_index_leaf_range": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "pb200_walk_index_build_ex": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_int,
                                          c_ptr]),
    "pb200_walk_bucket_workspace_bytes": (c_size, [c_i64]),
    "pb200_walk_bucket_plan": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "pb200_walk_bucket_fill": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_u64, c_ptr]),
    "pb200_walk_bucket_plan_ex": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_size, c_int, c_ptr]),
    "pb200_walk_bucket_fill_ex": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_u64, c_int, c_ptr]),
    "pb200_walk_topt_indexed_ex": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr, c_i64, c_int, c_int,
                                           c_int, c_u64, c_u32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                           c_ptr]),
    "pb200_walk_topt_indexed_multi": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr, c_i64, c_int, c_int,
                                              c_int, c_u64, c_u32, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                              c_ptr]),
    "pb200_u32_add": (c_int, [c_ptr, c_u32, c_ptr]),
    "pb200_count_topt": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pb200_cooc_pairs": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_int,
                                 c_ptr, c_ptr, c_ptr, c_ptr, c_u64, c_ptr, c_ptr]),
    "pb200_cooc_edges_workspace_bytes": (c_size, [c_i64]),
    "pb200_cooc_edges": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "pb200_ppr_push": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_i64, c_i64, c_ptr, c_i64, ctypes.c_double, c_int,
                               c_ptr, c_ptr, c_ptr]),
    "pb200_topk_rows_f64": (c_int, [c_ptr, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr]),
    "pb200_pool": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_int,
                           c_ptr, c_ptr]),
    "pb200_pool_sharded": (c_int, [c_ptr, c_int, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                   c_int, c_int, c_ptr, c_ptr]),
    "pb200_pool_sharded_ex": (c_int, [c_ptr, c_int, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                                      c_int, c_int, c_int, c_ptr, c_ptr]),
    "pb200_peer_alloc": (c_int, [c_size, c_ptr]),
    "pb200_peer_free": (c_int, [c_ptr]),
    "pb200_peer_export": (c_int, [c_ptr, c_ptr]),
    "pb200_peer_open": (c_int, [c_ptr, c_ptr]),
    "pb200_peer_close": (c_int, [c_ptr]),
    "pb200_peer_barrier": (c_int, [c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr]),
    "pb200_peer_barrier_ex": (c_int, [c_ptr, c_ptr, c_int, c_int, c_ptr, c_u64, c_ptr]),
    "pb200_gather_dense": (c_int, [c_ptr, c_int, c_ptr, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr,
                                   c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int,
                                   c_int, c_int, c_ptr, c_ptr]),
    "pb200_round_tf32": (c_int, [c_ptr, c_ptr, c_i64, c_ptr]),
    "pb200_topk_workspace_bytes": (c_size, [c_i64, c_i64, c_int, c_int]),
    "pb200_topk": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i32, c_ptr,
                           c_ptr, c_ptr, c_size, c_ptr]),
    "pb200_topk_tc_supported": (c_int, [c_i64, c_i64, c_int, c_int, c_int]),
    "pb200_topk_tc_workspace_bytes": (c_size, [c_i64, c_i64, c_int, c_int]),
    "pb200_topk_tc": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i32, c_ptr,
                              c_ptr, c_ptr, c_size, c_ptr, c_ptr]),
    "pb200_rank_of_target_workspace_bytes": (c_size, [c_i64]),
    "pb200_rank_of_target": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_size, c_ptr]),
    "pb200_topk_merge": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "pb200_lsh_encode": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "pb200_hamming_topk": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_i32, c_ptr, c_ptr,
                                   c_ptr]),
    "pb200_lsh_encode_tc_workspace_bytes": (c_size, [c_i64, c_int, c_int]),
    "pb200_lsh_encode_tc": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_ptr, c_size, c_ptr]),
    "pb200_hamming_topk_tc_supported": (c_int, [c_i64, c_i64, c_int, c_int]),
    "pb200_hamming_topk_tc_workspace_bytes": (c_size, [c_i64, c_i64, c_int, c_int]),
    "pb200_hamming_topk_tc": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_i32, c_ptr, c_ptr, c_ptr,
                                      c_size, c_ptr]),
    "pb200_ivf_search_tc_supported": (c_int, [c_i64, c_i64, c_int, c_int, c_int]),
    "pb200_ivf_search_tc_workspace_bytes": (c_size, [c_i64, c_i64, c_int, c_int, c_int, c_int]),
    "pb200_ivf_search_tc": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                    c_ptr, c_ptr, c_i64, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_size, c_ptr,
                                    c_ptr]),
    "pb200_lsh_tables_workspace_bytes": (c_size, [c_i64, c_int, c_int]),
    "pb200_lsh_build_tables": (c_int, [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_size,
                                       c_ptr]),
    "pb200_lsh_search_tables": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr,
                                        c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pb200_ivf_build_workspace_bytes": (c_size, [c_i64, c_int]),
    "pb200_ivf_build": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                c_size, c_ptr]),
    "pb200_ivf_centroid_update": (c_int, [c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr]),
    "pb200_ivf_search_ex": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr,
                                    c_ptr, c_ptr, c_ptr]),
    "pb200_lsh_search_tables_ex": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                           c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "pb200_ivf_search": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int,
                                 c_ptr, c_ptr, c_ptr]),
}

POOL_PINSAGE, POOL_LAYERS, POOL_AGGREGATOR, POOL_MEAN, POOL_MAX = range(5)
EPI_RELU, EPI_L2NORM, EPI_LAYERNORM, EPI_ROUND_TF32, IN_A1_TF32, IN_A2_TF32 = 1, 2, 4, 8, 16, 32
POOL_ROUND_TF32 = 0x100
PREC_FP32, PREC_TF32, PREC_AUTO = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "auto": PREC_AUTO}
METRIC_IP, METRIC_L2 = 0, 1
SHARD_BLOCKS, SHARD_CYCLIC = 0, 1
LEAF_WIDE, LEAF_COMPACT, LEAF_BUCKET, LEAF_BUCKET32 = 0, 1, 2, 3          # sampling-index leaf formats (pb200_walk_index_build_ex)

_lib = None


class NativeError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu into libpinsage_b200.so for sm_100a (nvcc cross-compiles w/o GPU)."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise NativeError("building libpinsage_b200.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded library; raises loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `make -C {CSRC_DIR}` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. This package has no CPU "
                "or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the ABI is incomplete
            fn.restype, fn.argtypes = res, args
        if handle.pb200_abi_version() != 1:
            raise NativeError("libpinsage_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().pb200_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'libpinsage_b200'} failed (status {rc}): {msg}")


def launch_count():
    return int(lib().pb200_launch_count())


def device_of(*tensors, device=None):
    """The CUDA device to run on; raises when there is none (no CPU fallback)."""
    if device is not None:
        dev = torch.device(device)
    else:
        dev = next((t.device for t in tensors if isinstance(t, torch.Tensor) and t.is_cuda), None)
        if dev is None:
            if not torch.cuda.is_available():
                raise NativeError("a CUDA device (B200, sm_100a) is required: this package has no "
                                  "CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise NativeError(f"device {dev} is not a CUDA device: this package has no CPU fallback")
    return dev


def dev_tensor(t, dtype, device):
    """Contiguous tensor of `dtype` on `device` (H2D copy if needed; plumbing only)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native ops need contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
