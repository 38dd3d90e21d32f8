"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): torchrun --nproc-per-node 2 runs
tools/multi_gpu_check.py, which asserts that row-sharded get_embeddings and item-/query-sharded
search over NCCL give bitwise the same results as the single-GPU path."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_results_identical_to_one_gpu():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_OK" in r.stdout
