"""-m gpu: the row-sharded path on >= 2 GPUs of one box (skipped on a 1-GPU box): one process per GPU, NCCL
allreduce of the per-run column sums inside the engine, results equal to the single-process oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_vb_matches_oracle(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
