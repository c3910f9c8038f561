"""Multi-GPU data-parallel parity (needs >= 2 GPUs; skipped on a single-GPU box): launches scripts/dp_parity.py under
torchrun over NCCL -- fp32 validation mode: 2 ranks x batch 1 == batch 2 single process <= 1e-5; bf16 product path with
the overlapped bucketed all-reduce: differently seeded ranks end up bit-identical after 3 steps."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_data_parallel_parity():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_parity.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "DP parity fp32 mode" in res.stdout and "identical across 2 differently seeded ranks" in res.stdout
