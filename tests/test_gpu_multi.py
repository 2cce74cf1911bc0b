"""GPU, >= 2 devices: cell-sharded optimize_ALS / online_iNMF over NCCL must reproduce the oracle."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_parity():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "tools", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0, res.stderr[-3000:]
    assert "MULTIGPU CHECK OK" in res.stdout
