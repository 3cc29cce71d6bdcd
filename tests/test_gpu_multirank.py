"""Launches tests/multigpu_check.py under torchrun on every visible GPU (needs >= 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multigpu_nccl_and_peer_window_agree_with_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0 and "MULTIGPU CHECK PASSED" in p.stdout
