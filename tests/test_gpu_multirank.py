"""Two ranks on two GPUs of one box (skipped on a single-GPU box): tools/multigpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_library_nccl_and_torch_allreduce():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTIGPU_OK world=2" in out.stdout


def test_single_process_group_over_all_gpus():
    """sdgpu_group_*: one host thread drives every visible GPU (also meaningful with one GPU: the group degenerates to a context)"""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "group_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "GROUP_OK" in out.stdout
