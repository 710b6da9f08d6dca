"""The library's own NCCL combine (hg_comm_init / hg_agg_combine, csrc/comm.cu) against the CPU oracle: world 1 on any GPU
box (exercises the REDUCE path end to end), world 2 when two GPUs are visible (tools/nccl_combine_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(cmd):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    return p.stdout


def test_combine_world1():
    out = _run([sys.executable, os.path.join(ROOT, "tools", "nccl_combine_check.py")])
    assert "rank 0/1: combine ok" in out


def test_combine_world2_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29517", os.path.join(ROOT, "tools", "nccl_combine_check.py")])
    assert "rank 0/2: combine ok" in out and "rank 1/2: combine ok" in out
