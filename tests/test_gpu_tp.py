"""Tensor-parallel decode (2 ranks, NVLink peer memory collectives) against the single-GPU path: needs 2 GPUs."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_tp2_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           str(ROOT / "tools" / "tp_check.py"), "tiny-gqa"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "-> OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
