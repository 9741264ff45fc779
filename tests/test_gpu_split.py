"""Split mode on >= 2 GPUs: one halo shared by all ranks with one grouped NCCL all-reduce per pass must reproduce
the single-GPU result bit for bit (same default options: symmetric self-term, cached external sums, incremental
passes) and agree with the CPU oracle; includes the 2e6-star halo of BASELINE configs[3]."""
import os
import subprocess
import sys

import pytest

from pyhalma_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_split_mode_matches_single_gpu():
    n = _lib.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "split_check.py"),
           "--giant", "2000000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("identical_on_all_ranks=True") == 5 and out.stdout.count("oracle_parity=True") == 4
