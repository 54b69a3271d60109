"""Multi-GPU parity (needs >= 2 B200 on the box; skipped on a 1-GPU box): rows of the kNN stage sharded over
2 ranks with the two NCCL all-gathers, and resamples strided over 2 ranks with the NCCL reduce, must
reproduce the single-rank result (tools/mgpu_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharding_matches_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "coclust: sharded" in r.stdout and "identical True" in r.stdout
