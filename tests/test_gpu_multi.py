"""Multi-GPU parity (needs >= 2 B200 on the box; the N > 1 cases are skipped on a 1-GPU box).  NCCL runs INSIDE the
library (SURVEY 8b "Threading", 8e): (1) one process driving several devices (mcb_multi: mcb_cgraph_bknn_multi,
mcb_coclust_multi), (2) one process per device (mcb_comm: mcb_cgraph_build_sharded, mcb_cgraph_gather_edges,
mcb_coclust_reduce).  Both must reproduce the single-device result: edges identical up to correlation ties, co-occurrence
counts bit-exact.  tools/multi_check.py is the worker (the launcher's gloo group only carries the NCCL id)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run(cmd, timeout=900):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    return r


def test_single_process_session_matches_single_device():
    """runs with 1 GPU too: a one-device session exercises the worker threads, the shared result table and the rank code"""
    r = _run([sys.executable, os.path.join(ROOT, "tools", "multi_check.py")])
    assert r.returncode == 0 and "MULTI_CHECK PASS" in r.stdout


def test_process_per_device_matches_single_device():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + os.getpid() % 400
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", str(port), os.path.join(ROOT, "tools", "multi_check.py")])
    assert r.returncode == 0 and "MULTI_CHECK PASS" in r.stdout
