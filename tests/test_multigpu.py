"""N>1 on real GPUs: the sharded scan (P2P mailbox exchange) must reproduce the single-GPU ordering and trace
bit for bit.  Spawns torchrun on 2 GPUs; skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

import fastneighbornet_b200 as fnn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(fnn.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_scan_equals_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mg_check.py"), "--quick"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("sharded == single: True") >= 6
