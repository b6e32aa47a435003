"""Multi-GPU tests (need >= 2 GPUs on the box; skipped otherwise): the detection exchange over NCCL / NVLink."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script, n, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_detection_gather_over_nccl_two_ranks():
    """DetectionGather / gather_detections over NCCL and PeerExchange (the repo's own push / wait kernels over NVLink peer
    memory) return, on every rank, bit-identical rows and counts to what each rank produced, including ranks whose
    detections overflow the fixed-size message, slot reuse and a rank that falls behind."""
    r = _torchrun("mp_gather_worker.py", 2)
    assert r.returncode == 0 and "NCCL_GATHER_OK PEER_EXCHANGE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
