"""Driver-visible parity of the sharded path on real GPUs: launches tools/check_ulysses_p2p.py under torchrun on two
GPUs (peer-memory Ulysses exchange with routed O / dQ vs the NCCL path, vs the un-sharded operator, multi-round with
skewed ranks, fwd-fwd-bwd-bwd).  Skips on boxes with fewer than two GPUs."""
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


@pytest.mark.timeout(600)
def test_ulysses_p2p_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_ulysses_p2p.py"), "--n-local", "2048",
           "--rounds", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert '"ok_all_ranks": true' in r.stdout
