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


def test_second_device_in_one_process():
    """One process, two GPUs (HF device_map="auto", the reference's subprocess_generate retry ladder,
    subprocess_eval.py:163-198): the per-device launch state (dynamic shared memory attribute, SM count) must be set
    up on EVERY device the library is used on -- round 1 kept it in per-process statics and failed on the second."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import sink_attention as sa
    from sink_attention import _lib
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        g = torch.Generator().manual_seed(0)
        mk = lambda H, N=384: torch.randn(1, H, N, 64, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
        q, k, v = mk(8), mk(1), mk(1)
        s_aux = torch.zeros(8, device=dev, requires_grad=True)
        o = sa.sink_flash_attention(q, k, v, 0, 128, s_aux)
        assert _lib.last_impl() == "tcgen05"
        o.float().square().sum().backward()
        assert _lib.last_impl() == "tcgen05-fused"
        qd = torch.randn(2, 8, 1, 64, generator=g).to(dev, torch.bfloat16)
        kd = torch.randn(2, 1, 500, 64, generator=g).to(dev, torch.bfloat16)
        od = sa.sink_decode_attention(qd, kd, kd, s_aux.detach())
        torch.cuda.synchronize(dev)
        outs.append((o.detach().cpu(), q.grad.cpu(), k.grad.cpu(), od.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b)
