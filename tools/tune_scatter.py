#!/usr/bin/env python
"""NVLink throughput of sfa_ulysses_scatter by grid size and loads in flight (torchrun, one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
from sink_attention import _lib  # noqa: E402
from sink_attention.sp_utils import _P2PBuffers  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, n, Hq, Hkv, D = 1, 8192, 64, 8, 64
q = torch.randn(B, n, Hq, D, device=dev).to(torch.bfloat16)
bufs = _P2PBuffers(None, B, n, Hq, Hkv, D, torch.bfloat16, dev)


def ev_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps * 1e3


nbytes = q.numel() * 2
for blocks in (4, 8, 16, 32):
    for unroll in (1, 2, 4, 8):
        os.environ["SFA_SCATTER_BLOCKS"], os.environ["SFA_SCATTER_UNROLL"] = str(blocks), str(unroll)
        t = ev_time(lambda: _lib.ulysses_scatter(q, bufs.peer[0], rank, 0, bufs.tot, 0))
        if rank == 0:
            print(f"blocks/SM {blocks:2d} unroll {unroll}: {t:6.1f} us  {nbytes * (world - 1) / world / t / 1e3:6.0f} GB/s remote", flush=True)
dist.barrier()
dist.destroy_process_group()
