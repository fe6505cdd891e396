#!/usr/bin/env python
"""TMA / cp.async load-throughput probe: GB/s for streaming a bf16 [H, N, 64] tensor (larger than L2) into
shared memory, by box shape and by the number of boxes in flight per SM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
from sink_attention import _probe as _lib  # noqa: E402

lib = _lib.load()
H, N = 256, 8192
src = torch.randn(H, N, 64, device="cuda").to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
nbytes = src.numel() * 2
st = torch.cuda.current_stream().cuda_stream


def run(box_n, box_h, stages, grid, mode):
    ts = []
    for it in range(4):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.sfa_probe_tma_bw(src.data_ptr(), H, N, box_n, box_h, stages, grid, mode, st)
        b.record()
        b.synchronize()
        assert rc == 0, lib.sfa_last_error()
        ts.append(a.elapsed_time(b))
    t = min(ts[1:])
    return nbytes / (t * 1e-3) / 1e9, t


print(f"tensor {nbytes / 1e6:.0f} MB")
for mode, name in ((0, "TMA"), (1, "cp.async")):
    for box_n, box_h in ((128, 1), (16, 8), (64, 1), (256, 1)):
        if box_n * box_h * 128 > 32768:
            continue
        for stages in (1, 2, 4, 6):
            for grid in (148, 296):
                if stages * box_n * box_h * 128 * (2 if grid > 148 else 1) > 200 * 1024:
                    continue
                gbs, t = run(box_n, box_h, stages, grid, mode)
                print(f"{name:8s} box {box_n:3d}x{box_h} ({box_n * box_h * 128 // 1024:2d} KB) stages {stages} grid {grid}: {gbs:7.0f} GB/s  ({t * 1e3:.0f} us)")
