#!/usr/bin/env python
"""TMEM -> register read throughput per SM by number of warps and tcgen05.ld width."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
from sink_attention import _probe as _lib  # noqa: E402

lib = _lib.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
sink = torch.zeros(1024, device="cuda")
st = torch.cuda.current_stream().cuda_stream
cols = {0: 16, 1: 32, 2: 64}
iters = 2000
for threads in (32, 128, 256, 512):
    for mode in range(3):
        for _ in range(2):
            assert lib.sfa_probe_tmem_rate(out.data_ptr(), sink.data_ptr(), mode, iters, threads, st) == 0
            torch.cuda.synchronize()
        cyc = out.item() / iters
        bytes_per_iter = threads * cols[mode] * 4
        print(f"{threads // 32:2d} warps, {cols[mode]:2d} columns per round trip: {cyc:7.1f} cycles per round trip, {bytes_per_iter / cyc:7.1f} B/clk/SM")
