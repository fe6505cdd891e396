#!/usr/bin/env python
"""Cycles per 16-element softmax step (FFMA + MUFU.EX2 + bf16 packing variants), by warps per scheduler."""
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
names = {0: "16 FFMA + 16 EX2", 1: "16 FFMA + 16 EX2 + 8 F2FP", 2: "16 FFMA + 16 EX2 + int pack", 3: "16 FFMA + 8 F2FP", 4: "16 FFMA"}
iters = 2000
for threads in (128, 256, 512):
    for mode in range(5):
        for _ in range(2):
            assert lib.sfa_probe_math_rate(out.data_ptr(), sink.data_ptr(), mode, iters, threads, st) == 0
            torch.cuda.synchronize()
        print(f"{threads // 128} warp(s)/scheduler  {names[mode]:30s}: {out.item() / iters:7.1f} cycles per step per warp")

print("with single-lane pollers (mbarrier.try_wait spin) on the same schedulers: 8 math warps + N poller warps")
for threads in (256, 384, 512, 768):
    for mode in (1,):
        for _ in range(2):
            assert lib.sfa_probe_math_rate(out.data_ptr(), sink.data_ptr(), 10 + mode, iters, threads, st) == 0
            torch.cuda.synchronize()
        print(f"2 math warps/scheduler + {(threads - 256) // 128} poller warp(s)/scheduler  {names[mode]:30s}: {out.item() / iters:7.1f} cycles per step per warp")

print("same step streamed from a larger code footprint (body unrolled U times, ~0.8 KB of SASS per copy):")
for threads in (128, 256):
    for U in (1, 8, 16, 64):
        for _ in range(2):
            assert lib.sfa_probe_math_rate(out.data_ptr(), sink.data_ptr(), 100 + U, 1984, threads, st) == 0
            torch.cuda.synchronize()
        print(f"{threads // 128} warp(s)/scheduler, unroll {U:2d}: {out.item() / 1984:7.1f} cycles per step per warp")
