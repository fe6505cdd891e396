#!/usr/bin/env python
"""UMMA issue cost: cycles per tcgen05.mma for one issuing thread (divergent lane-0 code vs warp-uniform code)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
from sink_attention import _probe as _lib  # noqa: E402

lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for uniform in (0, 2, 3, 4):
    for N in (16, 64, 144):
        for ksteps, reps in ((4, 16), (8, 16)):
            for _ in range(2):
                rc = lib.sfa_probe_mma_rate(out.data_ptr(), N, ksteps, reps, uniform, st)
                assert rc == 0
                torch.cuda.synchronize()
            n = ksteps * reps
            a, b = out.tolist()
            print(f"uniform={uniform} N={N:3d} mmas={n:3d}: issue {a / n:6.1f} cyc/mma, complete {b / n:6.1f} cyc/mma (ideal exec {128 * N / 256:.0f})")
