#!/usr/bin/env python
"""Small workload for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool racecheck  python tools/sanitize_target.py [--norace-barrier-off]
    compute-sanitizer --tool synccheck  python tools/sanitize_target.py
Runs the head_dim-64 tcgen05 forward, the fused backward (two CTAs with a shared key block, a sequence boundary inside
a CTA run) and the decode kernel on shapes small enough for the instrumented run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402
from sink_attention import _lib  # noqa: E402

if "--norace-barrier-off" in sys.argv:
    _lib.set_debug(1, 1)          # the round-1 kernel: no ordering barrier on the P image
B, Hq, Hkv, N, D, W = 2, 8, 1, 352, 64, 128
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda H: torch.randn(B, H, N, D, device="cuda", generator=g).to(torch.bfloat16)
q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
s_aux = torch.randn(Hq, device="cuda", generator=g) * 0.5
o, lse = sa.sink_flash_attention_with_lse(q, k, v, 0, W, s_aux)
print("fwd:", _lib.last_impl())
dq, dk, dv, ds = _lib.bwd(q, k, v, o, do, lse, 0, W, s_aux)
print("bwd:", _lib.last_impl())
qq = torch.randn(2, Hq, 1, D, device="cuda", generator=g).to(torch.bfloat16)
kk = torch.randn(2, Hkv, 300, D, device="cuda", generator=g).to(torch.bfloat16)
od = sa.sink_decode_attention(qq, kk, kk, s_aux)
print("decode:", _lib.last_impl())
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(o.float()).all() and torch.isfinite(dq.float()).all() and torch.isfinite(od.float()).all()))
