#!/usr/bin/env python
"""Timeline of CTA 0 of the head_dim-128 dQ kernel (clock64 stamps per role, `make trace` build, SFA_LIB=...trace.so) at
the shortened C2 shape.  Roles: 0 producer (1 tile loads, 2 K slot free, 3 V slot free), 1 issuer (2 K landed, 5 S issued,
6 V landed, 7 dP issued, 3 S+dP committed, 4 dS ready, 8 dQ issued, 9 dQ committed), 2 math (1 waiting for S, 2 S + dP
complete, 3 dS written, 4 deferred epilogue done)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402
from sink_attention import _lib  # noqa: E402

B, N, Hq, Hkv, D, S, W = 1, 8192, 32, 8, 128, 4, 4096
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
dt = torch.bfloat16
mk = lambda H: torch.randn(B, H, N, D, device=dev, generator=g).to(dt)
q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
o, lse = sa.sink_flash_attention_with_lse(q, k, v, S, W, None)
lib = _lib.load()
for _ in range(2):
    _lib.bwd(q, k, v, o, do, lse, S, W, None)
torch.cuda.synchronize()
buf = torch.zeros(8 * 256 * 2, dtype=torch.int64, device=dev)
lib.sfa_set_trace_buffer(buf.data_ptr())
lib.sfa_set_bwd_stages(int(os.environ.get("TRACE_STAGE", "2")))
_lib.bwd(q, k, v, o, do, lse, S, W, None)
torch.cuda.synchronize()
lib.sfa_set_trace_buffer(None)
lib.sfa_set_bwd_stages(7)
t = buf.cpu().view(8, 256, 2)
ev = []
for role in range(8):
    for j in range(256):
        tag, clk = int(t[role, j, 0]), int(t[role, j, 1])
        if clk:
            ev.append((clk, role, tag >> 32, tag & 0xffffffff))
ev.sort()
t0 = ev[0][0]
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for clk, role, code, idx in ev:
    if lo <= idx < lo + n:
        print(f"{clk - t0:8d}  role {role} code {code} #{idx}")
print("total cycles CTA 0:", ev[-1][0] - t0, "events", len(ev))
