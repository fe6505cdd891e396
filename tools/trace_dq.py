#!/usr/bin/env python
"""Timeline of CTA 0 of the dQ kernel (clock64 stamps per role) at the C1 shape: where the time per tile goes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402
from sink_attention import _lib  # noqa: E402

B, N, Hq, Hkv, D, S, W = 1, 8192, 64, 8, 64, 0, 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
dt = torch.bfloat16
q = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)
k = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt)
v = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt)
do = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)
s_aux = torch.randn(Hq, device=dev, generator=g)
o, lse = sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)
lib = _lib.load()
for _ in range(3):
    _lib.bwd(q, k, v, o, do, lse, S, W, s_aux)
torch.cuda.synchronize()
buf = torch.zeros(8 * 256 * 2, dtype=torch.int64, device=dev)
lib.sfa_set_trace_buffer(buf.data_ptr())
stage = int(os.environ.get('TRACE_STAGE', '2'))
if stage == 0:
    sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)
else:
    lib.sfa_set_bwd_stages(stage)
    _lib.bwd(q, k, v, o, do, lse, S, W, s_aux)
torch.cuda.synchronize()
lib.sfa_set_trace_buffer(None)
lib.sfa_set_bwd_stages(7)
t = buf.cpu().view(8, 256, 2)
ROLES = ["PROD", "I_S", "I_dP", "I_dQ", "EXP", "DS", "EPI", "-"]
if os.environ.get("TRACE_STAGE") == "0":
    ROLES = ["PROD", "I_S", "-", "I_PV", "SOFTMAX", "-", "EPI", "-"]
if os.environ.get("TRACE_STAGE") == "4":
    ROLES = ["PROD", "B:S^T", "B:dP^T", "A:dVdK", "MATH", "-", "-", "-"]
CODES = {1: "begin / wait", 2: "inputs ready", 3: "done", 4: "P ready / dep ok", 5: "batch loaded / bar", 6: "batch computed / ph1 done", 7: "st waited"}
ev = []
for role in range(8):
    for j in range(256):
        tag, clk = int(t[role, j, 0]), int(t[role, j, 1])
        if clk == 0:
            continue
        ev.append((clk, role, tag >> 32, tag & 0xffffffff))
ev.sort()
t0 = ev[0][0]
if os.environ.get('TRACE_WARPS'):
    for clk, role, code, idx in ev:
        if role == 5:
            print(f"{clk - t0:8d}  warp {idx % 100:2d} item {idx // 100}  {['S ready', 'pass1 done', 'max exchanged', 'arrived'][code - 1]}")
    sys.exit(0)
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for clk, role, code, idx in ev:
    if lo <= idx < lo + 2 and role == (int(sys.argv[2]) if len(sys.argv) > 2 else role):
        print(f"{clk - t0:8d}  {ROLES[role]:5s} {CODES[code]:14s} #{idx}")
