#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent forward kernel (clock64 stamps per role) at the C1 shape.
Build the timeline library first (`make -C sink-flash-attention-kernel_b200 trace`) and run with
SFA_LIB=.../libsinkfa_trace.so."""
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
s_aux = torch.randn(Hq, device=dev, generator=g)
lib = _lib.load()
for _ in range(3):
    sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)
torch.cuda.synchronize()
buf = torch.zeros(8 * 256 * 2, dtype=torch.int64, device=dev)
lib.sfa_set_trace_buffer(buf.data_ptr())
sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)
torch.cuda.synchronize()
lib.sfa_set_trace_buffer(None)
t = buf.cpu().view(8, 256, 2)
ROLES = {1: "ISS_S", 3: "ISS_PV", 4: "SOFT", 6: "EPI"}
CODES = {
    1: {1: "begin", 2: "inputs ready", 3: "S issued"},
    3: {1: "begin", 2: "P + V ready", 3: "PV issued"},
    4: {1: "begin", 2: "S full", 5: "max pass done", 6: "max stored", 4: "max exchanged", 7: "exp pass done", 3: "P arrived"},
    6: {1: "begin", 2: "O complete", 3: "store issued"},
}
ev = []
for role in ROLES:
    for j in range(256):
        tag, clk = int(t[role, j, 0]), int(t[role, j, 1])
        if clk == 0:
            continue
        ev.append((clk, role, tag >> 32, tag & 0xffffffff))
ev.sort()
t0 = ev[0][0]
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for clk, role, code, idx in ev:
    if lo <= idx < lo + n:
        print(f"{clk - t0:8d}  {ROLES[role]:6s} {CODES[role].get(code, str(code)):16s} #{idx}")
print("total cycles CTA 0:", ev[-1][0] - t0, "events", len(ev))
