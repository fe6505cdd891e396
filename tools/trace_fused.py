#!/usr/bin/env python
"""Timeline of CTA 0 of the fused backward kernel (clock64 stamps per role) at the C1 shape.
Build first with `make -C sink-flash-attention-kernel_b200 clean all EXTRA=-DSFA_TRACE=1`."""
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
buf = torch.zeros(9 * 256 * 2 + 1, dtype=torch.int64, device=dev)
buf[-1] = int(os.environ.get('TRACE_CTA', '0'))
lib.sfa_set_trace_buffer(buf.data_ptr())
lib.sfa_set_bwd_stages(6)
_lib.bwd(q, k, v, o, do, lse, S, W, s_aux)
torch.cuda.synchronize()
lib.sfa_set_trace_buffer(None)
lib.sfa_set_bwd_stages(7)
t = buf.cpu()[:-1].view(9, 256, 2)
ROLES = ["PROD", "ISS_A", "ISS_V", "MATH", "EPI0", "ISS_K", "EPI1", "ISS_Q", "DELTA"]
CODES = [
    {1: "begin", 3: "loads issued"},
    {1: "begin", 2: "S inputs ready", 3: "S issued", 4: "dP inputs ready", 5: "dP issued"},
    {1: "begin", 2: "P ready + drain ok", 3: "dV issued"},
    {1: "begin", 2: "S full + P free", 3: "pass 1 done", 4: "dP full + dS free", 5: "pass 2 done", 6: "delta read"},
    {1: "begin", 2: "tile UMMAs complete", 3: "dQ loaded", 4: "dQ stored", 5: "drain done", 6: "delta done", 7: "delta begin", 8: "delta inputs ready"},
    {1: "begin", 2: "dS ready + drain ok", 3: "dK issued"},
    {1: "begin", 2: "tile UMMAs complete", 3: "dQ loaded", 4: "dQ stored", 5: "drain done", 6: "delta done", 7: "delta begin", 8: "delta inputs ready"},
    {1: "begin", 2: "dS ready + dq_free ok", 3: "dQ issued"},
    {1: "begin", 2: "O landed", 3: "dO landed", 4: "delta slot free", 5: "delta done"},
]
ev = []
for role in range(9):
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
        print(f"{clk - t0:8d}  {ROLES[role]:5s} {CODES[role].get(code, str(code)):20s} #{idx}")
print("total cycles CTA 0:", ev[-1][0] - t0, "events", len(ev))
