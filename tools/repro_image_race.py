#!/usr/bin/env python
"""Single-GPU reproduction of the round-1 nondeterminism in the fused backward (dQ and dK differing between two runs
on the same inputs, dV / O / ds_aux exact), and the check that the fix holds.

Root cause: in bwd_fused64_kernel the math warps read P back from the shared-memory P image in pass 2 of tile n and
write P(n + 1) into it in pass 1 of tile n + 1.  The image is in ring-column order, which shifts by one key block per
tile, so the cell [row r, ring chunk c] written by the warp of part p for tile n + 1 is the cell the warp of part
p + 1 (same lane quarter) reads in pass 2 of tile n.  Only the tensor pipe's reads were ordered (p_free); a warp that
ran a full pass ahead of its neighbour made dS(n) = P(n + 1) o (dP - delta): dQ and dK wrong, dV untouched.  On one
GPU the warps stay close enough that it (almost) never fires; routed dQ stores over NVLink shifted the timing.

Knob 1 = 1 removes the ordering barrier again (the round-1 kernel), knob 0 delays a third of the math warps.
Prints how many elements differ from the undisturbed run for {barrier on, off} x {delays}.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402
from sink_attention import _lib  # noqa: E402

B, Hq, Hkv, N, W, D = 1, 64, 8, 8192, 128, 64
g = torch.Generator(device="cuda").manual_seed(7)
mk = lambda H: torch.randn(B, H, N, D, device="cuda", generator=g).to(torch.bfloat16)
q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
s_aux = torch.randn(Hq, device="cuda", generator=g) * 0.5
o, lse = sa.sink_flash_attention_with_lse(q, k, v, 0, W, s_aux)
base = _lib.bwd(q, k, v, o, do, lse, 0, W, s_aux)
assert _lib.last_impl() == "tcgen05-fused"
torch.cuda.synchronize()
names = ("dq", "dk", "dv", "ds_aux")
print(f"shape B={B} N={N} Hq={Hq} Hkv={Hkv} D={D} W={W}; baseline = undisturbed run with the ordering barrier")
for norace in (1, 0):
    for delay in (0, 500, 2000, 10000):
        _lib.set_debug(1, norace)
        _lib.set_debug(0, delay)
        worst = {n: 0 for n in names}
        nbad = {n: 0 for n in names}
        for _ in range(4):
            r = _lib.bwd(q, k, v, o, do, lse, 0, W, s_aux)
            torch.cuda.synchronize()
            for n, a, b in zip(names, r, base):
                d = (a.float() - b.float()).abs()
                nbad[n] = max(nbad[n], int((d > 0).sum()))
                worst[n] = max(worst[n], float(d.max()))
        tag = "barrier OFF (round-1 kernel)" if norace else "barrier ON  (fixed kernel)  "
        print(f"{tag} delay {delay:6d} ns: differing elements " +
              " ".join(f"{n}={nbad[n]}" for n in names) + "  max |diff| " + " ".join(f"{n}={worst[n]:.3g}" for n in names))
_lib.set_debug(0, 0)
_lib.set_debug(1, 0)
