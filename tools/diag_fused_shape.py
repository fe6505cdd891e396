#!/usr/bin/env python
"""Diagnostic: the fused backward at one shape against the CUDA-core kernels, with and without the pipeline-delay knob.
  python tools/diag_fused_shape.py B Hq Hkv N W hf [delay ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa
from sink_attention import _lib


def excess(got, ref, atol, rtol):
    g, r = got.float(), ref.float()
    return float(((g - r).abs() / (atol + rtol * r.abs())).max())

B, Hq, Hkv, N, W, hf = (int(x) for x in sys.argv[1:7])
delays = [int(x) for x in sys.argv[7:]] or [0, 2000]
D, S, dtype = 64, 0, torch.bfloat16
g = torch.Generator().manual_seed(B * 5 + Hq * 11 + N * 3 + W)
def mk(H):
    if hf:
        return torch.randn(B, N, H, D, generator=g).to("cuda", dtype).transpose(1, 2)
    return torch.randn(B, H, N, D, generator=g).to("cuda", dtype)
q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
s_aux = (torch.randn(Hq, generator=g) * 0.5 + 0.5).cuda()
o, lse = sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)
_lib.set_impl(_lib.IMPL_SIMT)
ref = _lib.bwd(q, k, v, o, do, lse, S, W, s_aux)
_lib.set_impl(_lib.IMPL_AUTO)
torch.cuda.synchronize()
for d in delays:
    _lib.set_debug(0, d)
    for rep in range(3):
        out = _lib.bwd(q, k, v, o, do, lse, S, W, s_aux)
        torch.cuda.synchronize()
        ex = [round(float(excess(a, b, 2e-2, 1e-2)), 3) for a, b in zip(out[:3], ref[:3])]
        bad = []
        for name, a, b in zip(("dq", "dk", "dv"), out[:3], ref[:3]):
            diff = (a.float() - b.float()).abs()
            idx = (diff > 2e-2 + 1e-2 * b.float().abs()).nonzero()
            if len(idx):
                bad.append((name, len(idx), idx[0].tolist(), idx[-1].tolist()))
        print(f"delay {d} rep {rep}: impl {_lib.last_impl()} excess dq/dk/dv {ex} ds_aux diff {float((out[3]-ref[3]).abs().max()):.2e} bad {bad}")
    _lib.set_debug(0, 0)
