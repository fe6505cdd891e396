#!/usr/bin/env python
"""Localises differences between the fused backward and the CUDA-core path under the extended geometry."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa
from sink_attention import _lib

def run(B, Hq, Hkv, D, W, Nq, Nkv):
    g = torch.Generator().manual_seed(1)
    dt = torch.bfloat16
    q = torch.randn(B, Hq, Nq, D, generator=g).to("cuda", dt)
    k = torch.randn(B, Hkv, Nkv, D, generator=g).to("cuda", dt)
    v = torch.randn(B, Hkv, Nkv, D, generator=g).to("cuda", dt)
    do = torch.randn(B, Hq, Nq, D, generator=g).to("cuda", dt)
    s = (torch.randn(Hq, generator=g) * 0.5).cuda()
    ext = _lib.make_ext(Nq, Nkv, Nkv - Nq)
    o, lse = _lib.fwd(q, k, v, 0, W, s, ext=ext)
    f = _lib.bwd(q, k, v, o, do, lse, 0, W, s, ext=ext); nf = _lib.last_impl()
    _lib.set_impl(_lib.IMPL_SIMT)
    r = _lib.bwd(q, k, v, o, do, lse, 0, W, s, ext=ext); nr = _lib.last_impl()
    _lib.set_impl(_lib.IMPL_AUTO)
    torch.cuda.synchronize()
    if Nkv <= 1024:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import sink_oracle as orc
        off = Nkv - Nq
        qf = torch.zeros(B, Hq, Nkv, D); qf[:, :, off:] = q.float().cpu()
        dof = torch.zeros(B, Hq, Nkv, D); dof[:, :, off:] = do.float().cpu()
        o_r, lse_r = orc.sink_attention_fwd(qf, k.float().cpu(), v.float().cpu(), 0, W, s.cpu())
        g_r = orc.sink_attention_bwd(qf, k.float().cpu(), v.float().cpu(), dof, 0, W, s.cpu())
        print("  vs oracle: o", (o.float().cpu() - o_r[:, :, off:]).abs().max().item(), "lse", (lse.cpu() - lse_r[:, :, off:]).abs().max().item())
        for name, a, b, c in zip(("dq", "dk", "dv"), f, r, (g_r[0][:, :, off:], g_r[1], g_r[2])):
            print(f"  {name}: fused-oracle {(a.float().cpu() - c).abs().max().item():.4f}  simt-oracle {(b.float().cpu() - c).abs().max().item():.4f}")
    print(f"B={B} Hq={Hq} Hkv={Hkv} W={W} Nq={Nq} Nkv={Nkv} q_off={Nkv-Nq}: {nf} vs {nr}")
    for name, a, b in zip(("dq", "dk", "dv"), f, r):
        d = (a.float() - b.float()).abs()
        bad = (d > 0.05 + 0.05 * b.float().abs())
        idx = bad.nonzero()
        pos = sorted(set(idx[:, 2].tolist()))
        print(f"  {name}: max|diff| {d.max().item():.4f}  bad {int(bad.sum())}  positions {pos[:12]}{'...' if len(pos) > 12 else ''} {pos[-4:] if len(pos) > 12 else ''}"
              f"  (b,h) {sorted(set((int(x[0]), int(x[1])) for x in idx[:200].tolist()))[:6]}")

run(1, 8, 1, 64, 128, 256, 384)
run(2, 16, 2, 64, 128, 256, 384)
run(1, 8, 1, 64, 128, 256, 256)
run(1, 8, 1, 64, 100, 200, 1000)
run(1, 8, 1, 64, 128, 4096, 4096 + 128)
run(1, 64, 8, 64, 128, 2048, 2048 + 128)
