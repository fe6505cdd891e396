#!/usr/bin/env python
"""Short ncu target: a few fwd+bwd steps of the C1 layer and a few C3 decode steps through the public API.

  python tools/prof_target.py [--cfg c1|c2s] [--iters 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402

CFG = {
    "c1": dict(B=1, N=8192, Hq=64, Hkv=8, D=64, S=0, W=128),            # BASELINE configs[1]
    "c2s": dict(B=1, N=8192, Hq=32, Hkv=8, D=128, S=4, W=4096),         # configs[2] shortened (B=1, N=8192)
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="c1")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--decode", type=int, default=1)
    a = ap.parse_args()
    c = CFG[a.cfg]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(42)
    dt = torch.bfloat16
    q = torch.randn(c["B"], c["Hq"], c["N"], c["D"], device=dev, generator=g).to(dt).requires_grad_(True)
    k = torch.randn(c["B"], c["Hkv"], c["N"], c["D"], device=dev, generator=g).to(dt).requires_grad_(True)
    v = torch.randn(c["B"], c["Hkv"], c["N"], c["D"], device=dev, generator=g).to(dt).requires_grad_(True)
    do = torch.randn(c["B"], c["Hq"], c["N"], c["D"], device=dev, generator=g).to(dt)
    s_aux = (torch.randn(c["Hq"], device=dev, generator=g) * 0.5).requires_grad_(True)
    for _ in range(a.iters):
        for t in (q, k, v, s_aux):
            t.grad = None
        o = sa.sink_flash_attention(q, k, v, c["S"], c["W"], s_aux)
        o.backward(do)
    torch.cuda.synchronize()
    if a.decode:
        qq = torch.randn(64, 64, 1, 64, device=dev, generator=g).to(dt)
        kk = torch.randn(64, 8, 4100, 64, device=dev, generator=g).to(dt)
        vv = torch.randn(64, 8, 4100, 64, device=dev, generator=g).to(dt)
        for _ in range(a.iters):
            sa.sink_decode_attention(qq, kk, vv, s_aux.detach())
        torch.cuda.synchronize()
    print("prof_target ok")


if __name__ == "__main__":
    main()
