"""UMMA probe for operands the math warps write themselves in the un-swizzled core-matrix layout
(sfa_probe_umma modes 5 and 6): checks the LBO / SBO convention of make_sdesc_ns."""
import sys, torch
sys.path.insert(0, "/root/repo/sink-flash-attention-kernel_b200")
from sink_attention import _probe as _lib
dtype = torch.bfloat16
for n, k in ((64, 128), (144, 128), (160, 128), (16, 128), (48, 64)):
    g = torch.Generator().manual_seed(n + k)
    a = torch.randn(k, 64, generator=g).to("cuda", dtype)
    b = torch.randn(k, n, generator=g).to("cuda", dtype)
    c = _lib.probe_umma(a, b, n, k, 5)
    torch.cuda.synchronize()
    ref = a.float().t() @ b.float()
    print(f"mode 5 (M=64, B un-swizzled MN-major) N={n} K={k}: max diff {(c[:64] - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.1f})", flush=True)
for n, k in ((64, 128), (64, 144), (128, 64), (64, 16)):
    g = torch.Generator().manual_seed(n + k)
    a = torch.randn(128, k, generator=g).to("cuda", dtype)
    b = torch.randn(k, n, generator=g).to("cuda", dtype)
    c = _lib.probe_umma(a, b, n, k, 6)
    torch.cuda.synchronize()
    ref = a.float() @ b.float()
    print(f"mode 6 (M=128, A un-swizzled K-major) N={n} K={k}: max diff {(c - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.1f})", flush=True)
