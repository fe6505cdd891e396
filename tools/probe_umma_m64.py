import sys, os, torch
sys.path.insert(0, "/root/repo/sink-flash-attention-kernel_b200")
from sink_attention import _lib
for dtype in (torch.bfloat16,):
    for n, k in ((64, 128), (128, 128), (144, 128), (128, 64), (16, 128)):
        g = torch.Generator().manual_seed(n + k)
        a = torch.randn(k, 64, generator=g).to("cuda", dtype)        # [K][M=64]
        b = torch.randn(k, n, generator=g).to("cuda", dtype)         # [K][N]
        if n % 64:
            print("skip", n); continue
        c = _lib.probe_umma(a, b, n, k, 3)
        torch.cuda.synchronize()
        ref = a.float().t() @ b.float()
        print(n, k, "max diff rows 0-63:", (c[:64] - ref).abs().max().item(), "ref max", ref.abs().max().item())
