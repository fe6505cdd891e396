"""UMMA operand-form probe (sfa_probe_umma): M = 64 with MN-major A and B from shared memory.
mode 3: one accumulator.  mode 4: two independent M = 64 accumulators in the SAME TMEM columns, the second at
lane offset 16 (it only sums the first half of K) -- the question is whether the hardware accepts the lane offset."""
import sys, os, torch
sys.path.insert(0, "/root/repo/sink-flash-attention-kernel_b200")
from sink_attention import _probe as _lib
for mode in (4,):
    for dtype in (torch.bfloat16,):
        for n, k in ((64, 128), (144, 128), (160, 128), (128, 64)):
            g = torch.Generator().manual_seed(n + k)
            a = torch.randn(k, 64, generator=g).to("cuda", dtype)        # [K][M=64]
            b = torch.randn(k, n, generator=g).to("cuda", dtype)         # [K][N]
            c = _lib.probe_umma(a, b, n, k, mode)
            torch.cuda.synchronize()
            ref = a.float().t() @ b.float()
            msg = f"mode {mode} N={n} K={k}: max diff lanes 0-15: {(c[:64] - ref).abs().max().item():.3e} (ref max {ref.abs().max().item():.1f})"
            if mode == 4:
                ref2 = a[: k // 2].float().t() @ b[: k // 2].float()
                msg += f"; lanes 16-31 vs half-K product: {(c[64:] - ref2).abs().max().item():.3e}"
            print(msg, flush=True)
