"""Development check of the fused backward kernel (bwdf_sm100.cu) against the CUDA-core fp32-math path."""
import sys, time, torch
sys.path.insert(0, "/root/repo/sink-flash-attention-kernel_b200")
import sink_attention as sa
from sink_attention import _lib

def run(B, Hq, Hkv, N, W, dtype=torch.bfloat16, hf_layout=False, seed=0):
    g = torch.Generator().manual_seed(seed + N + W)
    def mk(H):
        if hf_layout:
            return torch.randn(B, N, H, 64, generator=g).to("cuda", dtype).transpose(1, 2)
        return torch.randn(B, H, N, 64, generator=g).to("cuda", dtype)
    q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
    s_aux = (torch.randn(Hq, generator=g) * 0.5 + 1.0).cuda()
    o, lse = sa.sink_flash_attention_with_lse(q, k, v, 0, W, s_aux)
    out_t = _lib.bwd(q, k, v, o, do, lse, 0, W, s_aux)
    torch.cuda.synchronize()
    name = _lib.last_impl()
    _lib.set_impl(_lib.IMPL_SIMT)
    try:
        out_s = _lib.bwd(q, k, v, o, do, lse, 0, W, s_aux)
        torch.cuda.synchronize()
    finally:
        _lib.set_impl(_lib.IMPL_AUTO)
    msg = f"B={B} Hq={Hq} Hkv={Hkv} N={N} W={W} {str(dtype)[6:]} hf={int(hf_layout)} impl={name}:"
    ok = True
    for nm, a, b in zip(("dq", "dk", "dv"), out_t[:3], out_s[:3]):
        d = (a.float() - b.float()).abs()
        tol = 2e-2 + 1e-2 * b.float().abs()
        bad = (d > tol).sum().item()
        msg += f" {nm} max {d.max().item():.3e} bad {bad}"
        if bad:
            ok = False
            idx = (d > tol).nonzero()[0].tolist()
            msg += f" first bad {idx} got {a[tuple(idx)].item():.4f} ref {b[tuple(idx)].item():.4f}"
    print(("OK  " if ok else "FAIL") + " " + msg, flush=True)
    return ok

cases = [
    (1, 16, 2, 1024, 128),
    (1, 8, 1, 256, 128),
    (1, 8, 1, 200, 100),
    (2, 16, 2, 777, 128),
    (1, 8, 1, 4096, 128),
    (1, 8, 1, 512, 17),
    (1, 8, 1, 300, 1),
    (1, 4, 1, 640, 96),       # group 4 -> 32 positions per tile
    (1, 8, 2, 1000, 64),      # group 4
    (1, 64, 8, 8192, 128),    # C1
]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
allok = True
for c in cases:
    allok &= run(*c)
allok &= run(1, 16, 2, 1024, 128, hf_layout=True)
allok &= run(1, 16, 2, 1024, 128, dtype=torch.float16)
print("ALL OK" if allok else "SOME FAILED")
