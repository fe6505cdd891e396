"""Does concurrent shared-memory / TMEM traffic from other warps slow tcgen05.mma down?  (sfa_probe_mma_desc with the
background-load bits: 1 = 12 warps streaming st.shared.v4, 2 = streaming tcgen05.ld, 4 = streaming ld.shared.v4; 8 = random
operand data instead of zeros; 16, 32, 48 = alternate shapes / accumulators from one thread; 256, 512 = one or two MORE threads issue the same UMMA stream into their own accumulators)"""
import ctypes, sys, torch
sys.path.insert(0, "/root/repo/sink-flash-attention-kernel_b200")
from sink_attention import _probe as _lib
lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device="cuda")

def run(name, M, N, a_mn, b_mn, a, b, bg, n=128, ksteps=8, lane=0):
    prm = (ctypes.c_int * 16)(M, N, a_mn, b_mn, *a, *b, n, ksteps, lane, bg)
    rc = lib.sfa_probe_mma_desc(out.data_ptr(), prm, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.sfa_last_error()
    torch.cuda.synchronize()
    t = out.cpu()
    print(f"{name:44s} M={M:3d} N={N:3d} background={bg}: issue {t[0].item() / n:6.1f}  complete {t[1].item() / n:6.1f} cyc/mma", flush=True)

SW_K = (1, 16, 1024, 32)
SW_MN = (1, 16384, 1024, 2048)
for bg in (0, 256, 512):
    run("S form (A, B K-major sw128)", 128, 144, 0, 0, SW_K, SW_K, bg, ksteps=4)
    run("dV^T form (A MN sw128, B un-swizzled)", 64, 160, 1, 1, SW_MN, (0, 128, 2048, 256), bg)
    run("dQ form (A un-swizzled, B MN sw128)", 128, 64, 0, 1, (0, 2048, 128, 4096), (1, 18432, 1024, 2048), bg)
