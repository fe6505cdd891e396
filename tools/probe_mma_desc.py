"""Cycles per tcgen05.mma by operand layout (sfa_probe_mma_desc): which shared-memory layouts feed the tensor
core at full rate.  prm = M, N, a_mn, b_mn, a_swz, a_lbo, a_sbo, a_kstep, b_swz, b_lbo, b_sbo, b_kstep, n, ksteps, lane."""
import ctypes, sys, torch
sys.path.insert(0, "/root/repo/sink-flash-attention-kernel_b200")
from sink_attention import _probe as _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")

def run(name, M, N, a_mn, b_mn, a, b, n=128, ksteps=8, lane=0):
    prm = (ctypes.c_int * 16)(M, N, a_mn, b_mn, *a, *b, n, ksteps, lane, 0)
    rc = lib.sfa_probe_mma_desc(out.data_ptr(), prm, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.sfa_last_error()
    torch.cuda.synchronize()
    t = out.cpu()
    print(f"{name:78s} M={M:3d} N={N:3d}: issue {t[0].item() / n:6.1f}  complete {t[1].item() / n:6.1f} cyc/mma", flush=True)

SW_K = (1, 16, 1024, 32)          # K-major SWIZZLE_128B tile [rows][64]: k-step = 32 B
SW_MN = (1, 16384, 1024, 2048)    # MN-major SWIZZLE_128B [K rows][64]: k-step = 16 rows
for N in (64, 144):
    run("S / dP form: A K-major sw128, B K-major sw128", 128, N, 0, 0, SW_K, SW_K, ksteps=4)
run("dQ old: A K-major sw128 (stand-in), B MN-major sw128", 128, 64, 0, 1, SW_K, (1, 18432, 1024, 2048), ksteps=4)
# dQ with the un-swizzled dS image, layout L1 = [c/8][r/8][r%8][8]: SBO(row groups)=128, LBO(k halves)=2048, k-step 4096
run("dQ L1: A un-swizzled [c/8][r/8] (SBO 128, LBO 2048), B MN sw128", 128, 64, 0, 1, (0, 2048, 128, 4096), (1, 18432, 1024, 2048), ksteps=8)
# layout L2 = [r/8][c/8][r%8][8] with 20 col groups: SBO = 2560, LBO = 128, k-step 256
run("dQ L2: A un-swizzled [r/8][c/8] (SBO 2560, LBO 128), B MN sw128", 128, 64, 0, 1, (0, 128, 2560, 256), (1, 18432, 1024, 2048), ksteps=8)
for N in (16, 64, 80, 144, 160):
    run("dV^T sw: A MN sw128, B MN sw128 (LBO 16384)", 64, N, 1, 1, SW_MN, SW_MN)
    run("dV^T L1: A MN sw128, B un-swizzled [c/8][r/8] (LBO 128, SBO 2048)", 64, N, 1, 1, SW_MN, (0, 128, 2048, 256))
    run("dV^T L2: A MN sw128, B un-swizzled [r/8][c/8] (LBO 2560, SBO 128)", 64, N, 1, 1, SW_MN, (0, 2560, 128, 5120))
    run("M=128 stand-in: A MN sw128 (2 slabs), B L1", 128, N, 1, 1, SW_MN, (0, 128, 2048, 256))
run("dV^T L1 at lane offset 16", 64, 144, 1, 1, SW_MN, (0, 128, 2048, 256), lane=16)
