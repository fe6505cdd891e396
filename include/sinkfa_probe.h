/* libsinkfa_probe -- micro-architecture probes behind tools/probe_*.py (tcgen05 / TMA / TMEM rates and operand
 * layouts measured on the B200).  Diagnostics only: not part of the drop-in boundary (include/sinkfa.h) and not
 * linked into libsinkfa.so.  Return value: 0 ok, > 0 a cudaError_t; sfa_probe_last_error() describes it. */
#ifndef SINKFA_PROBE_H_
#define SINKFA_PROBE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* sfa_probe_last_error(void);

/* tcgen05/TMA self-test: C[M=128,N] = A[128,K] * B^T (+ variants).  Returns 0 and fills c (fp32, device).
 * mode 0: A,B K-major in smem; mode 1: B given as [K,N] (MN-major); mode 2: A through TMEM (TS form). */
int sfa_probe_umma(const void* a, const void* b, float* c, int N, int K, int mode, int dtype, void* stream);

/* UMMA issue-rate probe: out2 = device int64[2] <- {cycles to issue, cycles until complete} for reps*ksteps UMMAs */
int sfa_probe_mma_rate(void* out2, int N, int ksteps, int reps, int uniform, void* stream);
/* UMMA operand-layout timing probe: prm16 = {M, N, a_mn_major, b_mn_major, a_swizzle128, a_lbo, a_sbo, a_kstep_bytes,
   b_swizzle128, b_lbo, b_sbo, b_kstep_bytes, n_mmas, ksteps, d_lane_offset, 0}; out2 as sfa_probe_mma_rate */
int sfa_probe_mma_desc(void* out2, const int* prm16, void* stream);
/* math-pipe probe: out1 = device int64[1] <- cycles for `iters` 16-element softmax steps of one warp (see probe_sm100.cu) */
int sfa_probe_math_rate(void* out1, void* sink, int mode, int iters, int threads, void* stream);
/* TMEM read-throughput probe: out1 <- cycles for `iters` tcgen05.ld round trips per warp (mode 0: x16, 1: x32, 2: 2 x x32) */
int sfa_probe_tmem_rate(void* out1, void* sink, int mode, int iters, int threads, void* stream);
/* load-throughput probe (performance work): streams a bf16 [H,N,64] tensor through shared memory with TMA
 * boxes of box_n positions x box_h heads (mode 0) or per-thread cp.async (mode 1), `stages` boxes in flight. */
int sfa_probe_tma_bw(const void* src, int H, int N, int box_n, int box_h, int stages, int grid, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SINKFA_PROBE_H_ */
