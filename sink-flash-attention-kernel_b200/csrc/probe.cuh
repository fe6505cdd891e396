// Launchers of the micro-architecture probes (probe_sm100.cu).  They live in their own library,
// libsinkfa_probe.so (extern "C" surface: probe_api.cu, include/sinkfa_probe.h): diagnostics for performance work,
// not part of the product ABI.
#pragma once
#include "common.cuh"

namespace sfa {
cudaError_t probe_tma_bw(const void* src, int H, int N, int box_n, int box_h, int stages, int grid, int mode,
                         cudaStream_t st);
cudaError_t probe_mma_rate(long long* out, int N, int ksteps, int reps, int uniform, cudaStream_t st);
cudaError_t probe_mma_desc(long long* out, const int* prm16, cudaStream_t st);
cudaError_t probe_math_rate(long long* out, float* sink, int mode, int iters, int threads, cudaStream_t st);
cudaError_t probe_tmem_rate(long long* out, float* sink, int mode, int iters, int threads, cudaStream_t st);
cudaError_t probe_umma(const void* a, const void* b, float* c, int N, int K, int mode, int dtype, cudaStream_t st);

}  // namespace sfa
