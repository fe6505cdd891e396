// extern "C" surface of libsinkfa_probe.so (include/sinkfa_probe.h): tcgen05 / TMA / TMEM micro-probes used by
// tools/probe_*.py and by the UMMA descriptor self-test of the GPU suite.  Kept out of libsinkfa.so: the product
// library exports the operator entry points only.
#include <stdarg.h>
#include <stdio.h>

#include "probe.cuh"
#include "../../include/sinkfa_probe.h"

namespace sfa {
static thread_local char g_perr[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_perr, sizeof(g_perr), fmt, ap);
  va_end(ap);
}
void set_impl_name(const char*) {}
int debug_knob(int) { return 0; }
int device_sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}
static int ret(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
}  // namespace sfa

using namespace sfa;

extern "C" {

const char* sfa_probe_last_error(void) { return g_perr; }

int sfa_probe_tma_bw(const void* src, int H, int N, int box_n, int box_h, int stages, int grid, int mode, void* stream) {
  return ret(probe_tma_bw(src, H, N, box_n, box_h, stages, grid, mode, static_cast<cudaStream_t>(stream)), "sfa_probe_tma_bw");
}
int sfa_probe_mma_rate(void* out2, int N, int ksteps, int reps, int uniform, void* stream) {
  return ret(probe_mma_rate(static_cast<long long*>(out2), N, ksteps, reps, uniform, static_cast<cudaStream_t>(stream)),
             "sfa_probe_mma_rate");
}
int sfa_probe_mma_desc(void* out2, const int* prm16, void* stream) {
  return ret(probe_mma_desc(static_cast<long long*>(out2), prm16, static_cast<cudaStream_t>(stream)), "sfa_probe_mma_desc");
}
int sfa_probe_math_rate(void* out1, void* sink, int mode, int iters, int threads, void* stream) {
  return ret(probe_math_rate(static_cast<long long*>(out1), static_cast<float*>(sink), mode, iters, threads,
                             static_cast<cudaStream_t>(stream)), "sfa_probe_math_rate");
}
int sfa_probe_tmem_rate(void* out1, void* sink, int mode, int iters, int threads, void* stream) {
  return ret(probe_tmem_rate(static_cast<long long*>(out1), static_cast<float*>(sink), mode, iters, threads,
                             static_cast<cudaStream_t>(stream)), "sfa_probe_tmem_rate");
}
int sfa_probe_umma(const void* a, const void* b, float* c, int N, int K, int mode, int dtype, void* stream) {
  return ret(probe_umma(a, b, c, N, K, mode, dtype, static_cast<cudaStream_t>(stream)), "sfa_probe_umma");
}

}  // extern "C"
