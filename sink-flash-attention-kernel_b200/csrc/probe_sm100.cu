// tcgen05 / TMA self-test behind sfa_probe_umma: a one-CTA 128 x N x K GEMM in the three operand
// forms the attention kernels rely on.  The parity tests run it first, so a wrong descriptor
// convention shows up as a GEMM mismatch rather than as a confusing attention error.
//   mode 0: C = A[128,K] * B[N,K]^T      both operands K-major (S = Q K^T, dP = dO V^T)
//   mode 1: C = A[128,K] * B[K,N]        B MN-major           (dQ = dS K with smem A)
//   mode 2: as mode 1 with A fed from TMEM                    (O += P V, dV += P^T dO, ...)
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace sfa {
namespace {

template <typename T>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, const T* __restrict__ a_gmem,
                                                    float* __restrict__ c, int N, int K, int mode, int fmt) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_s = smem;                 // K/64 slabs of [128][64]
  unsigned char* b_s = smem + 4 * 16384;     // mode 0: K/64 slabs of [256][64]; mode 1/2: N/64 slabs of [256][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + 4 * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t kColA = 256;  // TMEM columns used for the A operand in mode 2

  if (threadIdx.x == 0) {
    const uint32_t bytes = (mode != 2 ? 128 * K * 2 : 0) + N * K * 2;
    mbar_expect_tx(bars, bytes);
    if (mode != 2)
      for (int s = 0; s < K / 64; ++s) tma_load_4d(a_s + s * 16384, &tmA, bars, s * 64, 0, 0, 0);
    const int nslab_b = (mode == 0) ? K / 64 : N / 64;
    for (int s = 0; s < nslab_b; ++s) tma_load_4d(b_s + s * 32768, &tmB, bars, s * 64, 0, 0, 0);
  }
  if (mode == 2) {  // A row per thread -> packed 16-bit pairs -> TMEM
    const T* row = a_gmem + static_cast<int64_t>(threadIdx.x) * K;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t pk[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) pk[e] = *reinterpret_cast<const uint32_t*>(row + k0 + 2 * e);
      tmem_st8(tl + kColA + (k0 >> 1), pk);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_wait(bars, 0);
    tc_fence_after();
    const uint32_t aa = smem_u32(a_s), ba = smem_u32(b_s);
    if (mode == 0) {
      const uint32_t idesc = make_idesc(fmt, 128, N, 0, 0);
      for (int kk = 0; kk < K / 16; ++kk)
        umma_ss(tmem, make_sdesc(aa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                make_sdesc(ba + (kk >> 2) * 32768 + (kk & 3) * 32, 16, 1024), idesc, kk > 0);
    } else if (mode == 1) {
      const uint32_t idesc = make_idesc(fmt, 128, N, 0, 1);
      for (int kk = 0; kk < K / 16; ++kk)
        umma_ss(tmem, make_sdesc(aa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                make_sdesc(ba + kk * 2048, 32768, 1024), idesc, kk > 0);
    } else {
      const uint32_t idesc = make_idesc(fmt, 128, N, 0, 1);
      for (int kk = 0; kk < K / 16; ++kk)
        umma_ts(tmem, tmem + kColA + kk * 8, make_sdesc(ba + kk * 2048, 32768, 1024), idesc, kk > 0);
    }
    umma_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tl + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) c[static_cast<int64_t>(threadIdx.x) * N + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

cudaError_t probe_umma(const void* a, const void* b, float* c, int N, int K, int mode, int dtype, cudaStream_t st) {
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return cudaErrorInvalidValue;
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256) return cudaErrorInvalidValue;
  if (mode == 0 && K % 64) return cudaErrorInvalidValue;
  if (mode != 0 && (N % 64 || N > 256)) return cudaErrorInvalidValue;
  if (mode == 1 && K % 64) return cudaErrorInvalidValue;
  TileMap ma, mb;
  Strides4 sa{(int64_t)128 * K, (int64_t)128 * K, K};
  if (!make_tile_map(&ma, a, dtype, K, 128, 1, 1, sa, 128, 1)) return cudaErrorInvalidValue;
  if (mode == 0) {
    Strides4 sb{(int64_t)N * K, (int64_t)N * K, K};
    if (!make_tile_map(&mb, b, dtype, K, N, 1, 1, sb, N, 1)) return cudaErrorInvalidValue;
  } else {
    Strides4 sb{(int64_t)N * K, (int64_t)N * K, N};
    if (!make_tile_map(&mb, b, dtype, N, K, 1, 1, sb, K, 1)) return cudaErrorInvalidValue;
  }
  const int smem = 1024 + 4 * 16384 + 4 * 32768 + 64;
  const int fmt = dtype == SFA_DTYPE_BF16 ? 1 : 0;
  cudaError_t e;
  if (dtype == SFA_DTYPE_BF16) {
    e = cudaFuncSetAttribute(probe_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    probe_kernel<__nv_bfloat16><<<1, 128, smem, st>>>(ma.map, mb.map, static_cast<const __nv_bfloat16*>(a), c, N, K, mode, fmt);
  } else {
    e = cudaFuncSetAttribute(probe_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    probe_kernel<__half><<<1, 128, smem, st>>>(ma.map, mb.map, static_cast<const __half*>(a), c, N, K, mode, fmt);
  }
  return cudaGetLastError();
}

}  // namespace sfa
