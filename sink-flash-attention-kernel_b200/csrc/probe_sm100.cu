// tcgen05 / TMA self-test behind sfa_probe_umma: a one-CTA 128 x N x K GEMM in the three operand
// forms the attention kernels rely on.  The parity tests run it first, so a wrong descriptor
// convention shows up as a GEMM mismatch rather than as a confusing attention error.
//   mode 0: C = A[128,K] * B[N,K]^T      both operands K-major (S = Q K^T, dP = dO V^T)
//   mode 1: C = A[128,K] * B[K,N]        B MN-major           (dQ = dS K with smem A)
//   mode 2: as mode 1 with A fed from TMEM                    (O += P V, dV += P^T dO, ...)
//   mode 3: M = 64, A [K,64] and B [K,N] both MN-major (SWIZZLE_128B, TMA-written)   (dV^T = dO^T P)
//   mode 4: mode 3 plus a second M = 64 accumulator in the same columns at lane offset 16
//   mode 5: mode 3 with B written by the threads in the un-swizzled core-matrix layout
//   mode 6: M = 128, A [128,K] K-major written by the threads un-swizzled, B [K,N] MN-major SWIZZLE_128B
#include "common.cuh"
#include "probe.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace sfa {
namespace {

template <typename T>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, const T* __restrict__ a_gmem,
                                                    const T* __restrict__ b_gmem, float* __restrict__ c, int N, int K, int mode, int fmt) {
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
    const int nslab_mn = (N + 63) / 64;
    const bool a_tma = (mode != 2 && mode != 6), b_tma = (mode != 5);
    const uint32_t bytes = (!a_tma ? 0 : (mode >= 3 ? 64 * K * 2 : 128 * K * 2)) +
                           (!b_tma ? 0 : (mode >= 3 ? nslab_mn * 64 * K * 2 : N * K * 2));
    mbar_expect_tx(bars, bytes);
    if (mode == 6) {
    } else if (mode >= 3) tma_load_4d(a_s, &tmA, bars, 0, 0, 0, 0);
    else if (mode != 2)
      for (int s = 0; s < K / 64; ++s) tma_load_4d(a_s + s * 16384, &tmA, bars, s * 64, 0, 0, 0);
    const int nslab_b = (mode == 0) ? K / 64 : (N + 63) / 64;
    if (b_tma)
      for (int s = 0; s < nslab_b; ++s) tma_load_4d(b_s + s * 32768, &tmB, bars, s * 64, 0, 0, 0);
  }
  if (mode == 5 && static_cast<int>(threadIdx.x) < K) {   // B row k -> [n / 8][k / 8][k % 8][n % 8]
    const int k = threadIdx.x;
    for (int n0 = 0; n0 < N; n0 += 8)
      *reinterpret_cast<uint4*>(b_s + (n0 >> 3) * 2048 + (k >> 3) * 128 + (k & 7) * 16) =
          *reinterpret_cast<const uint4*>(b_gmem + static_cast<int64_t>(k) * N + n0);
    fence_proxy_async_smem();
  }
  if (mode == 6) {                                          // A row r -> [k / 8][r / 8][r % 8][k % 8]
    const int r = threadIdx.x;
    for (int k0 = 0; k0 < K; k0 += 8)
      *reinterpret_cast<uint4*>(a_s + (k0 >> 3) * 2048 + (r >> 3) * 128 + (r & 7) * 16) =
          *reinterpret_cast<const uint4*>(a_gmem + static_cast<int64_t>(r) * K + k0);
    fence_proxy_async_smem();
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
    } else if (mode == 2) {
      const uint32_t idesc = make_idesc(fmt, 128, N, 0, 1);
      for (int kk = 0; kk < K / 16; ++kk)
        umma_ts(tmem, tmem + kColA + kk * 8, make_sdesc(ba + kk * 2048, 32768, 1024), idesc, kk > 0);
    } else {
      // mode 3: M = 64, A MN-major (tile stored [K rows][64 M-elements]), B MN-major ([K rows][N], 64-wide slabs)
      const uint32_t idesc = make_idesc(fmt, 64, N, 1, 1);
      if (mode < 5)
      for (int kk = 0; kk < K / 16; ++kk)
        umma_ss(tmem, make_sdesc(aa + kk * 2048, 16384, 1024), make_sdesc(ba + kk * 2048, 32768, 1024), idesc, kk > 0);
      if (mode == 5) {
        for (int kk = 0; kk < K / 16; ++kk)
          umma_ss(tmem, make_sdesc(aa + kk * 2048, 16384, 1024), make_sdesc_ns(ba + kk * 256, 128, 2048), idesc, kk > 0);
      } else if (mode == 6) {
        const uint32_t idesc6 = make_idesc(fmt, 128, N, 0, 1);
        for (int kk = 0; kk < K / 16; ++kk)
          umma_ss(tmem, make_sdesc_ns(aa + kk * 4096, 2048, 128), make_sdesc(ba + kk * 2048, 32768, 1024), idesc6, kk > 0);
      } else
      // mode 4: a second, independent M = 64 accumulator in the SAME columns at lane offset 16 (first half of K only)
      if (mode == 4)
        for (int kk = 0; kk < K / 32; ++kk)
          umma_ss(tmem + (16u << 16), make_sdesc(aa + kk * 2048, 16384, 1024), make_sdesc(ba + kk * 2048, 32768, 1024),
                  idesc, kk > 0);
    }
    umma_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tl + c0, v);
    tmem_ld_wait();
    if (mode >= 3 && mode != 6) {      // M = 64: row r lives in lane (r % 16) + 32 * (r / 16); rows 64..127 of c receive the other lanes
      const int row = (lane < 16) ? warp * 16 + lane : 64 + warp * 16 + (lane - 16);
#pragma unroll
      for (int e = 0; e < 16; ++e) c[static_cast<int64_t>(row) * N + c0 + e] = __uint_as_float(v[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) c[static_cast<int64_t>(threadIdx.x) * N + c0 + e] = __uint_as_float(v[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// UMMA issue-rate probe: one thread issues `reps` x (K/16) SS-form UMMAs (M = 128, N, K from zero-filled
// shared memory) and records clock64 after the issue loop and after the commit has landed.
__global__ void __launch_bounds__(128) mma_rate_kernel(long long* out, int N, int ksteps, int reps, int fmt, int uniform) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t idesc = make_idesc(fmt, 128, N, 0, 0);
  const uint64_t ad = make_sdesc(smem_u32(smem), 16, 1024), bd = make_sdesc(smem_u32(smem + 16384), 16, 1024);
  if (uniform) {
    if (warp == 1) {                       // whole warp runs the loop; one elected lane issues
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int kk = 0; kk < ksteps; ++kk)
          if (elect_one()) umma_ss(tmem + (r & 1) * 256, ad + (kk & 3) * 2, bd + (kk & 3) * 2, idesc, kk > 0);
      const long long t1 = clock64();
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, 0);
      const long long t2 = clock64();
      if (threadIdx.x == 32) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  } else if (threadIdx.x == 32) {
    const long long t0 = clock64();
    if (uniform == 0) {
      for (int r = 0; r < reps; ++r)
        for (int kk = 0; kk < ksteps; ++kk) umma_ss(tmem + (r & 1) * 256, ad + (kk & 3) * 2, bd + (kk & 3) * 2, idesc, kk > 0);
    } else if (uniform == 2) {      // two interleaved accumulation chains
      for (int r = 0; r < reps; ++r)
        for (int kk = 0; kk < ksteps; ++kk) umma_ss(tmem + (kk & 1) * 256, ad + (kk & 3) * 2, bd + (kk & 3) * 2, idesc, kk > 1);
    } else if (uniform == 3) {      // four interleaved chains
      for (int r = 0; r < reps; ++r)
        for (int kk = 0; kk < ksteps; ++kk) umma_ss(tmem + (kk & 3) * 128, ad + (kk & 3) * 2, bd + (kk & 3) * 2, idesc, kk > 3);
    } else {                        // fully unrolled issue, no loop arithmetic
      for (int r = 0; r < reps; ++r) {
        umma_ss(tmem, ad, bd, idesc, 0);
        umma_ss(tmem, ad + 2, bd + 2, idesc, 1);
        umma_ss(tmem, ad + 4, bd + 4, idesc, 1);
        umma_ss(tmem, ad + 6, bd + 6, idesc, 1);
      }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// UMMA operand-layout timing probe: one thread issues n UMMAs whose descriptors are built from raw numbers
// (prm[]: M, N, a_mn, b_mn, a_swz, a_lbo, a_sbo, a_kstep, b_swz, b_lbo, b_sbo, b_kstep, n, ksteps, d_lane), shared
// memory zero-filled: cycles per instruction for each layout the kernels consider.
struct DescPrm { int v[16]; };
__global__ void __launch_bounds__(512) mma_desc_kernel(long long* out, DescPrm p, int fmt) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBytes = 160 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  // p.v[15] bit 3: pseudo-random bf16 operands (|x| < 2) instead of zeros; bit 4: alternate with S-form UMMAs
  for (int i = threadIdx.x; i < kBytes / 4; i += blockDim.x) {
    uint32_t h = static_cast<uint32_t>(i) * 2654435761u;
    h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (p.v[15] & 8) ? ((h & 0x807f807fu) | 0x3f003f00u) : 0u;
  }
  const int warp = threadIdx.x >> 5;
  volatile uint32_t* done_flag = slot + 1;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 2, 1);
    mbar_init(bar + 3, 1);
    fence_barrier_init();
    *done_flag = 0u;
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int M = p.v[0], N = p.v[1], n = p.v[12], ksteps = p.v[13];
  // background load (p.v[15]): bit 0 = warps 4-15 stream st.shared.v4 into a scratch region, bit 1 = they stream
  // tcgen05.ld of 16 columns, bit 2 = they stream ld.shared.v4 -- while thread 32 issues the UMMAs
  if (warp >= 4) {
    const int bg = p.v[15] & 7;
    const uint32_t scratch = smem_u32(smem + 128 * 1024) + (threadIdx.x - 128) * 16;
    const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256;
    uint32_t acc = 0;
    while (*done_flag == 0u && bg != 0) {
      if (bg & 1)
        for (int i = 0; i < 4; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(scratch + i * 6144), "r"(acc) : "memory");
      if (bg & 2) {
        uint32_t x[16];
        tmem_ld16(tl, x);
        tmem_ld_wait();
        acc += x[3];
      }
      if (bg & 4)
        for (int i = 0; i < 4; ++i) {
          uint32_t a0, a1, a2, a3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(scratch + i * 6144) : "memory");
          acc += a0 + a3;
        }
    }
    if (acc == 0x12345678u) out[2] = acc;
  }
  const uint32_t idesc = make_idesc(fmt, M, N, p.v[2], p.v[3]);
  const uint32_t aa = smem_u32(smem), ba = smem_u32(smem + 64 * 1024);
  const int extra = (p.v[15] >> 8) & 3;     // extra issuing threads (lane 0 of warps 2, 3): same UMMAs, own accumulators
  if ((threadIdx.x == 64 && extra >= 1) || (threadIdx.x == 96 && extra >= 2)) {
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k2 = kk % ksteps;
      ad[kk] = p.v[4] ? make_sdesc(aa + k2 * p.v[7], p.v[5], p.v[6]) : make_sdesc_ns(aa + k2 * p.v[7], p.v[5], p.v[6]);
      bd[kk] = p.v[8] ? make_sdesc(ba + k2 * p.v[11], p.v[9], p.v[10]) : make_sdesc_ns(ba + k2 * p.v[11], p.v[9], p.v[10]);
    }
    const uint32_t d = tmem + (warp - 1) * 160;
    for (int i = 0; i < n; i += 8) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) umma_ss(d, ad[kk], bd[kk], idesc, kk > 0);
    }
    umma_commit(bar + 2 + (warp - 2));
    mbar_wait(bar + 2 + (warp - 2), 0);
  }
  if (threadIdx.x == 32) {
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k2 = kk % ksteps;
      ad[kk] = p.v[4] ? make_sdesc(aa + k2 * p.v[7], p.v[5], p.v[6]) : make_sdesc_ns(aa + k2 * p.v[7], p.v[5], p.v[6]);
      bd[kk] = p.v[8] ? make_sdesc(ba + k2 * p.v[11], p.v[9], p.v[10]) : make_sdesc_ns(ba + k2 * p.v[11], p.v[9], p.v[10]);
    }
    const uint32_t d = tmem + (static_cast<uint32_t>(p.v[14]) << 16);
    const uint32_t idesc_s = make_idesc(fmt, 128, 144, 0, 0);
    const uint64_t sa = make_sdesc(aa + 32768, 16, 1024), sb = make_sdesc(ba + 49152, 16, 1024);
    const int alt = p.v[15] >> 4;   // 0: one chain; 1: alternate with S-form UMMAs; 2: same shape, alternate accumulators;
                                    // 3: same shape, two accumulators in blocks of 8
    long long t0;
    if (alt == 0) {
      t0 = clock64();
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss(d, ad[kk], bd[kk], idesc, kk > 0);
      }
    } else if (alt == 1) {
      t0 = clock64();
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; kk += 2) {
          umma_ss(d, ad[kk], bd[kk], idesc, kk > 0);
          umma_ss(tmem + 288, sa + (kk >> 1) * 2, sb + (kk >> 1) * 2, idesc_s, kk > 0);
        }
      }
    } else if (alt == 2) {
      t0 = clock64();
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; kk += 2) {
          umma_ss(d, ad[kk], bd[kk], idesc, kk > 0);
          umma_ss(d + 256, ad[kk + 1], bd[kk + 1], idesc, kk > 0);
        }
      }
    } else {
      t0 = clock64();
      for (int i = 0; i < n; i += 16) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss(d, ad[kk], bd[kk], idesc, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss(d + 256, ad[kk], bd[kk], idesc, kk > 0);
      }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    *done_flag = 1u;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// TMEM read-throughput probe: every warp loops over tcgen05.ld (32x32b, x16 / x32 / 2 x x32 per wait) on its lane quarter.
__global__ void tmem_rate_kernel(long long* out, float* sink, int mode, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tl = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (it * 64) & 255;
    if (mode == 0) {
      uint32_t v[16];
      tmem_ld16(tl + col, v);
      tmem_ld_wait();
      acc ^= v[0] ^ v[15];
    } else if (mode == 1) {
      uint32_t v[32];
      tmem_ld32(tl + col, v);
      tmem_ld_wait();
      acc ^= v[0] ^ v[31];
    } else {
      uint32_t v[32], w[32];
      tmem_ld32(tl + col, v);
      tmem_ld32(tl + col + 32, w);
      tmem_ld_wait();
      acc ^= v[0] ^ w[31];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345u) sink[threadIdx.x] = 1.f;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

// Math-pipe probe: cycles per 16-element step of the softmax inner loop on registers only.
//   mode 0: 16 FFMA + 16 MUFU.EX2          mode 1: + 8 cvt.rn.bf16x2 (F2FP)      mode 2: + integer bf16 packing
//   mode 3: 16 FFMA + 8 F2FP (no MUFU)     mode 4: 16 FFMA only
__global__ void math_rate_kernel(long long* out, float* sink, int mode, int iters, float a, float b) {
  // modes >= 10: as mode-10, plus every warp beyond the first 8 is a poller spinning on an mbarrier that completes
  // only when the math warps are done (how much do spinning roles cost the math warps of their scheduler?)
  __shared__ uint64_t spin_bar;
  const bool pollers = mode >= 10;
  if (pollers) mode -= 10;
  if (threadIdx.x == 0) {
    mbar_init(&spin_bar, 256);
    fence_barrier_init();
  }
  __syncthreads();
  if (pollers && threadIdx.x >= 256) {
    if ((threadIdx.x & 31) == 0) mbar_wait(&spin_bar, 0);
    return;
  }
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = static_cast<float>(threadIdx.x + e) * 1e-3f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[8];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      x[e] = fmaf(x[e], a, b);
      if (mode <= 2) x[e] = fast_exp2(x[e]);
    }
    if (mode == 1 || mode == 3) {
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        __nv_bfloat162 v = __floats2bfloat162_rn(x[e], x[e + 1]);
        pk[e >> 1] = *reinterpret_cast<uint32_t*>(&v);
      }
    } else if (mode == 2) {
#pragma unroll
      for (int e = 0; e < 16; e += 2)
        pk[e >> 1] = __byte_perm(__float_as_uint(x[e]) + 0x8000u, __float_as_uint(x[e + 1]) + 0x8000u, 0x7632);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) pk[e] = 0;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc ^= pk[e];
  }
  const long long t1 = clock64();
  if (pollers) mbar_arrive(&spin_bar);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float ssum = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) ssum += x[e];
  if (ssum == 123.456f || acc == 0x12345u) sink[threadIdx.x] = ssum;
}

// Instruction-footprint probe: the mode-1 step (16 FFMA + 16 EX2 + 8 F2FP) with the loop body replicated U times,
// i.e. the same instruction mix streamed from a code footprint of U x ~0.8 KB (L0 I-cache ~6 KB per scheduler).
template <int U>
__global__ void math_unrolled_kernel(long long* out, float* sink, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = static_cast<float>(threadIdx.x + e) * 1e-3f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t pk[8];
#pragma unroll
      for (int e = 0; e < 16; ++e) x[e] = fast_exp2(fmaf(x[e], a, b));
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        __nv_bfloat162 v = __floats2bfloat162_rn(x[e], x[e + 1]);
        pk[e >> 1] = *reinterpret_cast<uint32_t*>(&v);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc ^= pk[e] + u;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float ssum = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) ssum += x[e];
  if (ssum == 123.456f || acc == 0x12345u) sink[threadIdx.x] = ssum;
}

// TMA load-throughput probe (performance work, tools/probe_tma.py): every CTA streams boxes of
// box_n positions x box_h heads x 64 channels (16-bit) through a ring of `stages` shared-memory
// buffers; nothing consumes the data.  mode 1 replaces TMA by per-thread cp.async (16 B each).
__global__ void __launch_bounds__(160) tma_bw_kernel(const __grid_constant__ CUtensorMap tm, const unsigned char* src,
                                                     int swap, int nblk, int nboxes, int box_n, int box_h,
                                                     int stages, int mode, long long head_stride_bytes) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int box_bytes = box_n * box_h * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * box_bytes);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full + s, mode == 0 ? 1 : 128);
      mbar_init(empty + s, 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (mode == 0) {
    if (warp == 0 && lane == 0) {
      int k = 0;
      for (int bx = blockIdx.x; bx < nboxes; bx += gridDim.x, ++k) {
        const int st = k % stages;
        mbar_wait(empty + st, ((k / stages) & 1) ^ 1);
        mbar_expect_tx(full + st, box_bytes);
        const int pb = bx % nblk, hg = bx / nblk;
        if (swap) tma_load_4d(smem + st * box_bytes, &tm, full + st, 0, hg * box_h, pb * box_n, 0);
        else tma_load_4d(smem + st * box_bytes, &tm, full + st, 0, pb * box_n, hg * box_h, 0);
      }
    } else if (warp == 1 && lane == 0) {
      int k = 0;
      for (int bx = blockIdx.x; bx < nboxes; bx += gridDim.x, ++k) {
        const int st = k % stages;
        mbar_wait(full + st, (k / stages) & 1);
        mbar_arrive(empty + st);
      }
    }
  } else {
    // warps 0-3: cp.async producers (16 B per thread per request, rows of 128 B, swizzled like the TMA box);
    // warp 4 lane 0: consumer
    if (warp < 4) {
      int k = 0;
      const int t = threadIdx.x;
      for (int bx = blockIdx.x; bx < nboxes; bx += gridDim.x, ++k) {
        const int st = k % stages;
        if (lane == 0) mbar_wait(empty + st, ((k / stages) & 1) ^ 1);
        __syncwarp();
        const int pb = bx % nblk, hg = bx / nblk;
        for (int idx = t; idx < box_n * box_h * 8; idx += 128) {
          const int row = idx >> 3, ch = idx & 7;
          const int hh = row / box_n, nn = row % box_n;
          const unsigned char* g = src + (long long)(hg * box_h + hh) * head_stride_bytes + (long long)(pb * box_n + nn) * 128 + ch * 16;
          const uint32_t d = smem_u32(smem + st * box_bytes + sw128_off(row, ch));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full + st)) : "memory");
      }
    } else if (lane == 0) {
      int k = 0;
      for (int bx = blockIdx.x; bx < nboxes; bx += gridDim.x, ++k) {
        const int st = k % stages;
        mbar_wait(full + st, (k / stages) & 1);
        mbar_arrive(empty + st);
      }
    }
  }
}

}  // namespace

cudaError_t probe_tma_bw(const void* src, int H, int N, int box_n, int box_h, int stages, int grid, int mode,
                         cudaStream_t st) {
  TileMap m;
  Strides4 sa{(int64_t)H * N * 64, (int64_t)N * 64, 64};
  if (!make_tile_map(&m, src, SFA_DTYPE_BF16, 64, N, H, 1, sa, box_n, box_h)) return cudaErrorInvalidValue;
  const int box_bytes = box_n * box_h * 128;
  const int smem = 1024 + stages * box_bytes + 2 * stages * 8 + 64;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const int nblk = N / box_n, nboxes = nblk * (H / box_h);
  tma_bw_kernel<<<grid, 160, smem, st>>>(m.map, static_cast<const unsigned char*>(src), m.swap_nh, nblk, nboxes, box_n,
                                         box_h, stages, mode, (long long)N * 128);
  return cudaGetLastError();
}

cudaError_t probe_tmem_rate(long long* out, float* sink, int mode, int iters, int threads, cudaStream_t st) {
  tmem_rate_kernel<<<148, threads, 0, st>>>(out, sink, mode, iters);
  return cudaGetLastError();
}

cudaError_t probe_math_rate(long long* out, float* sink, int mode, int iters, int threads, cudaStream_t st) {
  if (mode >= 100) {   // mode 100 + U: unrolled body
    switch (mode - 100) {
      case 1: math_unrolled_kernel<1><<<148, threads, 0, st>>>(out, sink, iters, 0.999f, -0.001f); break;
      case 8: math_unrolled_kernel<8><<<148, threads, 0, st>>>(out, sink, iters, 0.999f, -0.001f); break;
      case 16: math_unrolled_kernel<16><<<148, threads, 0, st>>>(out, sink, iters, 0.999f, -0.001f); break;
      case 64: math_unrolled_kernel<64><<<148, threads, 0, st>>>(out, sink, iters, 0.999f, -0.001f); break;
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  math_rate_kernel<<<148, threads, 0, st>>>(out, sink, mode, iters, 0.999f, -0.001f);
  return cudaGetLastError();
}

cudaError_t probe_mma_rate(long long* out, int N, int ksteps, int reps, int uniform, cudaStream_t st) {
  const int smem = 1024 + 16384 + 32768 + 64;
  cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  mma_rate_kernel<<<1, 128, smem, st>>>(out, N, ksteps, reps, 1, uniform);
  return cudaGetLastError();
}

cudaError_t probe_mma_desc(long long* out, const int* prm16, cudaStream_t st) {
  const int smem = 1024 + 160 * 1024 + 64;
  cudaError_t e = cudaFuncSetAttribute(mma_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  DescPrm p;
  for (int i = 0; i < 16; ++i) p.v[i] = prm16[i];
  mma_desc_kernel<<<1, 512, smem, st>>>(out, p, 1);
  return cudaGetLastError();
}

cudaError_t probe_umma(const void* a, const void* b, float* c, int N, int K, int mode, int dtype, cudaStream_t st) {
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return cudaErrorInvalidValue;
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256) return cudaErrorInvalidValue;
  if (mode == 0 && K % 64) return cudaErrorInvalidValue;
  if (mode != 0 && mode < 3 && (N % 64 || N > 256)) return cudaErrorInvalidValue;
  if (mode == 1 && K % 64) return cudaErrorInvalidValue;
  if (mode >= 3 && K > 128) return cudaErrorInvalidValue;
  TileMap ma, mb;
  Strides4 sa{(int64_t)128 * K, (int64_t)128 * K, K};
  if (mode == 6) {
    Strides4 s3{(int64_t)K * 64, (int64_t)K * 64, 64};   // unused placeholder map
    if (!make_tile_map(&ma, b, dtype, 64, K, 1, 1, s3, K, 1)) return cudaErrorInvalidValue;
  } else if (mode >= 3) {        // A given as [K][64]
    Strides4 s3{(int64_t)K * 64, (int64_t)K * 64, 64};
    if (!make_tile_map(&ma, a, dtype, 64, K, 1, 1, s3, K, 1)) return cudaErrorInvalidValue;
  } else if (!make_tile_map(&ma, a, dtype, K, 128, 1, 1, sa, 128, 1)) return cudaErrorInvalidValue;
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
    probe_kernel<__nv_bfloat16><<<1, 128, smem, st>>>(ma.map, mb.map, static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), c, N, K, mode, fmt);
  } else {
    e = cudaFuncSetAttribute(probe_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    probe_kernel<__half><<<1, 128, smem, st>>>(ma.map, mb.map, static_cast<const __half*>(a), static_cast<const __half*>(b), c, N, K, mode, fmt);
  }
  return cudaGetLastError();
}

}  // namespace sfa
