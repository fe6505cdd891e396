// Split-KV decode over the sink+window cache (reference: decode_kernel.py:28-226).
//
// Bandwidth-bound design for B200: one CTA per (split, kv-head, batch) serves ALL q heads of the
// GQA group from a single pass over K/V (the reference grids over q heads and re-reads each kv
// head H_q/H_kv times, decode_kernel.py:57-62,175).  Each of the 4 warps owns a private
// cp.async ring of 16-key blocks (16-byte coalesced requests, kStages blocks in flight per warp,
// no CTA-wide barrier in the loop); scores and PV run on mma.sync m16n8k16 with the q heads
// padded to the 16 MMA rows, online softmax uses quad shuffles.  The KV sequence may be given as
// two segments (sink buffer + ring window buffer, cache.py:185-216): softmax is order-invariant so
// the ring is read in place.  Partials (m, l, o) go to a workspace when splits > 1 and a small
// combine kernel folds in s_aux as the reference's virtual split (decode_kernel.py:205-224).
#include <stdlib.h>

#include "common.cuh"

namespace sfa {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* smem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
template <typename T> __device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int D> struct DecodeCfg {
  // keys per pipeline block: at D = 64 a 16-key block is only 4 KB of K+V and the per-block fixed work (softmax
  // update, shuffles, loop, cp.async wait) dominated (66 % of HBM peak vs 81 % at D = 128), so D = 64 uses 32 keys
  static constexpr int kKB = (D <= 64) ? 32 : 16;
  static constexpr int kStages = 3 - (D > 128 ? 1 : 0);     // x 4 warps x 2 x kBlkBytes: 96 KB -> 2 CTAs per SM
  static constexpr int kRowBytes = D * 2;
  static constexpr int kBlkBytes = kKB * kRowBytes;           // one block of K (or V)
  static constexpr int kWarpBytes = kStages * 2 * kBlkBytes;  // K+V ring of one warp
  static constexpr int kQBytes = 16 * kRowBytes;
  static constexpr int kMergeBytes = 4 * 16 * D * 4 + 4 * 16 * 2 * 4;
  static constexpr int kPipeBytes = 4 * kWarpBytes + kQBytes;
  static constexpr int kSmem = kPipeBytes > kMergeBytes ? kPipeBytes : kMergeBytes;
};

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-B chunk `c` of row `r` in a [16][D] 16-bit tile, XOR-swizzled for ldmatrix
template <int D> __device__ __forceinline__ int tile_off(int r, int c) { return r * (D * 2) + ((c ^ (r & 7)) << 4); }

template <typename T, int D>
__global__ void __launch_bounds__(128) decode_mma_kernel(DecodeParams p, int chunk_keys, int g_tiles) {
  using C = DecodeCfg<D>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int kvh = blockIdx.y / g_tiles, gt = blockIdx.y % g_tiles;
  const int b = blockIdx.z;
  const int G = p.Hq / p.Hkv;
  const int h0 = kvh * G + gt * 16;                 // first q head served by this CTA
  const int nh = min(16, G - gt * 16);              // real heads among the 16 MMA rows
  const int L = p.len[0] + p.len[1];
  const int k_begin = split * chunk_keys;
  const int k_end = min(L, k_begin + chunk_keys);

  unsigned char* q_s = smem + 4 * C::kWarpBytes;
  unsigned char* ring = smem + warp * C::kWarpBytes;

  // ---- Q tile [16][D] -> smem (rows >= nh are zero), then A fragments in registers
  for (int c = threadIdx.x; c < 16 * (D / 8); c += 128) {
    const int r = c / (D / 8), ch = c % (D / 8);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < nh) val = *reinterpret_cast<const uint4*>(static_cast<const T*>(p.q) + b * p.sq_b + (h0 + r) * p.sq_h + ch * 8);
    *reinterpret_cast<uint4*>(q_s + tile_off<D>(r, ch)) = val;
  }
  __syncthreads();
  uint32_t qf[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const int mat = lane >> 3, r = (lane & 7) + 8 * (mat & 1), ch = ks * 2 + (mat >> 1);
    ldsm_x4(qf[ks], q_s + tile_off<D>(r, ch));
  }

  const T* kb[2];
  const T* vb[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    kb[s] = static_cast<const T*>(p.k[s]) + b * p.sk[s].b + kvh * p.sk[s].h;
    vb[s] = static_cast<const T*>(p.v[s]) + b * p.sv[s].b + kvh * p.sv[s].h;
  }
  const int len0 = p.len[0];

  // blocks of kKB keys; warp w owns blocks w, w+4, ... of this split
  constexpr int KB = C::kKB;
  const int nblk_total = (k_end - k_begin + KB - 1) / KB;
  const int nblk = (nblk_total > warp) ? (nblk_total - warp + 3) / 4 : 0;

  auto load_block = [&](int it) {
    const int key0 = k_begin + (warp + 4 * it) * KB;
    unsigned char* ks_ = ring + (it % C::kStages) * 2 * C::kBlkBytes;
    unsigned char* vs_ = ks_ + C::kBlkBytes;
#pragma unroll
    for (int c = lane; c < KB * (D / 8); c += 32) {
      const int r = c / (D / 8), ch = c % (D / 8);
      const int key = key0 + r;
      const bool valid = key < k_end;
      const int seg = (key >= len0) ? 1 : 0;
      const int64_t pos = valid ? (seg ? key - len0 : key) : 0;
      const T* ksrc = kb[valid ? seg : 0] + pos * p.sk[valid ? seg : 0].n + ch * 8;
      const T* vsrc = vb[valid ? seg : 0] + pos * p.sv[valid ? seg : 0].n + ch * 8;
      cp_async16(ks_ + tile_off<D>(r, ch), ksrc, valid);
      cp_async16(vs_ + tile_off<D>(r, ch), vsrc, valid);
    }
  };

  float m_row[2] = {-INFINITY, -INFINITY};  // rows g and g+8 (log2 units)
  float l_row[2] = {0.f, 0.f};              // per-lane partial sums
  float o[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  const float sl2 = p.scale * kLog2e;

#pragma unroll
  for (int s = 0; s < C::kStages - 1; ++s) {
    if (s < nblk) load_block(s);
    cp_async_commit();
  }
  for (int it = 0; it < nblk; ++it) {
    if (it + C::kStages - 1 < nblk) load_block(it + C::kStages - 1);
    cp_async_commit();
    cp_async_wait<C::kStages - 1>();
    __syncwarp();
    const unsigned char* ks_ = ring + (it % C::kStages) * 2 * C::kBlkBytes;
    const unsigned char* vs_ = ks_ + C::kBlkBytes;
    const int key0 = k_begin + (warp + 4 * it) * KB;

    // S[16 heads x KB keys] = Q K^T   (KB/8 n-tiles of 8 keys)
    float sc[KB / 8][4];
#pragma unroll
    for (int n = 0; n < KB / 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
      for (int kb16 = 0; kb16 < KB / 16; ++kb16) {
        uint32_t bf[4];
        const int mat = lane >> 3, r = kb16 * 16 + (lane & 7) + 8 * (mat >> 1), ch = ks * 2 + (mat & 1);
        ldsm_x4(bf, ks_ + tile_off<D>(r, ch));
        mma16816<T>(sc[kb16 * 2], qf[ks], bf[0], bf[1]);
        mma16816<T>(sc[kb16 * 2 + 1], qf[ks], bf[2], bf[3]);
      }
    }
    // mask keys past the split end, move to log2 units
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < KB / 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key0 + n * 8 + 2 * (lane & 3) + (e & 1);
        const float s2 = (key < k_end) ? sc[n][e] * sl2 : -INFINITY;
        sc[n][e] = s2;
        mx[e >> 1] = fmaxf(mx[e >> 1], s2);
      }
    float alpha[2];
    bool moved = false;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_row[r], mx[r]);     // finite: every block holds >= 1 valid key
      moved |= (m_new != m_row[r]);
      alpha[r] = ex2_fast(m_row[r] - m_new);
      m_row[r] = m_new;
      l_row[r] *= alpha[r];
    }
    const bool rescale = __any_sync(0xffffffffu, moved);    // the running max rarely moves after the first blocks
    uint32_t pf[KB / 16][4];
#pragma unroll
    for (int kb16 = 0; kb16 < KB / 16; ++kb16) {
      float pv[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[n][e] = ex2_fast(sc[kb16 * 2 + n][e] - m_row[e >> 1]);
          l_row[e >> 1] += pv[n][e];
        }
      pf[kb16][0] = pack2<T>(pv[0][0], pv[0][1]);
      pf[kb16][1] = pack2<T>(pv[0][2], pv[0][3]);
      pf[kb16][2] = pack2<T>(pv[1][0], pv[1][1]);
      pf[kb16][3] = pack2<T>(pv[1][2], pv[1][3]);
    }
    // O[16 x D] = O*alpha + P V
#pragma unroll
    for (int n2 = 0; n2 < D / 16; ++n2) {
      if (rescale) {
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          float(&acc)[4] = o[n2 * 2 + hlf];
          acc[0] *= alpha[0];
          acc[1] *= alpha[0];
          acc[2] *= alpha[1];
          acc[3] *= alpha[1];
        }
      }
#pragma unroll
      for (int kb16 = 0; kb16 < KB / 16; ++kb16) {
        uint32_t vf[4];
        const int mat = lane >> 3, r = kb16 * 16 + (lane & 7) + 8 * (mat & 1), ch = n2 * 2 + (mat >> 1);
        ldsm_x4_t(vf, vs_ + tile_off<D>(r, ch));
        mma16816<T>(o[n2 * 2], pf[kb16], vf[0], vf[1]);
        mma16816<T>(o[n2 * 2 + 1], pf[kb16], vf[2], vf[3]);
      }
    }
    __syncwarp();
  }
  cp_async_wait<0>();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_row[r] += __shfl_xor_sync(0xffffffffu, l_row[r], 1);
    l_row[r] += __shfl_xor_sync(0xffffffffu, l_row[r], 2);
  }
  __syncthreads();  // all rings drained: reuse smem for the cross-warp merge

  float* mg_o = reinterpret_cast<float*>(smem);                      // [4][16][D]
  float* mg_ml = reinterpret_cast<float*>(smem) + 4 * 16 * D;        // [4][16][2]
  {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int n = 0; n < D / 8; ++n) {
      float* r0 = mg_o + ((warp * 16 + g) * D) + n * 8 + 2 * t;
      float* r1 = mg_o + ((warp * 16 + g + 8) * D) + n * 8 + 2 * t;
      r0[0] = o[n][0];
      r0[1] = o[n][1];
      r1[0] = o[n][2];
      r1[1] = o[n][3];
    }
    if (t == 0) {
      mg_ml[(warp * 16 + g) * 2 + 0] = m_row[0];
      mg_ml[(warp * 16 + g) * 2 + 1] = l_row[0];
      mg_ml[(warp * 16 + g + 8) * 2 + 0] = m_row[1];
      mg_ml[(warp * 16 + g + 8) * 2 + 1] = l_row[1];
    }
  }
  __syncthreads();
  const bool final_out = (p.splits == 1);
  for (int idx = threadIdx.x; idx < nh * D; idx += 128) {
    const int r = idx / D, d = idx % D;
    const int h = h0 + r;
    float mg = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) mg = fmaxf(mg, mg_ml[(w * 16 + r) * 2]);
    const float sa = (final_out && p.s_aux) ? p.s_aux[h] * kLog2e : -INFINITY;
    mg = fmaxf(mg, sa);
    float lg = (sa == -INFINITY) ? 0.f : exp2f(sa - mg);
    float og = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = mg_ml[(w * 16 + r) * 2];
      const float a = (mw == -INFINITY) ? 0.f : exp2f(mw - mg);
      lg += mg_ml[(w * 16 + r) * 2 + 1] * a;
      og += mg_o[(w * 16 + r) * D + d] * a;
    }
    if (final_out) {
      static_cast<T*>(p.o)[b * p.so_b + h * p.so_h + d] = from_f<T>(og / fmaxf(lg, 1e-8f));
    } else {
      const int64_t pr = ((int64_t)b * p.Hq + h) * p.splits + split;
      p.part_o[pr * D + d] = og;
      if (d == 0) {
        p.part_ml[pr * 2 + 0] = mg;
        p.part_ml[pr * 2 + 1] = lg;
      }
    }
  }
}

// Phase 2 (decode_kernel.py:201-226): merge the split partials with the s_aux virtual split.
template <typename T>
__global__ void decode_combine_kernel(DecodeParams p) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int64_t base = ((int64_t)b * p.Hq + h) * p.splits;
  const float sa = p.s_aux ? p.s_aux[h] * kLog2e : -INFINITY;
  float mg = sa;
  for (int s = 0; s < p.splits; ++s) mg = fmaxf(mg, p.part_ml[(base + s) * 2]);
  float lg = (sa == -INFINITY) ? 0.f : exp2f(sa - mg);
  for (int s = 0; s < p.splits; ++s) {
    const float ms = p.part_ml[(base + s) * 2];
    lg += (ms == -INFINITY) ? 0.f : p.part_ml[(base + s) * 2 + 1] * exp2f(ms - mg);
  }
  lg = fmaxf(lg, 1e-8f);
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float og = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float ms = p.part_ml[(base + s) * 2];
      if (ms != -INFINITY) og += p.part_o[(base + s) * p.D + d] * exp2f(ms - mg);
    }
    static_cast<T*>(p.o)[b * p.so_b + h * p.so_h + d] = from_f<T>(og / lg);
  }
}

template <typename T, int D>
cudaError_t launch(const DecodeParams& p, cudaStream_t st) {
  using C = DecodeCfg<D>;
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = ensure_dyn_smem(decode_mma_kernel<T, D>, C::kSmem, attr_done)) return e;
  const int G = p.Hq / p.Hkv;
  const int g_tiles = (G + 15) / 16;
  const int L = p.len[0] + p.len[1];
  int chunk = (L + p.splits - 1) / p.splits;
  chunk = ((chunk + 63) / 64) * 64;
  dim3 grid(p.splits, p.Hkv * g_tiles, p.B);
  decode_mma_kernel<T, D><<<grid, 128, C::kSmem, st>>>(p, chunk, g_tiles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || p.splits == 1) return e;
  decode_combine_kernel<T><<<dim3(p.Hq, p.B), (D < 128 ? 64 : 128), 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

bool mma_decode_supported(const DecodeParams& p, int dtype) {
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return false;
  if (p.D != 64 && p.D != 128 && p.D != 256) return false;
  // 16-byte vector loads: rows must be 16-B aligned
  auto ok = [](const void* ptr, int64_t a, int64_t b, int64_t c) {
    return (reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && (a % 8 == 0) && (b % 8 == 0) && (c % 8 == 0);
  };
  if (!ok(p.q, p.sq_b, p.sq_h, 8)) return false;
  for (int s = 0; s < 2; ++s) {
    if (p.len[s] <= 0) continue;
    if (!ok(p.k[s], p.sk[s].b, p.sk[s].h, p.sk[s].n)) return false;
    if (!ok(p.v[s], p.sv[s].b, p.sv[s].h, p.sv[s].n)) return false;
  }
  return true;
}

// Split count.  The kernel runs 2 CTAs per SM; B*Hkv*splits CTAs execute in ceil(ctas / 296) rounds and a partly
// filled last round wastes bandwidth, but every extra split costs a pipeline fill plus partial (m, l, o) traffic
// and any split > 1 costs the combine launch.  Measured at BASELINE configs[3] (B*Hkv = 512, 4100 keys, D = 64):
// splits 1 / 2 / 3 / 4 / 6 / 8 -> 112.6 / 118.8 / 120.8 / 118.8 / 127.0 / 133.1 us.  The cost model below
// (round efficiency x 4 % per extra split + 6 % for the second kernel) reproduces that ordering and still
// splits small batches enough to fill the machine.
int mma_decode_splits(int B, int Hkv, int total_len) {
  static const int forced = getenv("SFA_DECODE_SPLITS") ? atoi(getenv("SFA_DECODE_SPLITS")) : 0;   // experiments only
  if (forced > 0) return forced;
  const long long base = (long long)B * Hkv;
  int max_s = (total_len + 255) / 256;
  if (max_s > 64) max_s = 64;
  if (max_s < 1) max_s = 1;
  const double slots = 148.0 * 2.0;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= max_s; ++s) {
    const double rounds = (double)(base * s) / slots;
    const double eff = rounds / (double)(long long)(rounds + 0.999999);     // filled fraction of the rounds
    const double cost = (1.0 / eff) * (1.0 + 0.04 * (s - 1)) + (s > 1 ? 0.06 : 0.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

cudaError_t mma_decode(const DecodeParams& p, int dtype, cudaStream_t st) {
#define SFA_DEC(T)                                     \
  switch (p.D) {                                       \
    case 64: return launch<T, 64>(p, st);              \
    case 128: return launch<T, 128>(p, st);            \
    case 256: return launch<T, 256>(p, st);            \
  }
  if (dtype == SFA_DTYPE_BF16) { SFA_DEC(__nv_bfloat16) }
  if (dtype == SFA_DTYPE_FP16) { SFA_DEC(__half) }
#undef SFA_DEC
  return cudaErrorInvalidValue;
}

}  // namespace sfa
