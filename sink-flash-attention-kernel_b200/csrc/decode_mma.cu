// Split-KV decode over the sink+window cache (reference: decode_kernel.py:28-226).
//
// Bandwidth-bound design for B200: a CTA serves ALL q heads of the GQA group from a single pass over its K/V range (the reference grids over q heads and re-reads each kv
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

// V: pipeline variant (0 = default).  keys per pipeline block / ring depth are a trade between per-block fixed work
// (softmax update, shuffles, loop, cp.async wait), bytes in flight, and CTAs per SM (shared memory bound).
template <int D, int V = 0> struct DecodeCfg {
  static constexpr int kKB = (V == 2 || V == 3) ? 16 : (V == 4 ? 64 : ((D <= 64) ? 32 : 16));
  static constexpr int kStages = (V == 1) ? 2 : (V == 3 ? 4 : (V == 4 ? 2 : 3 - (D > 128 ? 1 : 0)));
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

// Work decomposition: the B * Hkv * g_tiles units (one KV head of one batch row, 16 q heads) x L keys are laid end to
// end and cut into gridDim.x EQUAL key ranges, one per CTA, all resident at once (2 CTAs per SM).  The round-1 grid of
// one CTA per unit ran BASELINE configs[3] (512 units) as 1.73 waves of 296 CTAs -- 13.5 % of the machine idle in the
// second wave.  A CTA's range covers the tail of one unit, possibly whole units, and the head of another: every
// (unit, key range) segment is one pass of the pipeline below; a unit cut into pieces leaves (m, l, o) partials that
// the combine kernel merges, a unit inside one range is finished here.
// kPaged: per-batch cache lengths (seq_lens) and / or a page table (block_table) -- a separate instantiation, the plain
// kernel is unchanged.  The key ranges are planned over the maximal length; a unit's range is clipped to its row's length
// (ranges past it leave empty partials), and every 32-key block lies inside one page (page_size % 32 == 0, ranges start
// on page boundaries).
template <typename T, int D, int V, bool kPaged = false>
__global__ void __launch_bounds__(128) decode_mma_kernel(DecodeParams p, int chunk_keys, int g_tiles) {
  using C = DecodeCfg<D, V>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = p.Hq / p.Hkv;
  const int L = p.len[0] + p.len[1];
  const int64_t f_total = static_cast<int64_t>(p.B) * p.Hkv * g_tiles * L;
  int64_t f = static_cast<int64_t>(blockIdx.x) * chunk_keys;
  const int64_t f_end = min(f + chunk_keys, f_total);
  unsigned char* q_s = smem + 4 * C::kWarpBytes;
  unsigned char* ring = smem + warp * C::kWarpBytes;
  const float sl2 = p.scale * kLog2e;
  constexpr int KB = C::kKB;

  while (f < f_end) {
  const int unit = static_cast<int>(f / L);
  const int k_begin = static_cast<int>(f - static_cast<int64_t>(unit) * L);
  const int64_t k_lim = static_cast<int64_t>(k_begin) + (f_end - f);
  const int k_end = k_lim < L ? static_cast<int>(k_lim) : L;
  f += k_end - k_begin;
  const int gt = unit % g_tiles;
  const int kvh = (unit / g_tiles) % p.Hkv, b = unit / (g_tiles * p.Hkv);
  int k_stop = k_end;                               // keys [k_begin, k_stop) exist for this batch row
  if constexpr (kPaged) {
    if (p.seq_lens != nullptr) k_stop = min(k_end, max(__ldg(p.seq_lens + b), k_begin));
  }
  const int h0 = kvh * G + gt * 16;                 // first q head served by this segment
  const int nh = min(16, G - gt * 16);              // real heads among the 16 MMA rows
  // pieces of this unit: CTAs first_cta .. last_cta touch it
  const int first_cta = static_cast<int>((static_cast<int64_t>(unit) * L) / chunk_keys);
  const int last_cta = static_cast<int>((static_cast<int64_t>(unit + 1) * L - 1) / chunk_keys);
  const int piece = static_cast<int>(blockIdx.x) - first_cta;

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
    // (64-bit batch term: a paged pool or a long contiguous cache exceeds 2^31 elements)
    kb[s] = static_cast<const T*>(p.k[s]) + ((kPaged && p.block_table != nullptr) ? 0 : static_cast<int64_t>(b) * p.sk[s].b) + kvh * p.sk[s].h;
    vb[s] = static_cast<const T*>(p.v[s]) + ((kPaged && p.block_table != nullptr) ? 0 : static_cast<int64_t>(b) * p.sv[s].b) + kvh * p.sv[s].h;
  }
  const int len0 = p.len[0];

  // blocks of kKB keys; warp w owns blocks w, w+4, ... of this segment
  const int nblk_total = (k_stop - k_begin + KB - 1) / KB;
  const int nblk = (nblk_total > warp) ? (nblk_total - warp + 3) / 4 : 0;

  // per-lane constants of the block copy: lane owns 16-byte chunk `lch` of rows lr0, lr0 + 32 / (D / 8), ...
  constexpr int kCPR = D / 8;                 // 16-byte chunks per row
  constexpr int kRowsPerIt = 32 / kCPR;       // rows covered by one warp-wide cp.async (D <= 256)
  const int lch = lane % kCPR, lr0 = lane / kCPR;
  auto load_block = [&](int it) {
    const int key0 = k_begin + (warp + 4 * it) * KB;
    unsigned char* ks_ = ring + (it % C::kStages) * 2 * C::kBlkBytes;
    unsigned char* vs_ = ks_ + C::kBlkBytes;
    if constexpr (kPaged) {
      // one page per block: logical page -> physical page, then the same copy as below inside the page
      int64_t pos0 = key0;
      int64_t kpage = 0, vpage = 0;
      if (p.block_table != nullptr) {
        const int phys = __ldg(p.block_table + b * p.bt_stride + (key0 >> p.lg_page));
        pos0 = key0 & (p.page_size - 1);
        kpage = static_cast<int64_t>(phys) * p.sk[0].b;
        vpage = static_cast<int64_t>(phys) * p.sv[0].b;
      }
      const T* ksrc = kb[0] + kpage + (pos0 + lr0) * p.sk[0].n + lch * 8;
      const T* vsrc = vb[0] + vpage + (pos0 + lr0) * p.sv[0].n + lch * 8;
      const int64_t kstep = kRowsPerIt * p.sk[0].n, vstep = kRowsPerIt * p.sv[0].n;
#pragma unroll
      for (int j = 0; j < KB / kRowsPerIt; ++j) {
        const int r = lr0 + j * kRowsPerIt;
        const bool valid = key0 + r < k_stop;           // rows past the row's length: zero fill, never dereferenced
        cp_async16(ks_ + tile_off<D>(r, lch), valid ? ksrc : kb[0], valid);
        cp_async16(vs_ + tile_off<D>(r, lch), valid ? vsrc : vb[0], valid);
        ksrc += kstep;
        vsrc += vstep;
      }
      return;
    }
    const bool in0 = key0 + KB <= len0, in1 = key0 >= len0;
    if (key0 + KB <= k_end && (in0 || in1)) {
      // fast path (all but the edge blocks): the block lies inside one segment and inside the range -- one pointer
      // per lane, advanced by a constant stride; no per-chunk bounds / segment arithmetic (the instruction count of
      // the copy loop, not HBM, limited the head_dim-64 kernel: 77 % of the copy bandwidth vs 82 % at head_dim 128)
      const int seg = in0 ? 0 : 1;
      const int64_t pos0 = (seg ? key0 - len0 : key0) + lr0;
      const T* ksrc = kb[seg] + pos0 * p.sk[seg].n + lch * 8;
      const T* vsrc = vb[seg] + pos0 * p.sv[seg].n + lch * 8;
      const int64_t kstep = kRowsPerIt * p.sk[seg].n, vstep = kRowsPerIt * p.sv[seg].n;
#pragma unroll
      for (int j = 0; j < KB / kRowsPerIt; ++j) {
        const int r = lr0 + j * kRowsPerIt;
        cp_async16(ks_ + tile_off<D>(r, lch), ksrc, true);
        cp_async16(vs_ + tile_off<D>(r, lch), vsrc, true);
        ksrc += kstep;
        vsrc += vstep;
      }
      return;
    }
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
        const float s2 = (key < k_stop) ? sc[n][e] * sl2 : -INFINITY;
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
  const bool final_out = (first_cta == last_cta);      // the whole unit sits inside this CTA's range
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
      const int64_t pr = ((int64_t)b * p.Hq + h) * p.splits + piece;
      p.part_o[pr * D + d] = og;
      if (d == 0) {
        p.part_ml[pr * 2 + 0] = mg;
        p.part_ml[pr * 2 + 1] = lg;
      }
    }
  }
  __syncthreads();      // the merge buffers alias the rings and the Q tile of the next segment
  }
}

// Phase 2 (decode_kernel.py:201-226): merge the split partials with the s_aux virtual split.
template <typename T>
__global__ void decode_combine_kernel(DecodeParams p, int chunk_keys, int g_tiles) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int G = p.Hq / p.Hkv;
  const int L = p.len[0] + p.len[1];
  const int kvh = h / G, gt = (h % G) / 16;
  const int64_t unit = (static_cast<int64_t>(b) * p.Hkv + kvh) * g_tiles + gt;
  const int first_cta = static_cast<int>((unit * L) / chunk_keys);
  const int last_cta = static_cast<int>(((unit + 1) * L - 1) / chunk_keys);
  const int np = last_cta - first_cta + 1;
  if (np == 1) return;                       // finished by the decode kernel itself
  const int64_t base = ((int64_t)b * p.Hq + h) * p.splits;
  const float sa = p.s_aux ? p.s_aux[h] * kLog2e : -INFINITY;
  float mg = sa;
  for (int s = 0; s < np; ++s) mg = fmaxf(mg, p.part_ml[(base + s) * 2]);
  float lg = (sa == -INFINITY) ? 0.f : exp2f(sa - mg);
  for (int s = 0; s < np; ++s) {
    const float ms = p.part_ml[(base + s) * 2];
    lg += (ms == -INFINITY) ? 0.f : p.part_ml[(base + s) * 2 + 1] * exp2f(ms - mg);
  }
  lg = fmaxf(lg, 1e-8f);
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float og = 0.f;
    for (int s = 0; s < np; ++s) {
      const float ms = p.part_ml[(base + s) * 2];
      if (ms != -INFINITY) og += p.part_o[(base + s) * p.D + d] * exp2f(ms - mg);
    }
    static_cast<T*>(p.o)[b * p.so_b + h * p.so_h + d] = from_f<T>(og / lg);
  }
}

// Key range per CTA and the worst-case number of pieces a unit is cut into (= workspace slots per head).
// One CTA per 2-CTA-per-SM slot when there is enough work, never less than kMinChunk keys per CTA.
struct DecodePlan {
  int chunk, ncta, max_pieces;
};
DecodePlan make_decode_plan(int B, int Hq, int Hkv, int L, int align = 1) {
  constexpr int kMinChunk = 256;
  const int g_tiles = (Hq / Hkv + 15) / 16;
  const int64_t total = static_cast<int64_t>(B) * Hkv * g_tiles * L;
  static const int forced = getenv("SFA_DECODE_CTAS") ? atoi(getenv("SFA_DECODE_CTAS")) : 0;   // experiments only
  const int64_t units = static_cast<int64_t>(B) * Hkv * g_tiles;
  const int64_t slots = static_cast<int64_t>(device_sm_count()) * 2;
  int64_t chunk;
  if (forced > 0) {
    chunk = (total + forced - 1) / forced;
  } else if (units >= slots && (units % slots == 0 || units % slots >= slots / 2 || units >= 4 * slots)) {
    // enough whole units to keep every slot busy: one unit per CTA, no cut, no partials, no combine launch.  The
    // kernel is bandwidth-bound, so the partly filled last wave costs little (its CTAs get the whole memory system):
    // measured at BASELINE configs[3] (512 units on 296 slots) 106.5 us this way vs 112.6 us with 296 equal ranges.
    chunk = L;
  } else {
    chunk = (total + slots - 1) / slots;          // few units (small batch) or a thin last wave: equal key ranges
  }
  if (align > 1 && chunk % align != 0) chunk += align - chunk % align;     // paged: ranges start on page boundaries
  DecodePlan pl;
  pl.chunk = static_cast<int>(chunk);
  pl.ncta = static_cast<int>((total + chunk - 1) / chunk);
  pl.max_pieces = static_cast<int>((L + chunk - 1) / chunk) + 1;
  return pl;
}

template <typename T, int D, int V = 0, bool kPaged = false>
cudaError_t launch(const DecodeParams& p, cudaStream_t st) {
  using C = DecodeCfg<D, V>;
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = ensure_dyn_smem(decode_mma_kernel<T, D, V, kPaged>, C::kSmem, attr_done)) return e;
  const int G = p.Hq / p.Hkv;
  const int g_tiles = (G + 15) / 16;
  const int L = p.len[0] + p.len[1];
  const DecodePlan pl = make_decode_plan(p.B, p.Hq, p.Hkv, L, (kPaged && p.block_table != nullptr) ? p.page_size : 1);
  if (pl.max_pieces > p.splits) return cudaErrorInvalidValue;      // workspace carved for fewer pieces
  decode_mma_kernel<T, D, V, kPaged><<<pl.ncta, 128, C::kSmem, st>>>(p, pl.chunk, g_tiles);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // units cut by a range boundary are merged here (every unit when the ranges are shorter than a unit; none when
  // the ranges happen to end on unit boundaries)
  const bool any_cut = !(pl.chunk % L == 0);
  if (!any_cut) return e;
  decode_combine_kernel<T><<<dim3(p.Hq, p.B), (D < 128 ? 64 : 128), 0, st>>>(p, pl.chunk, g_tiles);
  return cudaGetLastError();
}

}  // namespace

bool mma_decode_supported(const DecodeParams& p, int dtype) {
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return false;
  if (p.D != 64 && p.D != 128 && p.D != 256) return false;
  if (p.paged && p.block_table != nullptr && (p.page_size < 32 || (p.page_size & (p.page_size - 1)) != 0)) return false;
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

// Workspace slots per head (= the most pieces any unit can be cut into) for the balanced decomposition above.
// History: with one CTA per (unit, split) BASELINE configs[3] measured 112.6 / 118.8 / 120.8 / 118.8 / 127.0 / 133.1 us
// for 1 / 2 / 3 / 4 / 6 / 8 uniform splits -- every uniform choice leaves a partly filled last wave.
int mma_decode_splits(int B, int Hq, int Hkv, int total_len, int align) {
  return make_decode_plan(B, Hq, Hkv, total_len, align).max_pieces;
}

cudaError_t mma_decode(const DecodeParams& p, int dtype, cudaStream_t st) {
  if (p.paged) {
#define SFA_DECP(T)                                            \
  switch (p.D) {                                               \
    case 64: return launch<T, 64, 0, true>(p, st);             \
    case 128: return launch<T, 128, 0, true>(p, st);           \
    case 256: return launch<T, 256, 0, true>(p, st);           \
  }
    if (dtype == SFA_DTYPE_BF16) { SFA_DECP(__nv_bfloat16) }
    if (dtype == SFA_DTYPE_FP16) { SFA_DECP(__half) }
#undef SFA_DECP
    return cudaErrorInvalidValue;
  }
  static const int variant = getenv("SFA_DECODE_VARIANT") ? atoi(getenv("SFA_DECODE_VARIANT")) : 0;   // experiments only
#define SFA_DEC(T)                                     \
  switch (p.D) {                                       \
    case 64:                                           \
      if (variant == 1) return launch<T, 64, 1>(p, st); \
      if (variant == 2) return launch<T, 64, 2>(p, st); \
      if (variant == 3) return launch<T, 64, 3>(p, st); \
      if (variant == 4) return launch<T, 64, 4>(p, st); \
      return launch<T, 64>(p, st);                     \
    case 128: return launch<T, 128>(p, st);            \
    case 256: return launch<T, 256>(p, st);            \
  }
  if (dtype == SFA_DTYPE_BF16) { SFA_DEC(__nv_bfloat16) }
  if (dtype == SFA_DTYPE_FP16) { SFA_DEC(__half) }
#undef SFA_DEC
  return cudaErrorInvalidValue;
}

}  // namespace sfa
