// Forward sink attention on the 5th-gen tensor cores (reference kernel being replaced:
// _sink_flash_attn_fwd_kernel, sink_flash_attention.py:93-194).
//
// Tile = 128 MMA rows = G q-heads of one GQA group x P consecutive positions (G*P = 128), so a
// narrow window costs W+P-1 key columns per tile instead of W+127 and K/V are fetched once per
// group instead of once per q head.  Two-range KV walk with run-time bounds: sink tiles
// [0, S) first, then the causal band [max(q0-W+1,S), q0+P-1] in unaligned BN-row tiles (TMA boxes
// may start at any row; out-of-range rows are zero-filled and masked).
//
// Warp roles (192 threads): warps 0-3 softmax/epilogue (one MMA row per thread, TMEM lane ==
// row), warp 4 TMA producer, warp 5 tcgen05.mma issuer.
//   S = Q K^T      SS-form UMMA, K-major operands from SWIZZLE_128B TMA tiles, fp32 in TMEM
//   P              bf16/fp16, written over S in TMEM (tcgen05.st) -> A operand of the TS-form
//   O += P V       V consumed MN-major straight from its TMA tile; O stays in TMEM
// Online softmax is seeded with (m,l) = (s_aux,1) (sink_flash_attention.py:139-146), works in
// exp2 units with 1/sqrt(D)*log2(e) folded into one FFMA, and rescales O lazily (only when the
// running max moved by more than 2^8).  Epilogue: O/l -> 16-bit -> swizzled smem -> TMA store;
// LSE = m*ln2 + log(l) (natural log, sink term included, :192).
#include "attn_common.cuh"
#include "tmap.cuh"

namespace sfa {
namespace {

template <int D> struct FwdCfg {
  static constexpr int kDS = D / 64;                       // 64-channel slabs (one SWIZZLE_128B row each)
  static constexpr int kBNMax = (D == 64) ? 160 : 128;     // KV rows per tile (UMMA N of S), multiple of 16
  static constexpr int kStages = (D == 64) ? 2 : 1;        // K and V rings
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBNMax * D * 2;
  static constexpr int kSlabQ = 128 * 128;                 // bytes of one Q slab
  static constexpr int kSlabKV = kBNMax * 128;
  static constexpr uint32_t kTmemCols = 256;
  static constexpr uint32_t kColS = 0;                     // S (fp32) and P (16-bit, aliased) columns
  static constexpr uint32_t kColO = kBNMax;                // O accumulator columns
  static constexpr int kSmem = 1024 + kQBytes + 2 * kStages * kKVBytes + 256;
  static_assert(kBNMax + D <= 256, "TMEM budget");
};

struct FwdArgs {
  int N, S, W, Hq, G, P, BN, groups_per_kv;   // groups_per_kv = (Hq/Hkv)/G
  int q_swap, k_swap, v_swap, o_swap;
  int fmt;                                     // 0 f16, 1 bf16
  float sl2;                                   // scale * log2(e)
  const float* s_aux;
  float* lse;
  const int* seq_lo;     // packed sequences (no sink tokens): first key of the row's sequence; nullptr: none
  int64_t seq_bs;
};

template <typename T, int D>
__global__ void __launch_bounds__(192) fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                  const __grid_constant__ CUtensorMap tmK,
                                                  const __grid_constant__ CUtensorMap tmV,
                                                  const __grid_constant__ CUtensorMap tmO, const FwdArgs a) {
  using C = FwdCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = smem;
  unsigned char* k_s = q_s + C::kQBytes;
  unsigned char* v_s = k_s + C::kStages * C::kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + C::kStages * C::kKVBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* o_done = bars + 3;
  uint64_t* k_full = bars + 4;
  uint64_t* k_empty = k_full + C::kStages;
  uint64_t* v_full = k_empty + C::kStages;
  uint64_t* v_empty = v_full + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + C::kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * a.P;
  const int kvh = blockIdx.y / a.groups_per_kv;
  const int hq0 = blockIdx.y * a.G;     // == kvh*group + (blockIdx.y % groups_per_kv)*G
  const int b = blockIdx.z;
  TilePlan pl = make_plan(q0, a.P, a.N, a.S, a.W, a.BN);
  if (a.seq_lo != nullptr) clamp_plan(pl, a.seq_lo[b * a.seq_bs + q0], a.W, a.BN, bn_magic(a.BN));

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, C::kQBytes);
      for (int s = 0; s < C::kDS; ++s) tma_tile(q_s + s * C::kSlabQ, &tmQ, q_full, a.q_swap, s * 64, q0, hq0, b);
      for (int t = 0; t < pl.n_tiles; ++t) {
        int kstart, cols; bool is_sink;
        pl.tile(t, a.BN, kstart, cols, is_sink);
        const int st = t % C::kStages;
        const uint32_t ph = (t / C::kStages) & 1;
        mbar_wait(k_empty + st, ph ^ 1);
        mbar_expect_tx(k_full + st, a.BN * D * 2);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(k_s + st * C::kKVBytes + s * C::kSlabKV, &tmK, k_full + st, a.k_swap, s * 64, kstart, kvh, b);
        mbar_wait(v_empty + st, ph ^ 1);
        mbar_expect_tx(v_full + st, a.BN * D * 2);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(v_s + st * C::kKVBytes + s * C::kSlabKV, &tmV, v_full + st, a.v_swap, s * 64, kstart, kvh, b);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_pv = make_idesc(a.fmt, 128, D, 0, 1);
      mbar_wait(q_full, 0);
      for (int t = 0; t < pl.n_tiles; ++t) {
        int kstart, cols; bool is_sink;
        pl.tile(t, a.BN, kstart, cols, is_sink);
        const int st = t % C::kStages;
        const uint32_t ph = (t / C::kStages) & 1;
        mbar_wait(k_full + st, ph);
        tc_fence_after();
        const uint32_t idesc_s = make_idesc(a.fmt, 128, cols, 0, 0);
        const uint32_t qa = smem_u32(q_s), ka = smem_u32(k_s + st * C::kKVBytes);
#pragma unroll
        for (int s = 0; s < C::kDS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + C::kColS, make_sdesc(qa + s * C::kSlabQ + kk * 32, 16, 1024),
                    make_sdesc(ka + s * C::kSlabKV + kk * 32, 16, 1024), idesc_s, (s | kk) != 0);
        umma_commit(k_empty + st);
        umma_commit(s_full);
        mbar_wait(p_full, t & 1);
        tc_fence_after();
        mbar_wait(v_full + st, ph);
        tc_fence_after();
        const uint32_t va = smem_u32(v_s + st * C::kKVBytes);
        for (int kk = 0; kk < cols / 16; ++kk)
          umma_ts(tmem + C::kColO, tmem + C::kColS + kk * 8, make_sdesc(va + kk * 2048, C::kSlabKV, 1024), idesc_pv,
                  (t > 0 || kk > 0));
        umma_commit(v_empty + st);
        umma_commit(o_done);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (128 threads)
    const int r = threadIdx.x;                       // MMA row == TMEM lane
    const int pr = a.q_swap ? (r / a.G) : (r % a.P);  // position within the tile
    const int gr = a.q_swap ? (r % a.G) : (r / a.P);  // head within the packed group
    const int i = q0 + pr;
    const int h = hq0 + gr;
    const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float m_used = a.s_aux ? a.s_aux[h] * kLog2e : -INFINITY;
    float l = a.s_aux ? 1.f : 0.f;

    for (int t = 0; t < pl.n_tiles; ++t) {
      int kstart, cols; bool is_sink;
      pl.tile(t, a.BN, kstart, cols, is_sink);
      int c_lo, c_hi;   // attended columns of this row inside the tile: [c_lo, c_hi]
      row_range(is_sink, i, kstart, cols, a.S, a.W, c_lo, c_hi);
      if (a.seq_lo != nullptr && i < a.N) c_lo = max(c_lo, __ldg(a.seq_lo + b * a.seq_bs + i) - kstart);
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      // pass 1: row max over the attended columns
      float mx = -INFINITY;
      for (int c0 = 0; c0 < cols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tl + C::kColS + c0, v);
        tmem_ld_wait();
        if (c0 + 15 >= c_lo && c0 <= c_hi) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int c = c0 + e;
            const float s = __uint_as_float(v[e]);
            mx = (c >= c_lo && c <= c_hi) ? fmaxf(mx, s) : mx;
          }
        }
      }
      const float m_new = fmaxf(m_used, mx * a.sl2);
      const bool need = (m_new - m_used) > 8.0f;        // also true for -inf -> finite; false for NaN (-inf - -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? exp2f(m_used - m_new) : 1.f;
        if (need) {
          l *= alpha;
          m_used = m_new;
        }
        if (t > 0) {
          mbar_wait(o_done, (t - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tl + C::kColO + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st16(tl + C::kColO + c0, v);
          }
          tmem_st_wait();
        }
      }
      // pass 2: P = exp2(s*sl2 - m) -> 16-bit -> TMEM (over S), row sum
      const float neg_m = -m_used;
      float lsum = 0.f;
      for (int c0 = 0; c0 < cols; c0 += 16) {
        uint32_t v[16];
        uint32_t pk[8];
        tmem_ld16(tl + C::kColS + c0, v);
        tmem_ld_wait();
        if (c0 + 15 >= c_lo && c0 <= c_hi) {
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const int c = c0 + e;
            float p0 = fast_exp2(fmaf(__uint_as_float(v[e]), a.sl2, neg_m));
            float p1 = fast_exp2(fmaf(__uint_as_float(v[e + 1]), a.sl2, neg_m));
            p0 = (c >= c_lo && c <= c_hi) ? p0 : 0.f;
            p1 = (c + 1 >= c_lo && c + 1 <= c_hi) ? p1 : 0.f;
            lsum += p0 + p1;
            pk[e >> 1] = pack16_fast<T>(p0, p1);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) pk[e] = 0u;
        }
        tmem_st8(tl + C::kColS + (c0 >> 1), pk);
      }
      l += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // ---------------- epilogue
    mbar_wait(q_full, 0);   // Q smem is reused as the O staging buffer
    if (pl.n_tiles > 0) {
      mbar_wait(o_done, (pl.n_tiles - 1) & 1);
      tc_fence_after();
    }
    const float inv = (l > 0.f) ? 1.f / l : 0.f;
    const int ro = a.o_swap ? (pr * a.G + gr) : (gr * a.P + pr);   // row in O's box order
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint32_t v[16];
      if (pl.n_tiles > 0) {
        tmem_ld16(tl + C::kColO + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0u;
      }
      uint32_t pk[8];
#pragma unroll
      for (int e = 0; e < 16; e += 2)
        pk[e >> 1] = pack16<T>(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv);
      unsigned char* slab = q_s + (c0 >> 6) * C::kSlabQ;
      const int ch = (c0 & 63) >> 3;   // 16-byte chunk inside the 128-byte row
      *reinterpret_cast<uint4*>(slab + sw128_off(ro, ch)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(slab + sw128_off(ro, ch + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if (i < a.N) a.lse[(static_cast<int64_t>(b) * a.Hq + h) * a.N + i] = (l > 0.f) ? m_used * kLn2 + logf(l) : -INFINITY;
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      for (int s = 0; s < C::kDS; ++s) {
        if (a.o_swap) tma_store_4d(&tmO, q_s + s * C::kSlabQ, s * 64, hq0, q0, b);
        else tma_store_4d(&tmO, q_s + s * C::kSlabQ, s * 64, q0, hq0, b);
      }
      tma_store_commit();
      tma_store_wait_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, C::kTmemCols);
}

template <typename T, int D>
cudaError_t launch_fwd(const AttnParams& p, int dtype, cudaStream_t st) {
  using C = FwdCfg<D>;
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = ensure_dyn_smem(fwd_kernel<T, D>, C::kSmem, attr_done)) return e;
  const int group = p.Hq / p.Hkv;
  int G, P;
  pick_packing(p.Hq, p.Hkv, G, P);
  const int BN = pick_bn(p.W, p.N, P, C::kBNMax);

  TileMap mq, mk, mv, mo;
  if (!make_tile_map(&mq, p.q, dtype, p.D, p.N, p.Hq, p.B, p.sq, P, G)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mk, p.k, dtype, p.D, p.N, p.Hkv, p.B, p.sk, BN, 1)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mv, p.v, dtype, p.D, p.N, p.Hkv, p.B, p.sv, BN, 1)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mo, p.o, dtype, p.D, p.N, p.Hq, p.B, p.so, P, G)) return cudaErrorInvalidValue;

  FwdArgs a;
  a.N = p.N; a.S = p.S; a.W = p.W; a.Hq = p.Hq; a.G = G; a.P = P; a.BN = BN;
  a.groups_per_kv = group / G;
  a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh; a.o_swap = mo.swap_nh;
  a.fmt = (dtype == SFA_DTYPE_BF16) ? 1 : 0;
  a.sl2 = p.scale * kLog2e;
  a.s_aux = p.s_aux;
  a.lse = p.lse;
  a.seq_lo = p.seq_lo; a.seq_bs = p.seq_bs;
  dim3 grid((p.N + P - 1) / P, p.Hq / G, p.B);
  fwd_kernel<T, D><<<grid, 192, C::kSmem, st>>>(mq.map, mk.map, mv.map, mo.map, a);
  return cudaGetLastError();
}

}  // namespace

bool tc_fwd_supported(const AttnParams& p, int dtype) {
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return false;
  // 64 < D <= 128 in steps of 8 (80, 96, 112: north star "head_dim 64/80/128") runs on the head_dim-128 kernel with
  // TMA zero fill for the missing channels (the reference's Triton kernel cannot run these: tl.arange needs 2^k)
  if (p.D != 64 && !(p.D > 64 && p.D <= 128 && p.D % 8 == 0)) return false;
  if (p.N < 1) return false;
  return tma_compatible(p.q, p.sq, p.B, p.Hq, p.N) && tma_compatible(p.k, p.sk, p.B, p.Hkv, p.Nkv) && tma_compatible(p.v, p.sv, p.B, p.Hkv, p.Nkv) &&
         tma_compatible(p.o, p.so, p.B, p.Hq, p.N);
}

cudaError_t tc_fwd(const AttnParams& p, int dtype, cudaStream_t st) {
  if (dtype == SFA_DTYPE_BF16) return p.D == 64 ? launch_fwd<__nv_bfloat16, 64>(p, dtype, st) : launch_fwd<__nv_bfloat16, 128>(p, dtype, st);
  return p.D == 64 ? launch_fwd<__half, 64>(p, dtype, st) : launch_fwd<__half, 128>(p, dtype, st);
}

}  // namespace sfa
