// Persistent forward for 64 < head_dim <= 128, and for head_dim 64 with wide windows (reference kernel being replaced: _sink_flash_attn_fwd_kernel,
// sink_flash_attention.py:93-194).  At head_dim 128 the tensor pipe and the MUFU pipe need the same number of cycles
// per score (4 D flops against one exp2), so the kernel is built around keeping BOTH busy: a CTA works on TWO packed
// query tiles at once (A: position block 2k, B: block 2k + 1 of the same heads -- they share every K / V tile), each
// with its own S buffer and O accumulator in TMEM (2 x (128 + 128) = 512 columns) and its own softmax group, and
// the tensor pipe runs
//
//        S_A(n+1) | PV_B(n) | S_B(n+1) | PV_A(n+1) | S_A(n+2) | ...
//
// so the softmax of one tile (S -> row max -> P over the S columns) always sits under a PV + S of the other tile.
// UMMAs of one thread execute in issue order: S_X(n+1) is issued right behind the PV_X(n) that reads P_X(n) from the
// same columns, and a softmax thread that sees S_X(n+1) complete knows PV_X(n) is complete too (lazy O rescale
// without a further wait).
//
//   warps 0-3   softmax + epilogue of tile A   one thread per row (TMEM lane): 96 of the row's 128 scores stay in
//   warps 4-7   softmax + epilogue of tile B   registers between the max and the exp pass, no cross-thread exchange;
//                                              O / l -> 16-bit -> global rows directly
//   warp 8      TMA producer                   Q pair (single buffer), K ring (3 tiles), V ring (2 tiles)
//   warp 9      UMMA issuer
//
// The one-tile-per-CTA kernel of round 1 (fwd_sm100.cu, S -> softmax -> PV strictly in turn, two CTAs per SM) reached
// 31 % of the tensor peak at BASELINE configs[2].
#include "attn_common.cuh"
#include "tmap.cuh"

namespace sfa {
namespace {

template <int D_> struct F128Cfg {
  static constexpr int D = D_;                           // 128 (also 72 .. 120 by zero fill) or 64 (wide windows)
  static constexpr int kDS = D / 64;                     // 64-channel slabs
  static constexpr int kBNMax = 128;
  static constexpr int kKStages = (D == 64) ? 4 : 3, kVStages = (D == 64) ? 3 : 2;
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBNMax * D * 2;
  static constexpr int kSlabQ = 128 * 128;
  static constexpr int kSlabKV = kBNMax * 128;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS = 0;                   // S_A at 0, S_B at 128
  static constexpr uint32_t kColO = 256;                 // O_A at 256, O_B at 256 + D
  static constexpr int kThreads = 320;       // (registers are allocated per four warps: 10 warps get what 12 would)
  // no alignment slack: the dynamic shared memory is declared 1024-byte aligned (checked at kernel entry)
  static constexpr int kSmem = 2 * kQBytes + (kKStages + kVStages) * kKVBytes + 256;
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

struct F128Args {
  int B, N, S, W, Hq, G, P, BN, groups_per_kv, ny, npairs, total;
  unsigned long long bn_mul;
  int q_swap, k_swap, v_swap;
  int fmt;
  float sl2;
  const float* s_aux;
  float* lse;
  void* o;
  Strides4 so;
  int Dl;                // logical head_dim: channels Dl .. 127 are TMA zero fill on loads and are not stored
  const int* seq_lo;     // packed sequences (no sink tokens): first key of the row's sequence; nullptr: none
  int64_t seq_bs;
  int dbg_delay;         // test knob (sfa_set_debug 0): odd softmax warps sleep this many ns before their exp pass and
                         // tile B's group before its epilogue -- every result must stay bit-identical
};

// Super tile t -> (pair of position blocks, packed head group, batch), LAST positions first: the tiles with the longest
// bands start first and the short ones fill the tail of the launch; CTAs running at the same time work on neighbouring
// positions of the same heads (their K / V tiles meet in L2).
struct STile {
  int q0, y, b;
  TilePlan pl;
  __device__ __forceinline__ void set(const F128Args& a, int t) {
    const int per = a.ny * a.B;
    const int lvl = t / per, rem = t - lvl * per;
    y = rem % a.ny;
    b = rem / a.ny;
    q0 = (a.npairs - 1 - lvl) * 2 * a.P;
    pl = make_plan(q0, 2 * a.P, a.N, a.S, a.W, a.BN, a.bn_mul);
    if (a.seq_lo != nullptr) clamp_plan(pl, a.seq_lo[b * a.seq_bs + q0], a.W, a.BN, a.bn_mul);
  }
};

__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32p(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <typename T, int D_>
__global__ void __launch_bounds__(F128Cfg<D_>::kThreads, 1) fwd128_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV,
                                                                       const F128Args a) {
  using C = F128Cfg<D_>;
  constexpr int D = C::D;
  extern __shared__ __align__(1024) unsigned char smem_al[];
  unsigned char* smem = smem_al;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();               // SWIZZLE_128B tiles need 1024-byte aligned slabs
  unsigned char* q_s = smem;                                   // [2 tiles][kQBytes]
  unsigned char* k_s = q_s + 2 * C::kQBytes;                   // [kKStages][kKVBytes]
  unsigned char* v_s = k_s + C::kKStages * C::kKVBytes;        // [kVStages][kKVBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + C::kVStages * C::kKVBytes);
  uint64_t* q_full = bars;                       //      the Q pair of the super tile has landed
  uint64_t* q_empty = q_full + 1;                //      its last S UMMAs are complete
  uint64_t* k_full = q_empty + 1;                // [3]
  uint64_t* k_empty = k_full + C::kKStages;      // [3]  S_B(n) complete
  uint64_t* v_full = k_empty + C::kKStages;      // [2]
  uint64_t* v_empty = v_full + C::kVStages;      // [2]  PV_B(n) complete
  uint64_t* s_full = v_empty + C::kVStages;      // [2]  S_X(n) complete               (issuer -> softmax X)
  uint64_t* p_full = s_full + 2;                 // [2]  P_X(n) written over S_X(n)    (softmax X -> issuer)
  uint64_t* o_done = p_full + 2;                 // [2]  last PV_X of the tile done    (issuer -> softmax X)
  uint64_t* o_free = o_done + 2;                 // [2]  O_X read by the epilogue      (softmax X -> issuer)
  uint64_t* p_half = o_free + 2;                 // [2]  first 64 keys of P_X(n) written: PV of those keys starts under
                                                 //      the second half's exponentials   (softmax X -> issuer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < C::kKStages; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < C::kVStages; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(s_full + x, 1);
      mbar_init(p_full + x, 128);
      mbar_init(p_half + x, 128);
      mbar_init(o_done + x, 1);
      mbar_init(o_free + x, 128);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g = 0, tc = 0;
      STile st;
      for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++tc) {
        st.set(a, t);
        const int hq0 = st.y * a.G, kvh = st.y / a.groups_per_kv;
        mbar_wait(q_empty, (tc & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * C::kQBytes);
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int s = 0; s < C::kDS; ++s)
            tma_tile(q_s + x * C::kQBytes + s * C::kSlabQ, &tmQ, q_full, a.q_swap, s * 64, st.q0 + x * a.P, hq0, st.b);
        for (int n = 0; n < st.pl.n_tiles; ++n, ++g) {
          int kstart, cols; bool is_sink;
          st.pl.tile(n, a.BN, kstart, cols, is_sink);
          const int ks = g % C::kKStages, vs = g % C::kVStages;
          mbar_wait(k_empty + ks, ((g / C::kKStages) & 1) ^ 1);
          mbar_expect_tx(k_full + ks, a.BN * D * 2);
#pragma unroll
          for (int s = 0; s < C::kDS; ++s)
            tma_tile(k_s + ks * C::kKVBytes + s * C::kSlabKV, &tmK, k_full + ks, a.k_swap, s * 64, kstart, kvh, st.b);
          mbar_wait(v_empty + vs, ((g / C::kVStages) & 1) ^ 1);
          mbar_expect_tx(v_full + vs, a.BN * D * 2);
#pragma unroll
          for (int s = 0; s < C::kDS; ++s)
            tma_tile(v_s + vs * C::kKVBytes + s * C::kSlabKV, &tmV, v_full + vs, a.v_swap, s * 64, kstart, kvh, st.b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      const uint32_t idesc_pv = make_idesc(a.fmt, 128, D, 0, 1);
      const uint32_t qa0 = smem_u32(q_s), ka0 = smem_u32(k_s), va0 = smem_u32(v_s);
      auto issue_s = [&](int x, int g, int cols) {
        const uint32_t idesc_s = make_idesc(a.fmt, 128, cols, 0, 0);
        const uint32_t qa = qa0 + x * C::kQBytes, ka = ka0 + (g % C::kKStages) * C::kKVBytes;
#pragma unroll
        for (int s = 0; s < C::kDS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + C::kColS + x * 128, make_sdesc(qa + s * C::kSlabQ + kk * 32, 16, 1024),
                    make_sdesc(ka + s * C::kSlabKV + kk * 32, 16, 1024), idesc_s, (s | kk) != 0);
      };
      // key chunks [k0, k1) of PV_X(n): the first four (64 keys) go out as soon as that half of P is written
      auto issue_pv = [&](int x, int g, int cols, bool acc, int k0, int k1) {
        const uint32_t va = va0 + (g % C::kVStages) * C::kKVBytes;
        const int nk = cols >> 4;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (kk >= k0 && kk < k1 && kk < nk)
            umma_ts(tmem + C::kColO + x * D, tmem + C::kColS + x * 128 + kk * 8,
                    make_sdesc(va + kk * 2048, C::kSlabKV, 1024), idesc_pv, (acc || kk > 0));
      };
      int g = 0, tc = 0;
      STile st;
      for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++tc) {
        st.set(a, t);
        const int nt = st.pl.n_tiles;
        int kstart, cols, cols_n; bool is_sink;
        st.pl.tile(0, a.BN, kstart, cols, is_sink);
        mbar_wait(q_full, tc & 1);
        mbar_wait(k_full + g % C::kKStages, (g / C::kKStages) & 1);
        tc_fence_after();
        issue_s(0, g, cols);
        umma_commit(s_full + 0);
        issue_s(1, g, cols);
        umma_commit(s_full + 1);
        umma_commit(k_empty + g % C::kKStages);
        if (nt == 1) umma_commit(q_empty);
        for (int n = 0; n < nt; ++n, ++g) {
          const bool more = (n + 1 < nt);
          cols_n = cols;
          if (more) st.pl.tile(n + 1, a.BN, kstart, cols_n, is_sink);
          mbar_wait(v_full + g % C::kVStages, (g / C::kVStages) & 1);
          // ---- tile A: PV_A(n), S_A(n + 1)
          mbar_wait(p_half + 0, g & 1);
          if (n == 0 && tc > 0) mbar_wait(o_free + 0, (tc - 1) & 1);     // the epilogue has read the previous O_A
          tc_fence_after();
          issue_pv(0, g, cols, n > 0, 0, 4);
          mbar_wait(p_full + 0, g & 1);
          tc_fence_after();
          issue_pv(0, g, cols, n > 0, 4, 8);
          if (!more) umma_commit(o_done + 0);
          if (more) {
            mbar_wait(k_full + (g + 1) % C::kKStages, ((g + 1) / C::kKStages) & 1);
            tc_fence_after();
            issue_s(0, g + 1, cols_n);
            umma_commit(s_full + 0);
          }
          // ---- tile B: PV_B(n), S_B(n + 1)
          mbar_wait(p_half + 1, g & 1);
          if (n == 0 && tc > 0) mbar_wait(o_free + 1, (tc - 1) & 1);
          tc_fence_after();
          issue_pv(1, g, cols, n > 0, 0, 4);
          mbar_wait(p_full + 1, g & 1);
          tc_fence_after();
          issue_pv(1, g, cols, n > 0, 4, 8);
          umma_commit(v_empty + g % C::kVStages);
          if (!more) umma_commit(o_done + 1);
          if (more) {
            issue_s(1, g + 1, cols_n);
            umma_commit(s_full + 1);
            umma_commit(k_empty + (g + 1) % C::kKStages);
            if (n + 2 == nt) umma_commit(q_empty);       // the tile's last S UMMAs: Q may be replaced
          }
          cols = cols_n;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue of tile x
    const int x = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // MMA row == TMEM lane
    const int pr = a.q_swap ? (r / a.G) : (r % a.P);   // position within the tile
    const int gr = a.q_swap ? (r % a.G) : (r / a.P);   // head within the packed group
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t ts = tl + C::kColS + x * 128, to = tl + C::kColO + x * D;
    const uint64_t sl2_2 = pack_f32x2(a.sl2, a.sl2);
    int g = 0, tc = 0;
    STile st;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++tc) {
      st.set(a, t);
      const int i = st.q0 + x * a.P + pr;
      const int h = st.y * a.G + gr;
      int row_lo = 0;
      if (a.seq_lo != nullptr && i < a.N) row_lo = __ldg(a.seq_lo + st.b * a.seq_bs + i);
      float m_used = a.s_aux ? __ldg(a.s_aux + h) * kLog2e : -INFINITY;
      float l = a.s_aux ? 1.f : 0.f;
      for (int n = 0; n < st.pl.n_tiles; ++n, ++g) {
        int kstart, cols; bool is_sink;
        st.pl.tile(n, a.BN, kstart, cols, is_sink);
        int c_lo, c_hi;   // attended columns of this row inside the tile: [c_lo, c_hi]
        row_range(is_sink, i, kstart, cols, a.S, a.W, c_lo, c_hi);
        if (a.seq_lo != nullptr) c_lo = max(c_lo, row_lo - kstart);
        if (i >= a.N) c_hi = -1;
        mbar_wait(s_full + x, g & 1);
        tc_fence_after();
        if (a.dbg_delay && (warp & 1)) __nanosleep(a.dbg_delay);
        // Row sums in both paths: pair k (columns 2k, 2k + 1) goes to accumulator k & 3, masked elements add exactly
        // 0 -- a row gets the same bits whichever path its warp takes (the layout-invariance tests rely on it).
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        bool half_done = false;
        if (__all_sync(0xffffffffu, c_lo <= 0 && c_hi >= 127)) {
          // ---- interior tile: 96 of the row's 128 scores stay in registers between the max and the exp pass, the
          // last 32 are read twice (all 128 + the packed P exceed the 168 registers a thread of 12 warps can have)
          uint32_t s[96], u[32];
          tmem_ld32p(ts + 96, u);
          tmem_ld32p(ts, s);
          tmem_ld_wait();
          tmem_ld32p(ts + 32, s + 32);
          tmem_ld32p(ts + 64, s + 64);
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(u[e]));
#pragma unroll
          for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(s[e]));
          tmem_ld_wait();
#pragma unroll
          for (int e = 32; e < 96; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(s[e]));
          const float mxa = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          const float m_new = fmaxf(m_used, mxa * a.sl2);
          const bool need = (m_new - m_used) > 8.0f;        // also -inf -> finite; false for NaN (-inf - -inf)
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? exp2f(m_used - m_new) : 1.f;
            if (need) {
              l *= alpha;
              m_used = m_new;
            }
            if (n > 0) {      // PV_X(n - 1) was issued before S_X(n): complete by now
#pragma unroll 1
              for (int c0 = 0; c0 < D; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(to + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
                tmem_st16(to + c0, v);
              }
            }
          }
          const float neg_m = -m_used;
          const uint64_t negm_2 = pack_f32x2(neg_m, neg_m);
          uint32_t pk[64];
          auto exp_pair = [&](uint32_t s0, uint32_t s1, int k) {
            float x0, x1;
            unpack_f32x2(fma_f32x2(s0, s1, sl2_2, negm_2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            acc[k & 3] += p0 + p1;
            pk[k] = pack16_fast<T>(p0, p1);
          };
#pragma unroll
          for (int k = 0; k < 16; ++k) exp_pair(s[2 * k], s[2 * k + 1], k);
          tmem_ld32p(ts + 96, u);                    // second read of the last 32 scores, under the next 64 exponentials
#pragma unroll
          for (int k = 16; k < 32; ++k) exp_pair(s[2 * k], s[2 * k + 1], k);
          tmem_st32p(ts, pk);
          tmem_st_wait();                     // (the reload of the last 32 scores stays in flight: wait::st, not wait::ld)
          tc_fence_before();
          mbar_arrive(p_half + x);
          half_done = true;
#pragma unroll
          for (int k = 32; k < 48; ++k) exp_pair(s[2 * k], s[2 * k + 1], k);
          tmem_ld_wait();
#pragma unroll
          for (int k = 48; k < 64; ++k) exp_pair(u[2 * k - 96], u[2 * k - 95], k);
          tmem_st32p(ts + 32, pk + 32);
        } else {
          // ---- boundary tile (band start, diagonal, sink tokens, short last tile): 16-column chunks with the mask
          float mx = -INFINITY;
#pragma unroll 1
          for (int c0 = 0; c0 < cols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(ts + c0, v);
            tmem_ld_wait();
            const int lo = c_lo - c0, hi = c_hi - c0;
#pragma unroll
            for (int e = 0; e < 16; ++e) mx = (e >= lo && e <= hi) ? fmaxf(mx, __uint_as_float(v[e])) : mx;
          }
          const float m_new = fmaxf(m_used, mx * a.sl2);
          const bool need = (m_new - m_used) > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? exp2f(m_used - m_new) : 1.f;
            if (need) {
              l *= alpha;
              m_used = m_new;
            }
            if (n > 0) {
#pragma unroll 1
              for (int c0 = 0; c0 < D; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(to + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
                tmem_st16(to + c0, v);
              }
            }
          }
          const float neg_m = -m_used;
#pragma unroll 1
          for (int c0 = 0; c0 < cols; c0 += 16) {
            uint32_t v[16], pk[8];
            tmem_ld16(ts + c0, v);
            tmem_ld_wait();
            const int lo = c_lo - c0, hi = c_hi - c0;
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0 = fast_exp2(fmaf(__uint_as_float(v[e]), a.sl2, neg_m));
              float p1 = fast_exp2(fmaf(__uint_as_float(v[e + 1]), a.sl2, neg_m));
              p0 = (e >= lo && e <= hi) ? p0 : 0.f;
              p1 = (e + 1 >= lo && e + 1 <= hi) ? p1 : 0.f;
              acc[(e >> 1) & 3] += p0 + p1;
              pk[e >> 1] = pack16_fast<T>(p0, p1);
            }
            tmem_st8(ts + (c0 >> 1), pk);
          }
        }
        l += (acc[0] + acc[1]) + (acc[2] + acc[3]);
        tmem_st_wait();
        tc_fence_before();
        if (!half_done) mbar_arrive(p_half + x);      // boundary tiles hand both halves over at once
        mbar_arrive(p_full + x);
      }
      // ---------------- epilogue: O / l -> 16-bit -> this row of the output (256 contiguous bytes), LSE
      mbar_wait(o_done + x, tc & 1);
      tc_fence_after();
      if (a.dbg_delay && x == 1) __nanosleep(a.dbg_delay * 4);
      const float inv = (l > 0.f) ? 1.f / l : 0.f;
      T* orow = static_cast<T*>(a.o) + st.b * a.so.b + static_cast<int64_t>(h) * a.so.h + static_cast<int64_t>(i) * a.so.n;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32], pk[16];
        tmem_ld32p(to + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2)
          pk[e >> 1] = pack16<T>(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv);
        if (i < a.N) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (c0 + q * 8 < a.Dl)
              *reinterpret_cast<uint4*>(orow + c0 + q * 8) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(o_free + x);
      if (i < a.N)
        a.lse[(static_cast<int64_t>(st.b) * a.Hq + h) * a.N + i] = (l > 0.f) ? m_used * kLn2 + logf(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, C::kTmemCols);
}

template <typename T, int D_>
cudaError_t launch_fwd128(const AttnParams& p, int dtype, cudaStream_t st) {
  using C = F128Cfg<D_>;
  static std::atomic<unsigned long long> attr_done{0};
  if (cudaError_t e = ensure_dyn_smem(fwd128_kernel<T, D_>, C::kSmem, attr_done)) return e;
  const int group = p.Hq / p.Hkv;
  int G, P;
  pick_packing(p.Hq, p.Hkv, G, P);
  const int BN = pick_bn(p.W, p.N, 2 * P, C::kBNMax);
  TileMap mq, mk, mv;
  if (!make_tile_map(&mq, p.q, dtype, p.D, p.N, p.Hq, p.B, p.sq, P, G)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mk, p.k, dtype, p.D, p.N, p.Hkv, p.B, p.sk, BN, 1)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mv, p.v, dtype, p.D, p.N, p.Hkv, p.B, p.sv, BN, 1)) return cudaErrorInvalidValue;
  F128Args a;
  a.B = p.B; a.N = p.N; a.S = p.S; a.W = p.W; a.Hq = p.Hq; a.G = G; a.P = P; a.BN = BN;
  a.groups_per_kv = group / G;
  a.ny = p.Hq / G;
  a.npairs = ((p.N + P - 1) / P + 1) / 2;
  a.total = a.npairs * a.ny * p.B;
  a.bn_mul = bn_magic(BN);
  a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh;
  a.fmt = (dtype == SFA_DTYPE_BF16) ? 1 : 0;
  a.sl2 = p.scale * kLog2e;
  a.s_aux = p.s_aux;
  a.lse = p.lse;
  a.o = p.o; a.so = p.so; a.Dl = p.D;
  a.seq_lo = p.seq_lo; a.seq_bs = p.seq_bs;
  a.dbg_delay = debug_knob(0);
  const int sms = device_sm_count();
  const int grid = a.total < sms ? a.total : sms;
  fwd128_kernel<T, D_><<<grid, C::kThreads, C::kSmem, st>>>(mq.map, mk.map, mv.map, a);
  return cudaGetLastError();
}

}  // namespace

// O rows are written with 16-byte stores straight from the registers (no staging tile: the shared memory holds the Q
// pair and the K / V rings)
bool tc_fwd128_supported(const AttnParams& p, int dtype) {
  if (!tc_fwd_supported(p, dtype)) return false;      // head_dim 64 or 72 .. 128; the caller picks 64 for wide windows only
  if (p.q_off != 0 || p.Nkv != p.N || (p.seq_lo != nullptr && p.S > 0) || p.o_route != nullptr) return false;
  if (p.S <= 0 && p.W <= 0) return false;      // nothing attended: the one-tile-per-CTA kernel writes the O = 0 rows
  if (reinterpret_cast<uintptr_t>(p.o) % 16 || p.so.n % 8 || p.so.h % 8 || p.so.b % 8) return false;
  return true;
}

cudaError_t tc_fwd128(const AttnParams& p, int dtype, cudaStream_t st) {
  if (p.D == 64)
    return dtype == SFA_DTYPE_BF16 ? launch_fwd128<__nv_bfloat16, 64>(p, dtype, st) : launch_fwd128<__half, 64>(p, dtype, st);
  return dtype == SFA_DTYPE_BF16 ? launch_fwd128<__nv_bfloat16, 128>(p, dtype, st) : launch_fwd128<__half, 128>(p, dtype, st);
}

}  // namespace sfa
