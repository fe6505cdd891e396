// Fused backward for narrow sliding windows (head_dim 64, no sink tokens): dQ, dK and dV from ONE pass over
// the packed query tiles.  Replaces, for that case, the pair dq64_kernel + dkdv64_kernel (reference kernels
// _sink_flash_attn_bwd_dq_kernel sink_flash_attention.py:371-484 and _sink_flash_attn_bwd_dkdv_kernel :256-364
// plus the torch GQA group sum :648-651), which recomputed S / P / dP / dS twice and -- in the key-stationary
// dK/dV kernel -- spent half of its math on the masked corners of 128 x 128 tiles.
//
// Tile = 128 MMA rows (G q-heads of one KV head x P consecutive positions, P = 128 / G) against the nb * P keys
// its window can reach: key blocks pb-nb+1 .. pb of P keys each (block start may be negative: TMA zero-fills).
//
//   S  = Q K^T, dP = dO V^T          M = 128, N = nb*P          SS UMMAs, fp32 in TMEM
//   P  = exp2(S*c - lse), dS = P o (dP - delta)                 math warps: TMEM -> registers -> 16-bit in SHARED
//                                                               memory as [col/8][row/8][row%8][col%8] (un-swizzled
//                                                               core matrices) -- one image that is both a K-major A
//                                                               operand and an MN-major B operand
//   dQ   = dS K                      M = 128, N = 64            A = dS (smem), B = K tile (MN-major)
//   dV^T += dO^T P                   M = 64,  N = ring          A = dO tile (MN-major), B = P (smem)
//   dK^T += Q^T dS                   M = 64,  N = ring          A = Q tile (MN-major),  B = dS (smem)
//
// dK^T / dV^T live in a RING of R = nb + 1 key-block slots of P TMEM columns (block j -> slot j mod R); an M = 64
// accumulator only occupies lanes 0-15 of every lane quarter, so dV^T sits at lane offset 0 and dK^T at lane
// offset 16 of the SAME columns (validated by sfa_probe_umma mode 4).  P and dS are stored in ring-column
// order, so each k-step of dV^T / dK^T is at most two UMMAs (the ring minus the one slot that is being
// drained).  After the tile that last touches a block, the epilogue warps read its slot (dV^T and dK^T with
// one tcgen05.ld), zero it and write dK/dV -- the GQA group sum happened inside the contraction over the rows.
//
// A CTA owns a contiguous run of tiles.  Key blocks shared with the neighbouring CTA (the nb - 1 blocks before
// its first tile and the last nb - 1 blocks of its run) are written as fp32 partials and summed by a small
// fix-up kernel: no atomics, no inter-CTA waits, deterministic.
//
// Warp roles (19 warps): 0-11 math (lane quarter = warp & 3, a third of the 16-column chunks each), 12-15
// epilogue (dQ store, ring drain), 16 TMA producer, 17 UMMA issuer S / dP, 18 UMMA issuer dV^T / dK^T / dQ.
#include <stdlib.h>

#include "attn_common.cuh"
#include "tmap.cuh"

namespace sfa {
namespace {

struct FusedCfg {
  static constexpr int D = 64;
  static constexpr int kColsMax = 144;                 // keys per tile (UMMA N of S and dP)
  static constexpr int kRingCols = 160;                // (nb + 1) * P
  static constexpr int kQBytes = 128 * D * 2;          // Q / dO tile
  static constexpr int kKVBytes = kColsMax * D * 2;    // K / V tile
  static constexpr int kPBytes = (kRingCols / 8) * 2048;   // P or dS image: [col/8][row/8][row%8][8 x 16-bit]
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS = 0;
  static constexpr uint32_t kColP = kColsMax;          // dP
  static constexpr uint32_t kColQ = 2 * kColsMax;      // dQ
  static constexpr uint32_t kColR = 2 * kColsMax + D;  // ring
  static constexpr int kMathWarps = 12;
  static constexpr int kThreads = 19 * 32;
  static constexpr int kSmem = 1024 + 4 * kQBytes + 4 * kKVBytes + 2 * kPBytes + 512;
  static constexpr int kPartKeys = 128;                // keys per side of a CTA's fp32 partials
  static constexpr int kMaxCtas = 160;
  static_assert(kColR + kRingCols <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

struct FusedArgs {
  int B, N, W, Hq, Hkv, G, P, lgP, nb, R, cols, nch, nblk, total_tiles, tiles_per_cta;
  int q_swap, k_swap, v_swap;
  int fmt;       // 0 f16, 1 bf16
  float sl2;     // scale * log2(e)
  float scale;
  const float* lse;
  const float* delta;
  void* dq;
  void* dk;
  void* dv;
  Strides4 sdq, sdk, sdv;
  float* part;   // [grid][2 sides][kPartKeys][2 (dV, dK)][64] fp32
};

template <typename T> __device__ __forceinline__ void unpack16f(uint32_t u, float& a, float& b);
template <> __device__ __forceinline__ void unpack16f<__nv_bfloat16>(uint32_t u, float& a, float& b) {
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack16f<__half>(uint32_t u, float& a, float& b) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
  a = f.x;
  b = f.y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// This CTA's contiguous run of tiles; tile id = (b * Hkv + y) * nblk + pb.
struct FusedWalk {
  const FusedArgs& a;
  int tile, end, pb, y, b, it;
  bool seg_first;     // first tile of a sequence segment inside this CTA
  __device__ __forceinline__ explicit FusedWalk(const FusedArgs& a_) : a(a_), it(-1), seg_first(false) {
    tile = static_cast<int>(blockIdx.x) * a.tiles_per_cta;
    end = min(tile + a.tiles_per_cta, a.total_tiles);
    pb = tile % a.nblk;
    const int r = tile / a.nblk;
    y = r % a.Hkv;
    b = r / a.Hkv;
    --tile;
    --pb;
  }
  __device__ __forceinline__ bool next() {
    ++tile;
    ++it;
    if (tile >= end) return false;
    ++pb;
    seg_first = (it == 0);
    if (pb == a.nblk) {
      pb = 0;
      seg_first = true;
      if (++y == a.Hkv) {
        y = 0;
        ++b;
      }
    }
    return true;
  }
  __device__ __forceinline__ bool seq_end() const { return pb == a.nblk - 1; }
  __device__ __forceinline__ bool seg_last() const { return pb == a.nblk - 1 || tile == end - 1; }
};

// ring slot of the tile's first key block (block pb - nb + 1 == pb + 2 mod R), tracked without divisions
struct SlotTrack {
  int slot0;
  __device__ __forceinline__ void step(const FusedWalk& w, int R) {
    if (w.seg_first) slot0 = (w.pb + 2) % R;
    else if (++slot0 == R) slot0 = 0;
  }
};

template <typename T>
__global__ void __launch_bounds__(FusedCfg::kThreads, 1) bwd_fused64_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                             const __grid_constant__ CUtensorMap tmdO,
                                                                             const __grid_constant__ CUtensorMap tmK,
                                                                             const __grid_constant__ CUtensorMap tmV,
                                                                             const FusedArgs a) {
  using C = FusedCfg;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = smem;                          // [2][kQBytes]
  unsigned char* do_s = q_s + 2 * C::kQBytes;         // [2][kQBytes]
  unsigned char* k_s = do_s + 2 * C::kQBytes;         // [2][kKVBytes]
  unsigned char* v_s = k_s + 2 * C::kKVBytes;         // [2][kKVBytes]
  unsigned char* p_s = v_s + 2 * C::kKVBytes;         // P image
  unsigned char* ds_s = p_s + C::kPBytes;             // dS image
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_s + C::kPBytes);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = q_full + 2;     // [2]  dK^T(n) complete
  uint64_t* k_full = q_empty + 2;
  uint64_t* k_empty = k_full + 2;     //      dQ(n) complete
  uint64_t* do_full = k_empty + 2;
  uint64_t* do_empty = do_full + 2;   //      dV^T(n) complete
  uint64_t* v_full = do_empty + 2;
  uint64_t* v_empty = v_full + 2;     //      dP(n) complete
  uint64_t* s_full = v_empty + 2;     // S(n) complete                          (issuer A -> math)
  uint64_t* s_free = s_full + 1;      // S(n) read                              (math -> issuer A)
  uint64_t* dp_full = s_free + 1;     // dP(n) complete                         (issuer A -> math)
  uint64_t* dp_free = dp_full + 1;    // dP(n) read                             (math -> issuer A)
  uint64_t* p_ready = dp_free + 1;    // P(n) in shared memory                  (math -> issuer B)
  uint64_t* p_free = p_ready + 1;     // dV^T(n) complete: P image reusable     (issuer B -> math)
  uint64_t* ds_ready = p_free + 1;    // dS(n) in shared memory                 (math -> issuer B)
  uint64_t* ds_free = ds_ready + 1;   // dK^T(n), dQ(n) complete                (issuer B -> math)
  uint64_t* dq_done = ds_free + 1;    // all UMMAs of tile n complete           (issuer B -> epilogue)
  uint64_t* dq_free = dq_done + 1;    // dQ(n) read                             (epilogue -> issuer B)
  uint64_t* drain_done = dq_free + 1; // [2] ring drains of tile n finished     (epilogue -> issuer B), by n & 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drain_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < 16; ++s) mbar_init(bars + s, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, C::kMathWarps);
    mbar_init(dp_full, 1);
    mbar_init(dp_free, C::kMathWarps);
    mbar_init(p_ready, C::kMathWarps);
    mbar_init(p_free, 1);
    mbar_init(ds_ready, C::kMathWarps);
    mbar_init(ds_free, 1);
    mbar_init(dq_done, 1);
    mbar_init(dq_free, 4);
    mbar_init(drain_done + 0, 4);
    mbar_init(drain_done + 1, 4);
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp >= 12 && warp < 16) {      // the ring accumulates from the first tile on: start from zero
    const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t z[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) z[e] = 0u;
    for (int c = 0; c < C::kRingCols; c += 16) tmem_st16(tl + C::kColR + c, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int P = a.P, nb = a.nb, R = a.R;

  if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      FusedWalk w(a);
      const uint32_t kv_bytes = a.cols * C::D * 2;
      while (w.next()) {
        const int s = w.it & 1;
        const uint32_t eph = ((w.it >> 1) & 1) ^ 1;
        const int q0 = w.pb * P, hq0 = w.y * a.G, kstart = (w.pb - nb + 1) * P;
        mbar_wait(q_empty + s, eph);
        mbar_expect_tx(q_full + s, C::kQBytes);
        tma_tile(q_s + s * C::kQBytes, &tmQ, q_full + s, a.q_swap, 0, q0, hq0, w.b);
        mbar_wait(k_empty + s, eph);
        mbar_expect_tx(k_full + s, kv_bytes);
        tma_tile(k_s + s * C::kKVBytes, &tmK, k_full + s, a.k_swap, 0, kstart, w.y, w.b);
        mbar_wait(do_empty + s, eph);
        mbar_expect_tx(do_full + s, C::kQBytes);
        tma_tile(do_s + s * C::kQBytes, &tmdO, do_full + s, a.q_swap, 0, q0, hq0, w.b);
        mbar_wait(v_empty + s, eph);
        mbar_expect_tx(v_full + s, kv_bytes);
        tma_tile(v_s + s * C::kKVBytes, &tmV, v_full + s, a.v_swap, 0, kstart, w.y, w.b);
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ------------------------------------------------------------------ UMMA issuer A: S = Q K^T, dP = dO V^T
    if (lane == 0) {
      const uint32_t idesc = make_idesc(a.fmt, 128, a.cols, 0, 0);
      FusedWalk w(a);
      while (w.next()) {
        const int s = w.it & 1;
        const uint32_t fph = (w.it >> 1) & 1;
        const uint64_t qd = make_sdesc(smem_u32(q_s + s * C::kQBytes), 16, 1024);
        const uint64_t kd = make_sdesc(smem_u32(k_s + s * C::kKVBytes), 16, 1024);
        const uint64_t dod = make_sdesc(smem_u32(do_s + s * C::kQBytes), 16, 1024);
        const uint64_t vd = make_sdesc(smem_u32(v_s + s * C::kKVBytes), 16, 1024);
        mbar_wait(q_full + s, fph);
        mbar_wait(k_full + s, fph);
        if (w.it >= 1) mbar_wait(s_free, (w.it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C::kColS, qd + kk * 2, kd + kk * 2, idesc, kk != 0);
        umma_commit(s_full);
        mbar_wait(do_full + s, fph);
        mbar_wait(v_full + s, fph);
        if (w.it >= 1) mbar_wait(dp_free, (w.it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C::kColP, dod + kk * 2, vd + kk * 2, idesc, kk != 0);
        umma_commit(dp_full);
        umma_commit(v_empty + s);
      }
    }
    __syncwarp();
  } else if (warp == 18) {
    // ------------------------------------------------------------------ UMMA issuer B: dV^T, dK^T, dQ
    if (lane == 0) {
      const uint32_t idesc_dq = make_idesc(a.fmt, 128, C::D, 0, 1);
      const uint32_t p_a = smem_u32(p_s), ds_a = smem_u32(ds_s);
      const int sh = a.lgP - 4;                 // 16-column chunks per key block = 1 << sh
      FusedWalk w(a);
      SlotTrack st;
      st.slot0 = 0;
      while (w.next()) {
        st.step(w, R);
        const int s = w.it & 1;
        const uint32_t fph = (w.it >> 1) & 1;
        // active ring columns: everything but the slot of the block that left the window with the previous tile
        int x = st.slot0 + nb;
        if (x >= R) x -= R;
        const int n_lo = x * P, c_hi0 = (x + 1) * P, n_hi = (R - 1 - x) * P;
        const uint32_t idesc_lo = make_idesc(a.fmt, 64, n_lo, 1, 1), idesc_hi = make_idesc(a.fmt, 64, n_hi, 1, 1);
        const uint32_t do_a = smem_u32(do_s + s * C::kQBytes), q_a = smem_u32(q_s + s * C::kQBytes);
        const uint32_t k_a = smem_u32(k_s + s * C::kKVBytes);

        mbar_wait(p_ready, w.it & 1);
        mbar_wait(do_full + s, fph);
        // the slot the tile's newest block enters must have been drained (and zeroed)
        if (w.seg_first) {
          if (w.it >= 1) mbar_wait(drain_done + ((w.it - 1) & 1), ((w.it - 1) >> 1) & 1);
        } else if (w.it >= 2) {
          mbar_wait(drain_done + (w.it & 1), ((w.it - 2) >> 1) & 1);
        }
        tc_fence_after();
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_sdesc(do_a + kk * 2048, 16384, 1024);
          if (n_lo > 0) umma_ss(tmem + C::kColR, ad, make_sdesc_ns(p_a + kk * 256, 128, 2048), idesc_lo, 1);
          if (n_hi > 0)
            umma_ss(tmem + C::kColR + c_hi0, ad, make_sdesc_ns(p_a + (c_hi0 >> 3) * 2048 + kk * 256, 128, 2048), idesc_hi, 1);
        }
        umma_commit(p_free);
        umma_commit(do_empty + s);

        mbar_wait(ds_ready, w.it & 1);
        mbar_wait(q_full + s, fph);
        tc_fence_after();
        const uint32_t ring_k = tmem + C::kColR + (16u << 16);
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_sdesc(q_a + kk * 2048, 16384, 1024);
          if (n_lo > 0) umma_ss(ring_k, ad, make_sdesc_ns(ds_a + kk * 256, 128, 2048), idesc_lo, 1);
          if (n_hi > 0)
            umma_ss(ring_k + c_hi0, ad, make_sdesc_ns(ds_a + (c_hi0 >> 3) * 2048 + kk * 256, 128, 2048), idesc_hi, 1);
        }
        umma_commit(q_empty + s);

        mbar_wait(k_full + s, fph);
        if (w.it >= 1) mbar_wait(dq_free, (w.it - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int kk = 0; kk < a.nch; ++kk) {
          int slot = st.slot0 + (kk >> sh);
          if (slot >= R) slot -= R;
          const int rc = slot * P + ((kk & ((1 << sh) - 1)) << 4);
          umma_ss(tmem + C::kColQ, make_sdesc_ns(ds_a + (rc >> 3) * 2048, 2048, 128),
                  make_sdesc(k_a + kk * 2048, C::kKVBytes, 1024), idesc_dq, kk > 0);
        }
        umma_commit(ds_free);
        umma_commit(k_empty + s);
        umma_commit(dq_done);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                     // TMEM lane == MMA row
    const int pr = a.q_swap ? (r / a.G) : (r & (P - 1));
    const int gr = a.q_swap ? (r & (a.G - 1)) : (r >> a.lgP);
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);

    if (warp < C::kMathWarps) {
      // ---------------------------------------------------------------- math warps
      const int part = warp >> 2;                          // chunks part, part + 3, part + 6
      const int sh = a.lgP - 4;
      const uint32_t rowoff = static_cast<uint32_t>((r >> 3) * 128 + (r & 7) * 16);
      const uint32_t p_a = smem_u32(p_s) + rowoff, ds_a = smem_u32(ds_s) + rowoff;
      auto load_row = [&](const float* src, const FusedWalk& t, bool valid, float dflt) {
        float v = dflt;
        const int i = t.pb * P + pr;
        if (valid && i < a.N) {
          const int64_t row = (static_cast<int64_t>(t.b) * a.Hq + t.y * a.G + gr) * a.N + i;
          asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(src + row));
        }
        return v;
      };
      FusedWalk w(a), wn(a);
      SlotTrack st;
      st.slot0 = 0;
      bool has_next = wn.next();
      float l_next = load_row(a.lse, wn, has_next, INFINITY), d_next = load_row(a.delta, wn, has_next, 0.f);
      while (w.next()) {
        st.step(w, R);
        const float lse_i = l_next, delta = d_next;
        has_next = wn.next();
        l_next = load_row(a.lse, wn, has_next, INFINITY);
        d_next = load_row(a.delta, wn, has_next, 0.f);
        const float neg_l2 = (lse_i == -INFINITY) ? -INFINITY : -lse_i * kLog2e;   // lse = +-inf: P = 0
        const int i = w.pb * P + pr;
        const int kstart = (w.pb - nb + 1) * P;
        const int c_lo = max(i - a.W + 1, 0) - kstart;
        const int c_hi = (i < a.N) ? (i - kstart) : -1;

        uint32_t pk[3][8];
        uint32_t off[3];
        // ---- pass 1: P = exp2(S * c - lse), masked, 16-bit -> P image (ring-column order)
        mbar_wait_warp(s_full, w.it & 1);
        if (w.it >= 1) mbar_wait_warp(p_free, (w.it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int cb = part + 3 * k;
          if (cb < a.nch) {
            uint32_t sv[16];
            tmem_ld16(tl + C::kColS + cb * 16, sv);
            tmem_ld_wait();
            int slot = st.slot0 + (cb >> sh);
            if (slot >= R) slot -= R;
            const int rc = slot * P + ((cb & ((1 << sh) - 1)) << 4);
            off[k] = static_cast<uint32_t>(rc >> 3) * 2048u;
            const int lo = c_lo - cb * 16, hi = c_hi - cb * 16;      // attended elements of this chunk: [lo, hi]
            if (__all_sync(0xffffffffu, lo <= 0 && hi >= 15)) {
#pragma unroll
              for (int e = 0; e < 16; e += 2)
                pk[k][e >> 1] = pack16<T>(fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, neg_l2)),
                                          fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, neg_l2)));
            } else {
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, neg_l2));
                float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, neg_l2));
                p0 = (e >= lo && e <= hi) ? p0 : 0.f;
                p1 = (e + 1 >= lo && e + 1 <= hi) ? p1 : 0.f;
                pk[k][e >> 1] = pack16<T>(p0, p1);
              }
            }
            st_shared_v4(p_a + off[k], pk[k][0], pk[k][1], pk[k][2], pk[k][3]);
            st_shared_v4(p_a + off[k] + 2048u, pk[k][4], pk[k][5], pk[k][6], pk[k][7]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(s_free);
          mbar_arrive(p_ready);
        }
        // ---- pass 2: dS = P o (dP - delta), 16-bit -> dS image
        mbar_wait_warp(dp_full, w.it & 1);
        if (w.it >= 1) mbar_wait_warp(ds_free, (w.it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int cb = part + 3 * k;
          if (cb < a.nch) {
            uint32_t dv[16], dk[8];
            tmem_ld16(tl + C::kColP + cb * 16, dv);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0, p1;
              unpack16f<T>(pk[k][e >> 1], p0, p1);          // masked P is exactly 0 and dP is finite: dS = 0 there
              dk[e >> 1] = pack16<T>(p0 * (__uint_as_float(dv[e]) - delta), p1 * (__uint_as_float(dv[e + 1]) - delta));
            }
            st_shared_v4(ds_a + off[k], dk[0], dk[1], dk[2], dk[3]);
            st_shared_v4(ds_a + off[k] + 2048u, dk[4], dk[5], dk[6], dk[7]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dp_free);
          mbar_arrive(ds_ready);
        }
      }
    } else {
      // ---------------------------------------------------------------- epilogue warps: dQ store, ring drain
      const int which = lane >> 4;                          // 0: dV^T (lanes 0-15), 1: dK^T (lanes 16-31)
      const int dch = quarter * 16 + (lane & 15);           // channel of this lane's ring row
      const float osc = which ? a.scale : 1.f;
      T* const okv = static_cast<T*>(which ? a.dk : a.dv);
      const Strides4 skv = which ? a.sdk : a.sdv;
      float* const part_cta = a.part + static_cast<size_t>(blockIdx.x) * 2 * C::kPartKeys * 128 + which * 64 + dch;
      FusedWalk w(a);
      SlotTrack st;
      st.slot0 = 0;
      int pa = 0;
      while (w.next()) {
        st.step(w, R);
        if (w.seg_first) pa = w.pb;
        mbar_wait_warp(dq_done, w.it & 1);
        tc_fence_after();
        {
          uint32_t v[4][16];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld16(tl + C::kColQ + cc * 16, v[cc]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_free);
          const int i = w.pb * P + pr;
          if (i < a.N) {
            T* dst = static_cast<T*>(a.dq) + static_cast<int64_t>(w.b) * a.sdq.b +
                     static_cast<int64_t>(w.y * a.G + gr) * a.sdq.h + static_cast<int64_t>(i) * a.sdq.n;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint32_t pk[8];
#pragma unroll
              for (int e2 = 0; e2 < 16; e2 += 2)
                pk[e2 >> 1] = pack16<T>(__uint_as_float(v[cc][e2]) * a.scale, __uint_as_float(v[cc][e2 + 1]) * a.scale);
              *reinterpret_cast<uint4*>(dst + cc * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(dst + cc * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        // ring drain: the block this tile touched last -- or every live block at the end of a segment
        const bool last = w.seg_last();
        const bool tails = last && !w.seq_end();            // the run ends inside a sequence: blocks jb >= 1 are partial
        const int nd = last ? nb : 1;
        const int nh = P >> 4;
#pragma unroll 1
        for (int jb = 0; jb < nd; ++jb) {
          const int j = w.pb - nb + 1 + jb;
          if (j < 0) continue;
          int slot = st.slot0 + jb;
          if (slot >= R) slot -= R;
#pragma unroll 1
          for (int h = 0; h < nh; ++h) {
            uint32_t x[16], z[16];
            const uint32_t col = tl + C::kColR + slot * P + h * 16;
            tmem_ld16(col, x);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) z[e] = 0u;
            tmem_st16(col, z);
            const bool head = j < pa, tail = tails && jb >= 1;
            if (head || tail) {
              const int idx = head ? (j - (pa - nb + 1)) : (jb - 1);
              float* dst = part_cta + (static_cast<size_t>(tail ? 1 : 0) * C::kPartKeys + idx * P + h * 16) * 128;
#pragma unroll
              for (int e = 0; e < 16; ++e) dst[e * 128] = __uint_as_float(x[e]) * osc;
            } else {
              const int key0 = j * P + h * 16;
              T* dst = okv + static_cast<int64_t>(w.b) * skv.b + static_cast<int64_t>(w.y) * skv.h +
                       static_cast<int64_t>(key0) * skv.n + dch;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (key0 + e < a.N) dst[static_cast<int64_t>(e) * skv.n] = from_f<T>(__uint_as_float(x[e]) * osc);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(drain_done + (w.it & 1));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem, C::kTmemCols);
}

// Sums the fp32 partials of the key blocks shared by two neighbouring CTAs (tail of c - 1, head of c).
template <typename T>
__global__ void __launch_bounds__(256) bwd_fused_fixup_kernel(const FusedArgs a) {
  using C = FusedCfg;
  const int c = blockIdx.x + 1;
  const int t0 = c * a.tiles_per_cta;
  const int pa = t0 % a.nblk;
  if (pa == 0) return;                       // the boundary coincides with a sequence start: nothing shared
  const int seq = t0 / a.nblk, y = seq % a.Hkv, b = seq / a.Hkv;
  const float* tail = a.part + (static_cast<size_t>(c - 1) * 2 + 1) * C::kPartKeys * 128;
  const float* head = a.part + (static_cast<size_t>(c) * 2 + 0) * C::kPartKeys * 128;
  const int n = (a.nb - 1) * a.P * 128;
  const int key_base = (pa - a.nb + 1) * a.P;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int key = key_base + (e >> 7);
    if (key < 0 || key >= a.N) continue;
    const int which = (e >> 6) & 1, d = e & 63;
    const float v = tail[e] + head[e];
    T* o = static_cast<T*>(which ? a.dk : a.dv);
    const Strides4& s = which ? a.sdk : a.sdv;
    o[static_cast<int64_t>(b) * s.b + static_cast<int64_t>(y) * s.h + static_cast<int64_t>(key) * s.n + d] = from_f<T>(v);
  }
}

int fused_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// geometry of the fused path; false when the problem does not fit it
bool fused_geometry(const AttnParams& p, int& G, int& P, int& nb) {
  if (p.D != 64 || p.S != 0 || p.W < 1 || p.N < 1) return false;
  pick_packing(p.Hq, p.Hkv, G, P);
  if ((p.Hq / p.Hkv) != G) return false;              // one packed tile must hold the whole GQA group
  if (P != 16 && P != 32) return false;
  const int64_t weff = p.W < p.N ? p.W : p.N;
  const int64_t nb64 = (weff - 1 + P - 1) / P + 1;
  if (nb64 * P > FusedCfg::kColsMax || (nb64 + 1) * P > FusedCfg::kRingCols) return false;
  nb = static_cast<int>(nb64);
  if ((nb - 1) * P > FusedCfg::kPartKeys) return false;
  return true;
}

template <typename T>
cudaError_t launch_fused(const AttnParams& p, int dtype, float* part, cudaStream_t st) {
  using C = FusedCfg;
  int G, P, nb;
  if (!fused_geometry(p, G, P, nb)) return cudaErrorInvalidValue;
  FusedArgs a;
  a.B = p.B; a.N = p.N; a.W = p.W; a.Hq = p.Hq; a.Hkv = p.Hkv; a.G = G; a.P = P;
  a.lgP = (P == 16) ? 4 : 5;
  a.nb = nb; a.R = nb + 1; a.cols = nb * P; a.nch = a.cols / 16;
  a.nblk = (p.N + P - 1) / P;
  a.total_tiles = a.nblk * p.Hkv * p.B;
  int ctas = fused_sm_count();
  if (ctas > C::kMaxCtas) ctas = C::kMaxCtas;
  int tpc = (a.total_tiles + ctas - 1) / ctas;
  if (tpc < nb) tpc = nb;                            // a key block is shared by at most two CTAs
  a.tiles_per_cta = tpc;
  const int grid = (a.total_tiles + tpc - 1) / tpc;
  TileMap mq, mdo, mk, mv;
  if (!make_tile_map(&mq, p.q, dtype, C::D, p.N, p.Hq, p.B, p.sq, P, G)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mdo, p.dout, dtype, C::D, p.N, p.Hq, p.B, p.sdo, P, G)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mk, p.k, dtype, C::D, p.N, p.Hkv, p.B, p.sk, a.cols, 1)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mv, p.v, dtype, C::D, p.N, p.Hkv, p.B, p.sv, a.cols, 1)) return cudaErrorInvalidValue;
  if (mq.swap_nh != mdo.swap_nh) return cudaErrorInvalidValue;
  a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh;
  a.fmt = (dtype == SFA_DTYPE_BF16) ? 1 : 0;
  a.sl2 = p.scale * kLog2e;
  a.scale = p.scale;
  a.lse = p.lse;
  a.delta = p.delta;
  a.dq = p.dq; a.dk = p.dk; a.dv = p.dv;
  a.sdq = p.sdq; a.sdk = p.sdk; a.sdv = p.sdv;
  a.part = part;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(bwd_fused64_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  bwd_fused64_kernel<T><<<grid, C::kThreads, C::kSmem, st>>>(mq.map, mdo.map, mk.map, mv.map, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (grid > 1) {
    bwd_fused_fixup_kernel<T><<<grid - 1, 256, 0, st>>>(a);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace

size_t tc_bwd_fused_workspace_bytes() {
  return static_cast<size_t>(FusedCfg::kMaxCtas) * 2 * FusedCfg::kPartKeys * 128 * sizeof(float);
}

bool tc_bwd_fused_supported(const AttnParams& p, int dtype) {
  static const bool disabled = getenv("SFA_NO_FUSED_BWD") != nullptr;
  if (disabled) return false;
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return false;
  int G, P, nb;
  if (!fused_geometry(p, G, P, nb)) return false;
  if (!(tma_compatible(p.q, p.sq) && tma_compatible(p.k, p.sk) && tma_compatible(p.v, p.sv) &&
        tma_compatible(p.dout, p.sdo)))
    return false;
  // dQ rows are written with 16-byte stores
  if (reinterpret_cast<uintptr_t>(p.dq) % 16 || p.sdq.n % 8 || p.sdq.h % 8 || p.sdq.b % 8) return false;
  const bool q_swap = (p.Hq > 1 && p.N > 1) ? (p.sq.h < p.sq.n) : false;
  const bool do_swap = (p.Hq > 1 && p.N > 1) ? (p.sdo.h < p.sdo.n) : false;
  return q_swap == do_swap;
}

cudaError_t tc_bwd_fused(const AttnParams& p, int dtype, float* part, cudaStream_t st) {
  if (dtype == SFA_DTYPE_BF16) return launch_fused<__nv_bfloat16>(p, dtype, part, st);
  return launch_fused<__half>(p, dtype, part, st);
}

}  // namespace sfa
