// Backward sink attention on the 5th-gen tensor cores (reference kernels being replaced:
// _sink_flash_attn_bwd_dq_kernel sink_flash_attention.py:371-484 and
// _sink_flash_attn_bwd_dkdv_kernel :256-364, plus the torch GQA group-sum :648-651).
//
// Two persistent kernels share the forward's packed tile (128 MMA rows = G q-heads x P positions
// of one KV head) and its two-range KV walk:
//
//  dq_kernel   (Q-stationary)   per KV tile:  S = Q K^T,  dP = dO V^T      (SS UMMAs, fp32 in TMEM)
//                                             dS = exp2(S*c - lse) * (dP - delta)  -> 16-bit, written
//                                             over the already-consumed S/dP columns
//                                             dQ += dS K                   (TS UMMA, K tile MN-major)
//              epilogue: dQ * scale -> 16-bit -> swizzled smem -> TMA store.
//
//  dkdv_kernel (KV-stationary)  one CTA per 128-key tile; per packed Q chunk that can see it:
//                                             S^T = K Q^T, dP^T = V dO^T   (keys on the TMEM lanes)
//                                             P^T, dS^T -> 16-bit in TMEM
//                                             dV += P^T dO,  dK += dS^T Q  (TS UMMAs)
//              The q heads of a GQA group are MMA rows of the same chunk, so the group sum of
//              dK/dV happens inside the fp32 accumulators (no [B,Hq,N,D] temporaries).
//
// Warp roles (320 threads): warps 0-7 element-wise math + epilogue (two warps per TMEM lane
// quarter, each taking half of the columns), warp 8 TMA producer, warp 9 tcgen05.mma issuer.
#include <stdlib.h>
#include <type_traits>

#include "attn_common.cuh"
#include "tmap.cuh"

// The in-kernel delta variant of dq64_kernel (delta = rowsum(P o dP) in the dS warps; measured slower and less accurate,
// see fuses_delta below) is compiled in only with -DSFA_DQ64_FUSE_DELTA_CODE=1: cold code in a hot role is not free in
// these kernels (bwdf_sm100.cu: 10 KB of switched-off experiments cost the fused backward 3 us).
#ifndef SFA_DQ64_FUSE_DELTA_CODE
#define SFA_DQ64_FUSE_DELTA_CODE 0
#endif

namespace sfa {
namespace {

constexpr int kMathWarps = 8;
constexpr int kMathThreads = kMathWarps * 32;
constexpr int kThreads = kMathThreads + 64;

struct BwdArgs {
  int B, N, S, W, Hq, G, P, BN, groups_per_kv, ny, nblk, total_tiles;
  unsigned long long bn_mul;   // bn_magic(BN)
  int tiles_per_cta;   // > 0: CTA c owns the contiguous tiles [c*tpc, (c+1)*tpc); 0: tiles c, c+grid, ...
  int q_swap, k_swap, v_swap, dq_swap;
  int fmt;       // 0 f16, 1 bf16
  float sl2;     // scale * log2(e)
  float scale;
  const float* lse;
  const float* delta;
  long long* trace;   // optional timeline buffer (sfa_set_trace_buffer); nullptr in production
  // fused delta (dq64 only, launches in which every tile has exactly one KV item): the dS warps compute
  // delta = rowsum(P o dP) (== rowsum(dO o O): the s_aux column has V = 0) and write it for the dK/dV kernel
  int fuse_delta;
  float* delta_out;
  float* dsrow;        // -exp(s_aux - lse) * delta per row, or nullptr
  const float* s_aux;
  // extended geometry of the shared tile walker (attn_common.cuh); these kernels run with q_off = 0 and no packing
  int q_off;
  const int* seq_lo;
  int64_t seq_bs;
  int dbg_delay;       // test knob (sfa_set_debug 0): half of the math warps sleep this many ns inside their passes
};

// Timeline probe for performance work: CTA 0 appends (role, code, index, clock64) records.
// role 0 = TMA producer, 1 = MMA issuer, 2 = math thread 0.  256 records of 2 x int64 per role.
#ifndef SFA_TRACE
#define SFA_TRACE 0      // build with -DSFA_TRACE=1 to compile the timeline probe in
#endif
__device__ __forceinline__ void trace_ev(long long* trace, int role, int& cnt, int code, int idx) {
  if (SFA_TRACE && trace != nullptr && blockIdx.x == 0 && cnt < 256) {
    trace[(role * 256 + cnt) * 2] = (static_cast<long long>(code) << 32) | static_cast<unsigned>(idx);
    trace[(role * 256 + cnt) * 2 + 1] = clock64();
    ++cnt;
  }
}

template <typename T> __device__ __forceinline__ void unpack16(uint32_t u, float& a, float& b);
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(uint32_t u, float& a, float& b) {
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack16<__half>(uint32_t u, float& a, float& b) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
  a = f.x;
  b = f.y;
}

// dS = P o t with P held as a packed 16-bit pair and t in fp32: ONE rounding of the product (P's own rounding is the
// one the dV / PV operand has anyway).  A packed 16-bit multiply (P * round(t), rounded again) measurably widened the
// dK error where 22 000 terms meet in one key (GQA group of 32, window 700).
template <typename T> __device__ __forceinline__ uint32_t ds_pair(uint32_t p_pair, float t0, float t1) {
  float p0, p1;
  unpack16<T>(p_pair, p0, p1);
  return pack16_fast<T>(p0 * t0, p1 * t1);
}

// ================================================================================== dQ kernel
// Work items = (packed Q tile, KV tile) pairs in the order of the two-range walk, two TMEM slots:
// the UMMAs of item n+1 (S, dP) run while the math warps work on item n, and dQ(n) runs under the
// math of item n+1.  (head_dim 64 has its own kernel below; this one serves 64 < D <= 128.)
//
// Shared memory at head_dim 128: a tile's Q and dO are 32 KB each, a K or V item 24 KB.  Round 1 held two (Q, dO)
// stages and only two K slots, so K(n + 2) could be requested only when dQ(n) had completed -- right when S(n + 2)
// wanted it: one exposed TMA latency per item.  Now THREE 32 KB buffers rotate through the roles
// Q(t), dO(t), Q(t+1), dO(t+1), ... (use u = 2t / 2t + 1 -> buffer u % 3): Q(t + 1) is prefetched into the buffer
// dO(t - 1) left (after it served as the staging tile of dQ(t - 1)'s store), dO(t + 1) goes where Q(t) was as soon
// as the tile's last S is complete; that frees the room for a K ring of three.
#ifndef SFA_DQ_BN
#define SFA_DQ_BN 96
#define SFA_DQ_KST 3
#define SFA_DQ_VST 2
#endif
template <int D> struct DqCfg {
  static constexpr int kDS = D / 64;
  static constexpr int kBNMax = SFA_DQ_BN;                      // KV rows per item (UMMA N of S and dP)
  static constexpr int kKStages = SFA_DQ_KST;                     // K is held from S(n) to dQ(n)
  static constexpr int kVStages = SFA_DQ_VST;
  static constexpr int kQBufs = 3;                       // rotating Q / dO / staging buffers
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBNMax * D * 2;
  static constexpr int kSlabQ = 128 * 128;
  static constexpr int kSlabKV = kBNMax * 128;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kSlotCols = 2 * kBNMax;      // S then dP
  static constexpr uint32_t kColQ = 2 * kSlotCols;       // dQ accumulator
  static constexpr int kSmem = 1024 + kQBufs * kQBytes + (kKStages + kVStages) * kKVBytes + 512;
  static_assert(2 * kSlotCols + D <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

using ItemWalk = ItemWalkT<BwdArgs>;

template <typename T, int D>
__global__ void __launch_bounds__(kThreads, 1) dq_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                         const __grid_constant__ CUtensorMap tmdO,
                                                         const __grid_constant__ CUtensorMap tmK,
                                                         const __grid_constant__ CUtensorMap tmV,
                                                         const __grid_constant__ CUtensorMap tmdQ, const BwdArgs a) {
  using C = DqCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* qb_s = smem;                                  // [kQBufs][kQBytes]  use u -> buffer u % 3
  unsigned char* k_s = qb_s + C::kQBufs * C::kQBytes;
  unsigned char* v_s = k_s + C::kKStages * C::kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + C::kVStages * C::kKVBytes);
  uint64_t* qb_full = bars;                        // [kQBufs]  use u landed: phase u / 3
  uint64_t* qb_empty = qb_full + C::kQBufs;        // [kQBufs]  Q use: the tile's last S complete; dO use: dQ store has read the staging tile
  uint64_t* k_full = qb_empty + C::kQBufs;         // [kKStages]
  uint64_t* k_empty = k_full + C::kKStages;
  uint64_t* v_full = k_empty + C::kKStages;        // [kVStages]
  uint64_t* v_empty = v_full + C::kVStages;
  uint64_t* s_full = v_empty + C::kVStages;        // [2]  S and dP of the slot complete
  uint64_t* p_full = s_full + 2;                   // [2]
  uint64_t* dq_done = p_full + 2;
  uint64_t* dq_free = dq_done + 1;
  uint64_t* s1_full = dq_free + 1;                 // [2]  S of the slot complete: the exponentials start under the dP UMMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s1_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kMathWarps && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdQ);
    for (int s = 0; s < C::kQBufs; ++s) {
      mbar_init(qb_full + s, 1);
      mbar_init(qb_empty + s, 1);
    }
    for (int s = 0; s < C::kKStages; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
    }
    for (int s = 0; s < C::kVStages; ++s) {
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + s, 1);
      mbar_init(s1_full + s, 1);
      mbar_init(p_full + s, kMathThreads);
    }
    mbar_init(dq_done, 1);
    mbar_init(dq_free, kMathThreads);
    fence_barrier_init();
  }
  if (warp == kMathWarps + 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kMathWarps) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ItemWalk w(a);
      int tc = 0;
      int q_loaded = 0;        // tiles (w.it) whose Q has been requested
      // Q or dO tile of the tile (pb, y, b) as use u of the rotating buffers
      auto load_tile = [&](const CUtensorMap* tm, int u, int q0, int hq0, int b) {
        const int bf = u % C::kQBufs;
        mbar_wait(qb_empty + bf, ((u / C::kQBufs) & 1) ^ 1);
        mbar_expect_tx(qb_full + bf, C::kQBytes);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(qb_s + bf * C::kQBytes + s * C::kSlabQ, tm, qb_full + bf, a.q_swap, s * 64, q0, hq0, b);
      };
      while (w.next()) {
        const int hq0 = w.y * a.G, kvh = w.y / a.groups_per_kv;
        if (w.t == 0) {
          if (q_loaded <= w.it) {
            load_tile(&tmQ, 2 * w.it, w.q0, hq0, w.b);
            q_loaded = w.it + 1;
          }
          trace_ev(a.trace, 0, tc, 1, w.it);
          load_tile(&tmdO, 2 * w.it + 1, w.q0, hq0, w.b);       // waits for the previous tile's last S
        }
        // prefetch of the NEXT tile's Q, two items into this tile: its buffer was the staging tile of the previous
        // tile's dQ store (issued by the math warps after this tile's first item)
        if (w.t == 2 && q_loaded == w.it + 1 && w.tile + w.step < w.end) {
          int pb = w.pb, y = w.y, b = w.b;
          ItemWalk::advance(a, w.step, pb, y, b);
          load_tile(&tmQ, 2 * (w.it + 1), pb * a.P, y * a.G, b);
          q_loaded = w.it + 2;
        }
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const int kst = w.n % C::kKStages, vst = w.n % C::kVStages;
        mbar_wait(k_empty + kst, ((w.n / C::kKStages) & 1) ^ 1);
        trace_ev(a.trace, 0, tc, 2, w.n);        // K stage free
        mbar_expect_tx(k_full + kst, a.BN * D * 2);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(k_s + kst * C::kKVBytes + s * C::kSlabKV, &tmK, k_full + kst, a.k_swap, s * 64, kstart, kvh, w.b);
        mbar_wait(v_empty + vst, ((w.n / C::kVStages) & 1) ^ 1);
        trace_ev(a.trace, 0, tc, 3, w.n);        // V stage free
        mbar_expect_tx(v_full + vst, a.BN * D * 2);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(v_s + vst * C::kKVBytes + s * C::kSlabKV, &tmV, v_full + vst, a.v_swap, s * 64, kstart, kvh, w.b);
      }
    }
    __syncwarp();
  } else if (warp == kMathWarps + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_dq = make_idesc(a.fmt, 128, D, 0, 1);
      int tc = 0;
      // S = Q K^T and dP = dO V^T of the item `w` points at, into its TMEM slot
      auto issue_sdp = [&](const ItemWalk& w) {
        const int slot = w.n & 1;
        const int uq = 2 * w.it, ud = 2 * w.it + 1;
        const int bq = uq % C::kQBufs, bd = ud % C::kQBufs;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const int kst = w.n % C::kKStages, vst = w.n % C::kVStages;
        const uint32_t idesc_s = make_idesc(a.fmt, 128, cols, 0, 0);
        // descriptors of the first K-step; the next ones are +32 B (= +2 in the encoded address field)
        const uint64_t qd = make_sdesc(smem_u32(qb_s + bq * C::kQBytes), 16, 1024);
        const uint64_t dod = make_sdesc(smem_u32(qb_s + bd * C::kQBytes), 16, 1024);
        const uint64_t kd = make_sdesc(smem_u32(k_s + kst * C::kKVBytes), 16, 1024);
        const uint64_t vd = make_sdesc(smem_u32(v_s + vst * C::kKVBytes), 16, 1024);
        const uint32_t ts = tmem + slot * C::kSlotCols;
        if (w.t == 0) mbar_wait(qb_full + bq, (uq / C::kQBufs) & 1);
        mbar_wait(k_full + kst, (w.n / C::kKStages) & 1);
        tc_fence_after();
        trace_ev(a.trace, 1, tc, 2, w.n);        // K landed
#pragma unroll
        for (int s = 0; s < C::kDS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(ts, qd + ((s * C::kSlabQ + kk * 32) >> 4), kd + ((s * C::kSlabKV + kk * 32) >> 4), idesc_s, (s | kk) != 0);
        umma_commit(s1_full + slot);
        if (w.last_of_tile()) umma_commit(qb_empty + bq);       // Q is only read by S: its buffer takes the next tile's dO
        trace_ev(a.trace, 1, tc, 5, w.n);        // S UMMAs issued
        if (w.t == 0) mbar_wait(qb_full + bd, (ud / C::kQBufs) & 1);
        mbar_wait(v_full + vst, (w.n / C::kVStages) & 1);
        tc_fence_after();
        trace_ev(a.trace, 1, tc, 6, w.n);        // V landed
#pragma unroll
        for (int s = 0; s < C::kDS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(ts + C::kBNMax, dod + ((s * C::kSlabQ + kk * 32) >> 4), vd + ((s * C::kSlabKV + kk * 32) >> 4), idesc_s,
                    (s | kk) != 0);
        trace_ev(a.trace, 1, tc, 7, w.n);        // dP UMMAs issued
        umma_commit(v_empty + vst);
        umma_commit(s_full + slot);
        trace_ev(a.trace, 1, tc, 3, w.n);        // S, dP issued
      };
      ItemWalk w_sdp(a), w_dq(a);
      if (w_sdp.next()) issue_sdp(w_sdp);
      while (w_dq.next()) {
        if (w_sdp.next()) issue_sdp(w_sdp);          // item n+1 runs under the math of item n
        const int slot = w_dq.n & 1;
        const int kst = w_dq.n % C::kKStages;
        int kstart, cols; bool is_sink;
        w_dq.pl.tile(w_dq.t, a.BN, kstart, cols, is_sink);
        mbar_wait(p_full + slot, (w_dq.n >> 1) & 1);
        tc_fence_after();
        trace_ev(a.trace, 1, tc, 4, w_dq.n);     // dS ready
        if (w_dq.t == 0 && w_dq.it >= 1) {
          mbar_wait(dq_free, (w_dq.it - 1) & 1);
          tc_fence_after();
        }
        // dS is held as 16-bit pairs: first half of the key columns over dP, second half over S
        const uint32_t ts = tmem + slot * C::kSlotCols;
        const uint64_t kd = make_sdesc(smem_u32(k_s + kst * C::kKVBytes), C::kSlabKV, 1024);
        const int hcol = ((cols / 16 + 1) / 2) * 16;
        const uint32_t dqa = tmem + C::kColQ;
        const uint32_t a_lo = ts + C::kBNMax, a_hi = ts + hcol - (hcol >> 1);   // + c0/2 in both halves
        const int nk = cols >> 4;
#pragma unroll
        for (int kk = 0; kk < C::kBNMax / 16; ++kk)
          if (kk < nk)
            umma_ts(dqa, ((kk * 16 < hcol) ? a_lo : a_hi) + kk * 8, kd + kk * (2048 >> 4), idesc_dq, (w_dq.t > 0 || kk > 0));
        trace_ev(a.trace, 1, tc, 8, w_dq.n);     // dQ UMMAs issued
        umma_commit(k_empty + kst);
        if (w_dq.last_of_tile()) umma_commit(dq_done);
        trace_ev(a.trace, 1, tc, 9, w_dq.n);     // dQ committed
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ element-wise math + epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;                  // MMA row == TMEM lane
    const int pr = a.q_swap ? (r / a.G) : (r % a.P);
    const int gr = a.q_swap ? (r % a.G) : (r / a.P);
    const int ro = a.dq_swap ? (pr * a.G + gr) : (gr * a.P + pr);   // row in dQ's box order
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);

    // lse and delta of this thread's row in the tile (pb, y, b): loaded one tile ahead, consumed (and only then
    // waited for) when that tile starts
    auto load_row = [&](bool valid, int pb, int y, int b, float& l, float& dl) {
      l = INFINITY;                      // rows past N: P = exp2(s - inf) = 0
      dl = 0.f;
      const int i = pb * a.P + pr;
      if (valid && i < a.N) {
        const int64_t row = (static_cast<int64_t>(b) * a.Hq + y * a.G + gr) * a.N + i;
        // volatile: the loads must be issued here (a tile ahead of their use), not sunk to the use
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(l) : "l"(a.lse + row));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dl) : "l"(a.delta + row));
      }
    };
    struct Pending { int valid, it, q0, hq0, b; } pend = {0, 0, 0, 0, 0};
    auto epilogue = [&](const Pending& e) {
      // staging tile of the dQ store = the buffer the tile's dO sat in (all of the tile's UMMAs are complete)
      const int bd = (2 * e.it + 1) % C::kQBufs;
      unsigned char* stage = qb_s + bd * C::kQBytes;
      mbar_wait(dq_done, e.it & 1);
      tc_fence_after();
      uint32_t v[D / 32][16];
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) tmem_ld16(tl + C::kColQ + half * (D / 2) + cc * 16, v[cc]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc) {
        const int c0 = half * (D / 2) + cc * 16;
        uint32_t pk[8];
#pragma unroll
        for (int e2 = 0; e2 < 16; e2 += 2)
          pk[e2 >> 1] = pack16<T>(__uint_as_float(v[cc][e2]) * a.scale, __uint_as_float(v[cc][e2 + 1]) * a.scale);
        unsigned char* slab = stage + (c0 >> 6) * C::kSlabQ;
        const int chn = (c0 & 63) >> 3;
        *reinterpret_cast<uint4*>(slab + sw128_off(ro, chn)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(slab + sw128_off(ro, chn + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, kMathThreads);
      if (threadIdx.x == 0) {
        for (int s = 0; s < C::kDS; ++s) tma_tile_store(&tmdQ, stage + s * C::kSlabQ, a.dq_swap, s * 64, e.q0, e.hq0, e.b);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(qb_empty + bd);        // the buffer may take the Q of the tile after next
      }
    };

    float l_next, dl_next, neg_l2 = -INFINITY, delta = 0.f;
    const uint64_t sl2_2 = pack_f32x2(a.sl2, a.sl2), one2 = pack_f32x2(1.f, 1.f);
    uint64_t negl2_2 = 0, ndelta2 = 0;
    int i = 0, mtc = 0;
    ItemWalk w(a);
    {
      int pb = w.pb, y = w.y, b = w.b;
      ItemWalk::advance(a, w.step, pb, y, b);
      load_row(w.tile + w.step < w.end, pb, y, b, l_next, dl_next);
    }
    while (w.next()) {
      if (w.t == 0) {
        neg_l2 = (l_next == -INFINITY) ? -INFINITY : -l_next * kLog2e;   // lse = -inf: nothing attended, P = 0
        delta = dl_next;
        negl2_2 = pack_f32x2(neg_l2, neg_l2);
        ndelta2 = pack_f32x2(-delta, -delta);
        i = w.q0 + pr;
        int pb = w.pb, y = w.y, b = w.b;
        ItemWalk::advance(a, w.step, pb, y, b);
        load_row(w.tile + w.step < w.end, pb, y, b, l_next, dl_next);
      }
      const int slot = w.n & 1;
      const uint32_t ts = tl + slot * C::kSlotCols;
      int kstart, cols; bool is_sink;
      w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
      int c_lo, c_hi;
      row_range(is_sink, i, kstart, cols, a.S, a.W, c_lo, c_hi);
      if (a.seq_lo != nullptr && i < a.N) c_lo = max(c_lo, __ldg(a.seq_lo + w.b * a.seq_bs + i) - kstart);   // packed sequences
      if (i >= a.N) c_hi = -1;
      const int nch = cols / 16;
      const int hch = (nch + 1) / 2;
      const int hcol = hch * 16;
      const int ch0 = half ? hch : 0, ch1 = half ? nch : hch;
      if (threadIdx.x == 0) trace_ev(a.trace, 2, mtc, 1, w.n);     // waiting for S
      // ---- phase 1 (under the dP UMMAs): P = exp2(S * c - lse) of this thread's columns -> packed 16-bit in registers
      mbar_wait(s1_full + slot, (w.n >> 1) & 1);
      tc_fence_after();
      constexpr int kMaxCh = (C::kBNMax / 16 + 1) / 2;
      uint32_t xv[kMaxCh][16], pp[kMaxCh][8];
#pragma unroll
      for (int jj = 0; jj < kMaxCh; ++jj)
        if (ch0 + jj < ch1) tmem_ld16(ts + (ch0 + jj) * 16, xv[jj]);
      tmem_ld_wait();
      // every column of the item attended by every row of the warp (all items but the ends of a band): no per-chunk tests
      const bool interior = __all_sync(0xffffffffu, c_lo <= 0 && c_hi >= cols - 1);
#pragma unroll
      for (int jj = 0; jj < kMaxCh; ++jj)
        if (ch0 + jj < ch1) {
          const int c0 = (ch0 + jj) * 16;
          if (interior || __all_sync(0xffffffffu, (c0 >= c_lo) && (c0 + 15 <= c_hi))) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float x0, x1;
              unpack_f32x2(fma_f32x2(xv[jj][e], xv[jj][e + 1], sl2_2, negl2_2), x0, x1);
              pp[jj][e >> 1] = pack16_fast<T>(fast_exp2(x0), fast_exp2(x1));
            }
          } else if (c0 + 15 >= c_lo && c0 <= c_hi) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int c = c0 + e;
              float p0 = fast_exp2(fmaf(__uint_as_float(xv[jj][e]), a.sl2, neg_l2));
              float p1 = fast_exp2(fmaf(__uint_as_float(xv[jj][e + 1]), a.sl2, neg_l2));
              p0 = (c >= c_lo && c <= c_hi) ? p0 : 0.f;
              p1 = (c + 1 >= c_lo && c + 1 <= c_hi) ? p1 : 0.f;
              pp[jj][e >> 1] = pack16_fast<T>(p0, p1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) pp[jj][e] = 0u;
          }
          __syncwarp();
        }
      // ---- phase 2: dS = P o (dP - delta): t = dP - delta rounded to 16 bit, one packed multiply with the P pair (masked
      // P is exactly 0 and dP is finite: dS = 0 there).  This pass sits on the dP -> dS -> dQ critical chain: the packed
      // multiply is 5 % faster than the fp32 product here, and dQ sums over one row's keys only (its error stays at
      // 0.4 of the on-device bar where dK, summing 22 000 terms per key, needs the single rounding: ds_pair)
      mbar_wait(s_full + slot, (w.n >> 1) & 1);
      tc_fence_after();
      if (a.dbg_delay && half == 1) __nanosleep(a.dbg_delay);
      if (threadIdx.x == 0) trace_ev(a.trace, 2, mtc, 2, w.n);     // S, dP complete
#pragma unroll
      for (int jj = 0; jj < kMaxCh; ++jj)
        if (ch0 + jj < ch1) tmem_ld16(ts + C::kBNMax + (ch0 + jj) * 16, xv[jj]);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < kMaxCh; ++jj)
        if (ch0 + jj < ch1) {
          const int c0 = (ch0 + jj) * 16;
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float t0, t1;
            unpack_f32x2(fma_f32x2(xv[jj][e], xv[jj][e + 1], one2, ndelta2), t0, t1);
            pk[e >> 1] = mul16x2<T>(pp[jj][e >> 1], pack16_fast<T>(t0, t1));
          }
          const uint32_t dst = half ? (ts + hcol + ((c0 - hcol) >> 1)) : (ts + C::kBNMax + (c0 >> 1));
          tmem_st8(dst, pk);
        }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + slot);
      if (threadIdx.x == 0) trace_ev(a.trace, 2, mtc, 3, w.n);     // dS written
      if (pend.valid) {
        epilogue(pend);
        pend.valid = 0;
        if (threadIdx.x == 0) trace_ev(a.trace, 2, mtc, 4, w.n);   // deferred epilogue done
      }
      if (w.last_of_tile()) pend = Pending{1, w.it, w.q0, w.y * a.G, w.b};
    }
    if (pend.valid) epilogue(pend);
    if (threadIdx.x == 0) tma_store_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMathWarps + 1) tmem_dealloc(tmem, C::kTmemCols);
}

// ================================================================================== dQ kernel, head_dim 64
// Measured on B200 (tools/probe_mma.py, tools/trace_dq.py): one tcgen05.mma costs its issuing thread
// ~70-90 cycles whatever its N up to 144, every mbarrier round trip on a single-thread role ~200-300
// cycles, MUFU.EX2 runs at 16/clk/SM (a 128 x 144 tile of exps is >= 1152 cycles) and a TMA load
// takes 1-2.5 us under load.  Hence a deep, warp-specialised pipeline in which no role does two jobs:
//
//   warp 12      TMA producer          Q/K rings (S side), dO/V rings (dP side)
//   warp 13      UMMA issuer S         S(n)  = Q K^T   -> S buffer n & 1 (two items ahead of dQ)
//   warp 14      UMMA issuer dP        dP(n) = dO V^T  -> the single dP region
//   warp 15      UMMA issuer dQ        dQ   += dS(n) K (TS form, dS read from the S buffer)
//   warps 0-3    exp                   P  = exp2(S*c - lse)           -> 16-bit over S columns [0, 72)
//   warps 4-7    dS                    dS = P * (dP - delta), masked  -> 16-bit over S columns [72, 144)
//   warps 8-11   epilogue              dQ * scale -> 16-bit -> swizzled smem -> TMA store
//
// KV items are up to 144 columns (one item per tile at window 128).
struct Dq64Cfg {
  static constexpr int D = 64;
  static constexpr int kBNMax = 144;
  static constexpr int kKStages = 4;      // K(n) is held from S(n) (two items ahead) to dQ(n)
  static constexpr int kVStages = 2;
  static constexpr int kQStages = 3;      // Q tiles (S side)
  static constexpr int kOStages = 3;      // dO tiles (dP side)
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBNMax * D * 2;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS = 0;               // S buffers at 0 and kBNMax; P at +0, dS at +kBNMax/2
  static constexpr uint32_t kColP = 2 * kBNMax;      // dP
  static constexpr uint32_t kColQ = 3 * kBNMax;      // dQ accumulator
  static constexpr int kThreads = 16 * 32;
  static constexpr int kSmem = 1024 + (kQStages + kOStages + 1) * kQBytes + (kKStages + kVStages) * kKVBytes + 512;
  static_assert(3 * kBNMax + D <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

template <typename T>
__global__ void __launch_bounds__(Dq64Cfg::kThreads, 1) dq64_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmdO,
                                                                    const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV,
                                                                    const __grid_constant__ CUtensorMap tmdQ,
                                                                    const BwdArgs a) {
  using C = Dq64Cfg;
  constexpr int D = C::D;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = smem;                                   // [kQStages][kQBytes]
  unsigned char* do_s = q_s + C::kQStages * C::kQBytes;        // [kOStages][kQBytes]
  unsigned char* stage_s = do_s + C::kOStages * C::kQBytes;    // dQ staging
  unsigned char* k_s = stage_s + C::kQBytes;                   // [kKStages][kKVBytes]
  unsigned char* v_s = k_s + C::kKStages * C::kKVBytes;        // [kVStages][kKVBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + C::kVStages * C::kKVBytes);
  uint64_t* q_full = bars;
  uint64_t* q_empty = q_full + C::kQStages;
  uint64_t* do_full = q_empty + C::kQStages;
  uint64_t* do_empty = do_full + C::kOStages;
  uint64_t* k_full = do_empty + C::kOStages;
  uint64_t* k_empty = k_full + C::kKStages;
  uint64_t* v_full = k_empty + C::kKStages;
  uint64_t* v_empty = v_full + C::kVStages;
  uint64_t* s_full = v_empty + C::kVStages;        // [2]  S(n) complete                       (issuer S -> exp)
  uint64_t* p1_done = s_full + 2;                  // [2]  P(n) written                        (exp -> dS warps)
  uint64_t* dp_full = p1_done + 2;                 //      dP(n) complete                      (issuer dP -> dS warps)
  uint64_t* p_full = dp_full + 1;                  // [2]  dS(n) written, dP(n) consumed       (dS warps -> issuers dQ, dP)
  uint64_t* sbuf_free = p_full + 2;                // [2]  dQ(n) complete: S buffer n & 1 free (issuer dQ -> issuer S)
  uint64_t* dq_done = sbuf_free + 2;               //      tile's dQ complete                  (issuer dQ -> epilogue)
  uint64_t* dq_free = dq_done + 1;                 //      dQ accumulator read                 (epilogue -> issuer dQ)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdQ);
    for (int s = 0; s < C::kQStages; ++s) { mbar_init(q_full + s, 1); mbar_init(q_empty + s, 1); }
    for (int s = 0; s < C::kOStages; ++s) { mbar_init(do_full + s, 1); mbar_init(do_empty + s, 1); }
    for (int s = 0; s < C::kKStages; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < C::kVStages; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + s, 1);
      mbar_init(p1_done + s, 128);
      mbar_init(p_full + s, 128);
      mbar_init(sbuf_free + s, 1);
    }
    mbar_init(dp_full, 1);
    mbar_init(dq_done, 1);
    mbar_init(dq_free, 128);
    fence_barrier_init();
  }
  if (warp == 13) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ItemWalk w(a);
      while (w.next()) {
        const int hq0 = w.y * a.G, kvh = w.y / a.groups_per_kv;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const int kst = w.n % C::kKStages, vst = w.n % C::kVStages;
        if (w.t == 0) {
          const int qs = w.it % C::kQStages;
          mbar_wait(q_empty + qs, ((w.it / C::kQStages) & 1) ^ 1);
          mbar_expect_tx(q_full + qs, C::kQBytes);
          tma_tile(q_s + qs * C::kQBytes, &tmQ, q_full + qs, a.q_swap, 0, w.q0, hq0, w.b);
        }
        mbar_wait(k_empty + kst, ((w.n / C::kKStages) & 1) ^ 1);
        mbar_expect_tx(k_full + kst, a.BN * D * 2);
        tma_tile(k_s + kst * C::kKVBytes, &tmK, k_full + kst, a.k_swap, 0, kstart, kvh, w.b);
        if (w.t == 0) {
          const int os = w.it % C::kOStages;
          mbar_wait(do_empty + os, ((w.it / C::kOStages) & 1) ^ 1);
          mbar_expect_tx(do_full + os, C::kQBytes);
          tma_tile(do_s + os * C::kQBytes, &tmdO, do_full + os, a.q_swap, 0, w.q0, hq0, w.b);
        }
        mbar_wait(v_empty + vst, ((w.n / C::kVStages) & 1) ^ 1);
        mbar_expect_tx(v_full + vst, a.BN * D * 2);
        tma_tile(v_s + vst * C::kKVBytes, &tmV, v_full + vst, a.v_swap, 0, kstart, kvh, w.b);
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ------------------------------------------------------------------ UMMA issuer S
    if (lane == 0) {
      ItemWalk w(a);
      int tc = 0;
      while (w.next()) {
        trace_ev(a.trace, 1, tc, 1, w.n);
        const int qs = w.it % C::kQStages, kst = w.n % C::kKStages, sb = w.n & 1;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const uint32_t idesc_s = make_idesc(a.fmt, 128, cols, 0, 0);
        const uint64_t qd = make_sdesc(smem_u32(q_s + qs * C::kQBytes), 16, 1024);
        const uint64_t kd = make_sdesc(smem_u32(k_s + kst * C::kKVBytes), 16, 1024);
        if (w.t == 0) mbar_wait(q_full + qs, (w.it / C::kQStages) & 1);
        mbar_wait(k_full + kst, (w.n / C::kKStages) & 1);
        if (w.n >= 2) mbar_wait(sbuf_free + sb, ((w.n - 2) >> 1) & 1);     // dQ(n-2) has consumed dS(n-2)
        tc_fence_after();
        trace_ev(a.trace, 1, tc, 2, w.n);
        const uint32_t ts = tmem + C::kColS + sb * C::kBNMax;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(ts, qd + kk * 2, kd + kk * 2, idesc_s, kk != 0);
        umma_commit(s_full + sb);
        if (w.last_of_tile()) umma_commit(q_empty + qs);
        trace_ev(a.trace, 1, tc, 3, w.n);
      }
    }
    __syncwarp();
  } else if (warp == 14) {
    // ------------------------------------------------------------------ UMMA issuer dP
    if (lane == 0) {
      ItemWalk w(a);
      int tc = 0;
      while (w.next()) {
        trace_ev(a.trace, 2, tc, 1, w.n);
        const int os = w.it % C::kOStages, vst = w.n % C::kVStages;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const uint32_t idesc_s = make_idesc(a.fmt, 128, cols, 0, 0);
        const uint64_t dod = make_sdesc(smem_u32(do_s + os * C::kQBytes), 16, 1024);
        const uint64_t vd = make_sdesc(smem_u32(v_s + vst * C::kKVBytes), 16, 1024);
        if (w.t == 0) mbar_wait(do_full + os, (w.it / C::kOStages) & 1);
        mbar_wait(v_full + vst, (w.n / C::kVStages) & 1);
        if (w.n >= 1) mbar_wait(p_full + ((w.n - 1) & 1), ((w.n - 1) >> 1) & 1);   // dP(n-1) has been read
        tc_fence_after();
        trace_ev(a.trace, 2, tc, 2, w.n);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C::kColP, dod + kk * 2, vd + kk * 2, idesc_s, kk != 0);
        umma_commit(dp_full);
        umma_commit(v_empty + vst);
        if (w.last_of_tile()) umma_commit(do_empty + os);
        trace_ev(a.trace, 2, tc, 3, w.n);
      }
    }
    __syncwarp();
  } else if (warp == 15) {
    // ------------------------------------------------------------------ UMMA issuer dQ
    if (lane == 0) {
      const uint32_t idesc_dq = make_idesc(a.fmt, 128, D, 0, 1);
      ItemWalk w(a);
      int tc = 0;
      while (w.next()) {
        trace_ev(a.trace, 3, tc, 1, w.n);
        const int kst = w.n % C::kKStages, sb = w.n & 1;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const uint64_t kd = make_sdesc(smem_u32(k_s + kst * C::kKVBytes), C::kKVBytes, 1024);
        const uint32_t ads = tmem + C::kColS + sb * C::kBNMax + C::kBNMax / 2;    // dS(n), 8 columns per 16 keys
        const int nk = cols >> 4;
        mbar_wait(p_full + sb, (w.n >> 1) & 1);
        if (w.t == 0 && w.it > 0) mbar_wait(dq_free, (w.it - 1) & 1);
        tc_fence_after();
        trace_ev(a.trace, 3, tc, 2, w.n);
#pragma unroll
        for (int kk = 0; kk < C::kBNMax / 16; ++kk)
          if (kk < nk) umma_ts(tmem + C::kColQ, ads + kk * 8, kd + kk * (2048 >> 4), idesc_dq, (w.t > 0 || kk > 0));
        umma_commit(sbuf_free + sb);
        umma_commit(k_empty + kst);
        if (w.last_of_tile()) umma_commit(dq_done);
        trace_ev(a.trace, 3, tc, 3, w.n);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ math roles: TMEM lane == MMA row
    const int role = warp >> 2;                          // 0 exp, 1 dS, 2 epilogue
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int pr = a.q_swap ? (r / a.G) : (r % a.P);
    const int gr = a.q_swap ? (r % a.G) : (r / a.P);
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);

    // one fp32 per (tile, row) from a [B,Hq,N] array, loaded one tile ahead of its use
    auto load_row = [&](const float* src, bool valid, int pb, int y, int b, float dflt) {
      float v = dflt;
      const int i = pb * a.P + pr;
      if (valid && i < a.N) {
        const int64_t row = (static_cast<int64_t>(b) * a.Hq + y * a.G + gr) * a.N + i;
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(src + row));
      }
      return v;
    };

    if (role == 0) {
      // ---------------------------------------------------------------- exp warps
      ItemWalk w(a);
      float l_next, neg_l2 = -INFINITY;
      int i = 0, mtc = 0;
      {
        int pb = w.pb, y = w.y, b = w.b;
        ItemWalk::advance(a, w.step, pb, y, b);
        l_next = load_row(a.lse, w.tile + w.step < w.end, pb, y, b, INFINITY);
      }
      while (w.next()) {
        if (w.t == 0) {
          neg_l2 = (l_next == -INFINITY) ? -INFINITY : -l_next * kLog2e;   // lse = +-inf (no row / nothing attended): P = 0
          i = w.q0 + pr;
          int pb = w.pb, y = w.y, b = w.b;
          ItemWalk::advance(a, w.step, pb, y, b);
          l_next = load_row(a.lse, w.tile + w.step < w.end, pb, y, b, INFINITY);
        }
        const int sb = w.n & 1;
        const uint32_t ts = tl + C::kColS + sb * C::kBNMax;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        int c_lo, c_hi;
        row_range(is_sink, i, kstart, cols, a.S, a.W, c_lo, c_hi);
        if (a.seq_lo != nullptr && i < a.N) c_lo = max(c_lo, __ldg(a.seq_lo + w.b * a.seq_bs + i) - kstart);   // packed sequences
        if (i >= a.N) c_hi = -1;
        const int nch = cols >> 4;
        if (threadIdx.x == 0) trace_ev(a.trace, 4, mtc, 1, w.n);
        mbar_wait(s_full + sb, (w.n >> 1) & 1);
        tc_fence_after();
        if (threadIdx.x == 0) trace_ev(a.trace, 4, mtc, 2, w.n);
        // In place, one 16-column chunk per trip, ONE code path (masked elements are zeroed later by the dS warps):
        // ncu showed the math warps starved for instructions (stall_no_inst 50-80 % in unrolled multi-variant
        // bodies), so the loop body is kept small enough to stay resident in the instruction cache.
#pragma unroll 1
        for (int cb = 0; cb < nch; ++cb) {
          uint32_t sv[16], pk[8];
          tmem_ld16(ts + cb * 16, sv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 2)
            pk[e >> 1] = pack16_fast<T>(fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, neg_l2)),
                                        fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, neg_l2)));
          tmem_st8(ts + cb * 8, pk);
        }
        tmem_st_wait();
        if (threadIdx.x == 0) trace_ev(a.trace, 4, mtc, 7, w.n);
        tc_fence_before();
        mbar_arrive(p1_done + sb);
        if (threadIdx.x == 0) trace_ev(a.trace, 4, mtc, 3, w.n);
      }
    } else if (role == 1) {
      // ---------------------------------------------------------------- dS warps
      ItemWalk w(a);
      // prefetched per tile: delta of this row (preprocessed), or -- with fused delta -- its lse (for ds_aux)
      const float* pre_src = (SFA_DQ64_FUSE_DELTA_CODE && a.fuse_delta) ? a.lse : a.delta;
      const bool pre_on = !(SFA_DQ64_FUSE_DELTA_CODE && a.fuse_delta) || a.dsrow != nullptr;
      float pre_next, pre_cur = 0.f, delta = 0.f;
      int i = 0, mtc = 0;
      {
        int pb = w.pb, y = w.y, b = w.b;
        ItemWalk::advance(a, w.step, pb, y, b);
        pre_next = load_row(pre_src, pre_on && (w.tile + w.step < w.end), pb, y, b, 0.f);
      }
      while (w.next()) {
        if (w.t == 0) {
          pre_cur = pre_next;
          delta = pre_cur;
          i = w.q0 + pr;
          int pb = w.pb, y = w.y, b = w.b;
          ItemWalk::advance(a, w.step, pb, y, b);
          pre_next = load_row(pre_src, pre_on && (w.tile + w.step < w.end), pb, y, b, 0.f);
        }
        const int sb = w.n & 1;
        const uint32_t ts = tl + C::kColS + sb * C::kBNMax;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        int c_lo, c_hi;
        row_range(is_sink, i, kstart, cols, a.S, a.W, c_lo, c_hi);
        if (a.seq_lo != nullptr && i < a.N) c_lo = max(c_lo, __ldg(a.seq_lo + w.b * a.seq_bs + i) - kstart);   // packed sequences
        if (i >= a.N) c_hi = -1;
        const int nch = cols >> 4;
        if (threadIdx.x == 128) trace_ev(a.trace, 5, mtc, 1, w.n);
        mbar_wait(p1_done + sb, (w.n >> 1) & 1);
        if (threadIdx.x == 128) trace_ev(a.trace, 5, mtc, 4, w.n);
        mbar_wait(dp_full, w.n & 1);
        tc_fence_after();
        if (threadIdx.x == 128) trace_ev(a.trace, 5, mtc, 2, w.n);
        if (SFA_DQ64_FUSE_DELTA_CODE && a.fuse_delta) {
          // delta = sum over the attended columns of P * dP (one extra sweep over this tile's single KV item)
          float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int cb = 0; cb < nch; ++cb) {
            uint32_t pv[8], dv[16];
            tmem_ld8(ts + cb * 8, pv);
            tmem_ld16(tl + C::kColP + cb * 16, dv);
            tmem_ld_wait();
            const int lo = c_lo - cb * 16, hi = c_hi - cb * 16;
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0, p1;
              unpack16<T>(pv[e >> 1], p0, p1);
              p0 = (e >= lo && e <= hi) ? p0 : 0.f;
              p1 = (e + 1 >= lo && e + 1 <= hi) ? p1 : 0.f;
              d4[e & 3] = fmaf(p0, __uint_as_float(dv[e]), d4[e & 3]);
              d4[(e + 1) & 3] = fmaf(p1, __uint_as_float(dv[e + 1]), d4[(e + 1) & 3]);
            }
          }
          delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
          if (i < a.N) {
            const int h = w.y * a.G + gr;
            const int64_t row = (static_cast<int64_t>(w.b) * a.Hq + h) * a.N + i;
            a.delta_out[row] = delta;
            if (a.dsrow != nullptr)      // pre_cur = lse of this row; ds_aux[h] = -sum exp(s_aux - lse) * delta  (:653-665)
              a.dsrow[row] = (pre_cur == -INFINITY) ? 0.f : -__expf(__ldg(a.s_aux + h) - pre_cur) * delta;
          }
        }
        // one 16-column chunk per trip, one (always masked) code path -- see the exp warps
#pragma unroll 1
        for (int cb = 0; cb < nch; ++cb) {
          uint32_t pv[8], dv[16], pk[8];
          const int c0 = cb * 16;
          tmem_ld8(ts + cb * 8, pv);
          tmem_ld16(tl + C::kColP + c0, dv);
          tmem_ld_wait();
          const int lo = c_lo - c0, hi = c_hi - c0;       // attended elements of this chunk: [lo, hi]
          if (__all_sync(0xffffffffu, lo <= 0 && hi >= 15)) {     // interior chunk of the band: no mask
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0, p1;
              unpack16<T>(pv[e >> 1], p0, p1);
              pk[e >> 1] = pack16_fast<T>(p0 * (__uint_as_float(dv[e]) - delta), p1 * (__uint_as_float(dv[e + 1]) - delta));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0, p1;
              unpack16<T>(pv[e >> 1], p0, p1);
              p0 = (e >= lo && e <= hi) ? p0 : 0.f;         // dP and delta are finite: P = 0 => dS = 0
              p1 = (e + 1 >= lo && e + 1 <= hi) ? p1 : 0.f;
              pk[e >> 1] = pack16_fast<T>(p0 * (__uint_as_float(dv[e]) - delta), p1 * (__uint_as_float(dv[e + 1]) - delta));
            }
          }
          tmem_st8(ts + C::kBNMax / 2 + (c0 >> 1), pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full + sb);
        if (threadIdx.x == 128) trace_ev(a.trace, 5, mtc, 3, w.n);
      }
    } else {
      // ---------------------------------------------------------------- epilogue warps
      const int ro = a.dq_swap ? (pr * a.G + gr) : (gr * a.P + pr);   // row in dQ's box order
      const int et = threadIdx.x - 256;                                 // 0..127
      ItemWalk w(a);
      int mtc = 0;
      while (w.next()) {
        if (!w.last_of_tile()) continue;
        if (et == 0) trace_ev(a.trace, 6, mtc, 1, w.n);
        if (et == 0) tma_store_wait_read0();     // previous store has finished reading the staging buffer
        named_bar_sync(2, 128);
        mbar_wait(dq_done, w.it & 1);
        tc_fence_after();
        if (et == 0) trace_ev(a.trace, 6, mtc, 2, w.n);
        uint32_t v[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld16(tl + C::kColQ + cc * 16, v[cc]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_free);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t pk[8];
#pragma unroll
          for (int e2 = 0; e2 < 16; e2 += 2)
            pk[e2 >> 1] = pack16<T>(__uint_as_float(v[cc][e2]) * a.scale, __uint_as_float(v[cc][e2 + 1]) * a.scale);
          *reinterpret_cast<uint4*>(stage_s + sw128_off(ro, cc * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(stage_s + sw128_off(ro, cc * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
          tma_tile_store(&tmdQ, stage_s, a.dq_swap, 0, w.q0, w.y * a.G, w.b);
          tma_store_commit();
          trace_ev(a.trace, 6, mtc, 3, w.n);
        }
      }
      if (et == 0) tma_store_wait_all0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, C::kTmemCols);
}

// ================================================================================== dK/dV kernel
// One CTA per (128-key tile, kv head, batch).  Keys sit on the TMEM lanes; the packed Q chunks that can see the
// tile stream through a Q ring (3 tiles) and a dO ring (2 tiles).  TMEM is full at head_dim 128 (S^T, dP^T, dK, dV:
// 4 x 128 columns), so S^T and dP^T are single-buffered and the overlap comes from the ORDER of the tensor pipe:
//
//   tensor pipe:  dV(c) | S(c+1) | dK(c) | dP(c+1) | dV(c+1) | S(c+2) | ...
//   math warps:           [ dS(c) ]  [   P(c+1)   ]  [ dS(c+1) ]  [  P(c+2)  ] ...
//
// P(c+1) (the MUFU-bound pass) runs under dK(c) + dP(c+1), dS(c+1) under dV(c+1) + S(c+2).  P^T overwrites the
// consumed S^T columns as 16-bit, dS^T the dP^T columns; UMMAs of one thread execute in issue order, so S(c+1) may
// be issued right behind the dV(c) that still reads P(c) from the same columns.  (Round 1's version ran
// S, dP -> math -> dV, dK strictly one after the other: 21 % of the tensor peak at BASELINE configs[2].)
template <int D> struct DkvCfg {
  static constexpr int kDS = D / 64;
  static constexpr int kBK = 128;                       // keys per CTA
  static constexpr int kQStages = 3;                    // Q(c) lives from S(c) to dK(c): two tiles in use + one in flight
  static constexpr int kDoStages = 2;                   // dO(c) lives from dP(c) to dV(c)
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBK * D * 2;
  static constexpr int kSlabQ = 128 * 128;
  static constexpr int kSlabKV = kBK * 128;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS = 0;                  // S^T (fp32) -> P^T (16-bit)
  static constexpr uint32_t kColP = 128;                // dP^T (fp32) -> dS^T (16-bit)
  static constexpr uint32_t kColK = 256;                // dK accumulator [keys][D]
  static constexpr uint32_t kColV = 256 + D;            // dV accumulator
  // no alignment slack: the dynamic shared memory is declared 1024-byte aligned (checked at kernel entry)
  static constexpr int kSmem = 2 * kKVBytes + (kQStages + kDoStages) * kQBytes + 2 * 2 * 128 * 4 + 256;
  static_assert(256 + 2 * D <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

struct DkvArgs {
  long long* trace;
  int B, N, S, W, Hq, Hkv, G, P, groups_per_kv;
  int ntiles;    // key tiles per (kv head, batch); the launch is 1-D: tiles holding sink keys first (they see every row)
  // Key tiles that hold sink tokens are visited by EVERY later row of the sequence: one CTA walking all N / P
  // position blocks outlasts the rest of the launch when batch x kv heads is small (4 sink tokens cost 0.9 ms at
  // B=1 N=8192 Hkv=8 D=64).  Their position-block range is cut into n_split CTAs that leave fp32 partials
  // [tile][split][key][dK | dV][D]; dkdv_split_reduce_kernel sums them in a fixed order (deterministic).
  int split_tiles, n_split;
  float* split_part;
  int dbg_delay;       // test knob (sfa_set_debug 0): half of the math warps sleep before their passes
  int q_swap, k_swap, v_swap;
  int fmt;
  float sl2, scale;
  const float* lse;
  const float* delta;
  void* dk;
  void* dv;
  Strides4 sdk, sdv;
  const int* seq_hi;   // packed sequences: one past the last query position that may attend key j (nullptr: none)
  int64_t seq_bs;
  int Dl;        // logical head_dim (<= the kernel's D): channels Dl .. D-1 are TMA zero fill and are not stored
};

// position blocks [pb_lo, pb_hi] (P positions each) whose queries can attend a key in [j0, j0+BK)
__device__ __forceinline__ void chunk_range(const DkvArgs& a, int b, int j0, int bk, int& pb_lo, int& pb_hi) {
  const int j1 = min(j0 + bk, a.N) - 1;                 // last key of the tile
  int i_max = -1;
  if (j0 < a.S) i_max = a.N - 1;                        // sink keys are seen by every later query
  if (a.W > 0) i_max = max(i_max, min(a.N - 1, j1 + a.W - 1));
  if (a.seq_hi != nullptr) i_max = min(i_max, __ldg(a.seq_hi + b * a.seq_bs + j1) - 1);   // no row beyond the last key's sequence
  pb_lo = j0 / a.P;
  pb_hi = (i_max >= j0) ? i_max / a.P : pb_lo - 1;
}

// 1-D launch of the dK/dV kernels: CTA id -> (key tile, kv head * batch index, split of a sink-holding tile).  The
// split CTAs come first: they are the longest.
__device__ __forceinline__ void dkv_block(const DkvArgs& a, int& tile_x, int& kvh, int& b, int& split) {
  const int id = blockIdx.x;
  const int heavy = a.split_tiles * a.n_split * a.Hkv * a.B;
  int yz;
  if (id < heavy) {
    int u = id;
    split = u % a.n_split;
    u /= a.n_split;
    tile_x = u % a.split_tiles;
    yz = u / a.split_tiles;
  } else {
    const int r = id - heavy, rest = a.ntiles - a.split_tiles;
    split = 0;
    tile_x = a.split_tiles + r % rest;
    yz = r / rest;
  }
  kvh = yz % a.Hkv;
  b = yz / a.Hkv;
}
// this CTA's share of the position blocks [pb_lo, pb_hi] of a split tile
__device__ __forceinline__ void dkv_split_range(const DkvArgs& a, int tile_x, int split, int& pb_lo, int& pb_hi) {
  if (a.n_split > 1 && tile_x < a.split_tiles) {
    const int npb_all = max(pb_hi - pb_lo + 1, 0);
    const int per = (npb_all + a.n_split - 1) / a.n_split;
    pb_lo += split * per;
    pb_hi = min(pb_hi, pb_lo + per - 1);
  }
}
// fp32 partial rows of a split CTA: [(b, kvh)][tile][split][key 0..127][dK (D) | dV (D)]
__device__ __forceinline__ float* dkv_part_row(const DkvArgs& a, int D, int tile_x, int kvh, int b, int split, int kr) {
  const size_t slot = ((static_cast<size_t>(b) * a.Hkv + kvh) * a.split_tiles + tile_x) * a.n_split + split;
  return a.split_part + (slot * 128 + kr) * static_cast<size_t>(2 * D);
}

// sums the splits of the sink-holding key tiles in a fixed order and writes the 16-bit dK / dV rows
template <typename T>
__global__ void dkdv_split_reduce_kernel(const DkvArgs a, int D) {
  const int tile_x = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int kr = threadIdx.x >> 1, hf = threadIdx.x & 1;      // two threads per key row: dK | dV
  const int j = tile_x * 128 + kr;
  if (j >= a.N) return;
  T* out = static_cast<T*>(hf ? a.dv : a.dk) + b * (hf ? a.sdv.b : a.sdk.b) + kvh * (hf ? a.sdv.h : a.sdk.h) +
           static_cast<int64_t>(j) * (hf ? a.sdv.n : a.sdk.n);
  for (int c0 = 0; c0 < a.Dl; c0 += 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < a.n_split; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(dkv_part_row(a, D, tile_x, kvh, b, s, kr) + hf * D + c0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[c0] = from_f<T>(acc.x);
    out[c0 + 1] = from_f<T>(acc.y);
    out[c0 + 2] = from_f<T>(acc.z);
    out[c0 + 3] = from_f<T>(acc.w);
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(kThreads, 1) dkdv_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                           const __grid_constant__ CUtensorMap tmdO,
                                                           const __grid_constant__ CUtensorMap tmK,
                                                           const __grid_constant__ CUtensorMap tmV, const DkvArgs a) {
  using C = DkvCfg<D>;
  extern __shared__ __align__(1024) unsigned char smem_aligned[];
  unsigned char* smem = smem_aligned;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();               // SWIZZLE_128B tiles need 1024-byte aligned slabs
  unsigned char* k_s = smem;
  unsigned char* v_s = k_s + C::kKVBytes;
  unsigned char* q_s = v_s + C::kKVBytes;                      // [kQStages][kQBytes]
  unsigned char* do_s = q_s + C::kQStages * C::kQBytes;        // [kDoStages][kQBytes]
  float* row_l2 = reinterpret_cast<float*>(do_s + C::kDoStages * C::kQBytes);   // [2][128]  -lse*log2e per chunk row
  float* row_dl = row_l2 + 2 * 128;                                             // [2][128]  -delta per chunk row
  uint64_t* bars = reinterpret_cast<uint64_t*>(row_dl + 2 * 128);
  uint64_t* kv_full = bars;
  uint64_t* q_full = kv_full + 1;              // [3]
  uint64_t* q_empty = q_full + C::kQStages;    // [3]  dK(c) complete
  uint64_t* do_full = q_empty + C::kQStages;   // [2]
  uint64_t* do_empty = do_full + C::kDoStages; // [2]  dV(c) complete
  uint64_t* s_full = do_empty + C::kDoStages;  // S^T(c) complete            (issuer -> math)
  uint64_t* dp_full = s_full + 1;              // dP^T(c) complete           (issuer -> math)
  uint64_t* p_ready = dp_full + 1;             // P^T(c) written over S^T    (math -> issuer)
  uint64_t* ds_ready = p_ready + 1;            // dS^T(c) written over dP^T  (math -> issuer)
  uint64_t* acc_done = ds_ready + 1;
  uint64_t* p_half = acc_done + 1;             // first two 16-row groups of both column halves of P^T(c) written: the dV
                                               // k-steps over those rows start under the rest of the exp pass
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile_x, kvh, b, split;
  dkv_block(a, tile_x, kvh, b, split);
  const int j0 = tile_x * C::kBK;
  int pb_lo, pb_hi;
  chunk_range(a, b, j0, C::kBK, pb_lo, pb_hi);
  dkv_split_range(a, tile_x, split, pb_lo, pb_hi);
  const bool to_part = (a.n_split > 1 && tile_x < a.split_tiles);
  const int npb = max(pb_hi - pb_lo + 1, 0);
  const int nchunks = npb * a.groups_per_kv;             // chunk c -> (group c % gpk, position block pb_lo + c / gpk)

  if (warp == kMathWarps && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < C::kQStages; ++s) {
      mbar_init(q_full + s, 1);
      mbar_init(q_empty + s, 1);
    }
    for (int s = 0; s < C::kDoStages; ++s) {
      mbar_init(do_full + s, 1);
      mbar_init(do_empty + s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, kMathThreads);
    mbar_init(p_half, kMathThreads);
    mbar_init(ds_ready, kMathThreads);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == kMathWarps + 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kMathWarps) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * C::kKVBytes);
      for (int s = 0; s < C::kDS; ++s) {
        tma_tile(k_s + s * C::kSlabKV, &tmK, kv_full, a.k_swap, s * 64, j0, kvh, b);
        tma_tile(v_s + s * C::kSlabKV, &tmV, kv_full, a.v_swap, s * 64, j0, kvh, b);
      }
      int grp = 0, pb = pb_lo, qs = 0, qph = 1, ds = 0, dph = 1;
      for (int c = 0; c < nchunks; ++c) {
        const int hq0 = (kvh * a.groups_per_kv + grp) * a.G;
        mbar_wait(q_empty + qs, qph);
        mbar_expect_tx(q_full + qs, C::kQBytes);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(q_s + qs * C::kQBytes + s * C::kSlabQ, &tmQ, q_full + qs, a.q_swap, s * 64, pb * a.P, hq0, b);
        mbar_wait(do_empty + ds, dph);
        mbar_expect_tx(do_full + ds, C::kQBytes);
        for (int s = 0; s < C::kDS; ++s)
          tma_tile(do_s + ds * C::kQBytes + s * C::kSlabQ, &tmdO, do_full + ds, a.q_swap, s * 64, pb * a.P, hq0, b);
        if (++grp == a.groups_per_kv) { grp = 0; ++pb; }
        if (++qs == C::kQStages) { qs = 0; qph ^= 1; }
        if (++ds == C::kDoStages) { ds = 0; dph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == kMathWarps + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && nchunks > 0) {
      const uint32_t idesc_s = make_idesc(a.fmt, 128, 128, 0, 0);    // S^T = K Q^T : M = keys, N = chunk rows
      const uint32_t idesc_acc = make_idesc(a.fmt, 128, D, 0, 1);    // dV += P^T dO : B = dO tile, MN-major
      const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s);
      auto issue_s = [&](uint32_t qa) {
#pragma unroll
        for (int s = 0; s < C::kDS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + C::kColS, make_sdesc(ka + s * C::kSlabKV + kk * 32, 16, 1024),
                    make_sdesc(qa + s * C::kSlabQ + kk * 32, 16, 1024), idesc_s, (s | kk) != 0);
      };
      auto issue_dp = [&](uint32_t doa) {
#pragma unroll
        for (int s = 0; s < C::kDS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + C::kColP, make_sdesc(va + s * C::kSlabKV + kk * 32, 16, 1024),
                    make_sdesc(doa + s * C::kSlabQ + kk * 32, 16, 1024), idesc_s, (s | kk) != 0);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(smem_u32(q_s));
      umma_commit(s_full);
      mbar_wait(do_full, 0);
      tc_fence_after();
      issue_dp(smem_u32(do_s));
      umma_commit(dp_full);
      int qs = 0, qph = 0, ds = 0, dph = 0;       // stage / phase of chunk c
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t qa = smem_u32(q_s + qs * C::kQBytes), doa = smem_u32(do_s + ds * C::kQBytes);
        int qs1 = qs + 1, qph1 = qph, ds1 = ds + 1, dph1 = dph;
        if (qs1 == C::kQStages) { qs1 = 0; qph1 ^= 1; }
        if (ds1 == C::kDoStages) { ds1 = 0; dph1 ^= 1; }
        const bool more = (c + 1 < nchunks);
        // dV += P(c)^T dO(c): contraction over the 128 chunk rows, 16 per UMMA; rows [0,64) were packed into 32-bit
        // columns [0,32) and rows [64,128) into [64,96) of each region (one range per math-warp half)
        // (k-steps 0, 1 and 4, 5 -- the first two row groups of each math-warp half -- as soon as they are written)
        mbar_wait(p_half, c & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if ((kk & 3) < 2)
            umma_ts(tmem + C::kColV, tmem + C::kColS + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8),
                    make_sdesc(doa + kk * 2048, C::kSlabQ, 1024), idesc_acc, (c > 0 || kk > 0));
        mbar_wait(p_ready, c & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if ((kk & 3) >= 2)
            umma_ts(tmem + C::kColV, tmem + C::kColS + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8),
                    make_sdesc(doa + kk * 2048, C::kSlabQ, 1024), idesc_acc, 1);
        umma_commit(do_empty + ds);
        if (more) {            // S(c+1) behind dV(c) in the pipe: it overwrites the columns P(c) is read from
          mbar_wait(q_full + qs1, qph1);
          tc_fence_after();
          issue_s(smem_u32(q_s + qs1 * C::kQBytes));
          umma_commit(s_full);
        }
        mbar_wait(ds_ready, c & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tmem + C::kColK, tmem + C::kColP + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8),
                  make_sdesc(qa + kk * 2048, C::kSlabQ, 1024), idesc_acc, (c > 0 || kk > 0));
        umma_commit(q_empty + qs);
        if (more) {
          mbar_wait(do_full + ds1, dph1);
          tc_fence_after();
          issue_dp(smem_u32(do_s + ds1 * C::kQBytes));
          umma_commit(dp_full);
        }
        qs = qs1; qph = qph1; ds = ds1; dph = dph1;
      }
      umma_commit(acc_done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ element-wise math + epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const int kr = quarter * 32 + lane;                 // key row == TMEM lane
    const int j = j0 + kr;
    // packed sequences: rows at or beyond e_j belong to a later sequence than key j
    const int e_j = (a.seq_hi != nullptr && j < a.N) ? __ldg(a.seq_hi + b * a.seq_bs + j) : 0x7fffffff;
    const int jw_lo = j0 + quarter * 32, jw_hi = jw_lo + 31;          // this warp's keys
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int tid = threadIdx.x;                        // 0..255
    const int sh_p = 31 - __clz(a.P), sh_g = 31 - __clz(a.G);          // P and G are powers of two
    const uint64_t sl2_2 = pack_f32x2(a.sl2, a.sl2), one2 = pack_f32x2(1.f, 1.f);
    // -lse*log2e and -delta of the 128 chunk rows (row order = the Q tile's box order), staged one chunk ahead
    const int st_pr = a.q_swap ? (tid >> sh_g) : (tid & (a.P - 1));
    const int st_gr = a.q_swap ? (tid & (a.G - 1)) : (tid >> sh_p);
    // raw lse / delta of this thread's chunk row (threads 0-127), loaded one chunk ahead; the values are only touched
    // when they are stored to shared memory at the end of the chunk (no scoreboard wait in between)
    auto load_stats = [&](int grp, int pb, float& l, float& d) {
      l = -INFINITY;
      d = 0.f;
      const int i = pb * a.P + st_pr;
      if (tid < 128 && i < a.N) {
        const int64_t row = (static_cast<int64_t>(b) * a.Hq + (kvh * a.groups_per_kv + grp) * a.G + st_gr) * a.N + i;
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(l) : "l"(a.lse + row));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(d) : "l"(a.delta + row));
      }
    };
    auto store_stats = [&](int buf, float l, float d) {
      if (tid < 128) {
        row_l2[buf * 128 + tid] = (l == -INFINITY) ? -INFINITY : -l * kLog2e;     // lse = -inf: nothing attended, P = 0
        row_dl[buf * 128 + tid] = -d;
      }
    };
    int grp = 0, pb = pb_lo;
    if (nchunks > 0) {
      float l0, d0;
      load_stats(grp, pb, l0, d0);
      store_stats(0, l0, d0);
      named_bar_sync(1, kMathThreads);
    }
    for (int c = 0; c < nchunks; ++c) {
      const int q0 = pb * a.P;
      int grp_n = grp + 1, pb_n = pb;
      if (grp_n == a.groups_per_kv) { grp_n = 0; ++pb_n; }
      float l_n = -INFINITY, d_n = 0.f;
      if (c + 1 < nchunks) load_stats(grp_n, pb_n, l_n, d_n);         // in flight during the two passes
      const float* rl = row_l2 + (c & 1) * 128;
      const float* rd = row_dl + (c & 1) * 128;
      // this thread: columns [half*64, half*64+64) of its key row; 16-bit results go to the low half of the
      // 32-bit columns it has already consumed: P^T over S^T, dS^T over dP^T
      uint32_t pp[32];
      // ---- pass 1: P^T = exp2(S^T * c - lse), masked
      mbar_wait(s_full, c & 1);
      tc_fence_after();
      if (a.dbg_delay && half == 1) __nanosleep(a.dbg_delay);
      // every row of the chunk attends every key of the tile (all but the few chunks at the two ends of the band):
      // CTA-uniform, so the common chunk pays for no per-group position arithmetic at all
      const bool interior = (j0 + C::kBK - 1 <= q0) && (q0 + a.P - 1 < a.N) && (a.seq_hi == nullptr) &&
                            ((j0 + C::kBK - 1 < a.S) || (a.W > 0 && j0 >= q0 + a.P - a.W));
      auto pass1 = [&](auto tag) {
        constexpr bool kInterior = decltype(tag)::value;
        uint32_t sv[2][16];
        tmem_ld16(tl + C::kColS + half * 64, sv[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c0 = half * 64 + g * 16;
          tmem_ld_wait();
          if (g < 3) tmem_ld16(tl + C::kColS + c0 + 16, sv[(g + 1) & 1]);
          const uint32_t (&x)[16] = sv[g & 1];
          bool any = true, all = true;
          if constexpr (!kInterior) {
            // positions covered by these 16 chunk rows
            int i_lo, i_hi;
            if (a.q_swap) {
              i_lo = q0 + (c0 >> sh_g);
              i_hi = q0 + ((c0 + 15) >> sh_g);
            } else {
              i_lo = q0 + (c0 & (a.P - 1));
              i_hi = (a.P >= 16) ? i_lo + 15 : q0 + a.P - 1;
              if (a.P < 16) i_lo = q0;
            }
            // warp-uniform: no key of this warp is attended by any of these rows / every key by every row
            any = (jw_lo <= i_hi) && ((jw_lo < a.S) || (jw_hi >= i_lo - a.W + 1)) && (i_lo < a.N);
            all = (jw_hi <= i_lo) && ((jw_hi < a.S) || (a.W > 0 && jw_lo >= i_hi - a.W + 1)) && (i_hi < a.N) &&
                  (a.seq_hi == nullptr);
          }
          uint32_t pk[8];
          if (all) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const uint64_t l2 = *reinterpret_cast<const uint64_t*>(rl + c0 + e);
              float x0, x1;
              unpack_f32x2(fma_f32x2(x[e], x[e + 1], sl2_2, l2), x0, x1);
              pk[e >> 1] = pack16_fast<T>(fast_exp2(x0), fast_exp2(x1));
            }
          } else if (any) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int r0 = c0 + e, r1 = r0 + 1;
              const int i0 = q0 + (a.q_swap ? (r0 >> sh_g) : (r0 & (a.P - 1)));
              const int i1 = q0 + (a.q_swap ? (r1 >> sh_g) : (r1 & (a.P - 1)));
              const float2 l2 = *reinterpret_cast<const float2*>(rl + r0);
              float p0 = fast_exp2(fmaf(__uint_as_float(x[e]), a.sl2, l2.x));
              float p1 = fast_exp2(fmaf(__uint_as_float(x[e + 1]), a.sl2, l2.y));
              const bool ok0 = attended(i0, j, a.S, a.W) && (j < a.N) && (i0 < e_j);
              const bool ok1 = attended(i1, j, a.S, a.W) && (j < a.N) && (i1 < e_j);
              pk[e >> 1] = pack16_fast<T>(ok0 ? p0 : 0.f, ok1 ? p1 : 0.f);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) pk[e] = 0u;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) pp[g * 8 + e] = pk[e];
          // each half packs into the front of its OWN source range (32-bit columns [half*64, half*64+32)),
          // so it never overwrites fp32 columns the other half (or this thread's next load) has yet to read
          tmem_st8(tl + C::kColS + half * 64 + g * 8, pk);
          if (g == 1) {            // row groups 0 and 1 of this half are in TMEM
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_half);
          }
        }
      };
      if (interior) pass1(std::true_type{});
      else pass1(std::false_type{});
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- pass 2: dS^T = P^T o (dP^T - delta) from the packed P pair (masked P is exactly 0 and dP is finite there:
      // dS = 0)
      mbar_wait(dp_full, c & 1);
      tc_fence_after();
      if (a.dbg_delay && half == 0 && (quarter & 1)) __nanosleep(a.dbg_delay);
      {
        uint32_t dv[2][16];
        tmem_ld16(tl + C::kColP + half * 64, dv[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c0 = half * 64 + g * 16;
          tmem_ld_wait();
          if (g < 3) tmem_ld16(tl + C::kColP + c0 + 16, dv[(g + 1) & 1]);
          const uint32_t (&x)[16] = dv[g & 1];
          uint32_t pd[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const uint64_t nd2 = *reinterpret_cast<const uint64_t*>(rd + c0 + e);
            float t0, t1;
            unpack_f32x2(fma_f32x2(x[e], x[e + 1], one2, nd2), t0, t1);
            pd[e >> 1] = ds_pair<T>(pp[g * 8 + (e >> 1)], t0, t1);
          }
          tmem_st8(tl + C::kColP + half * 64 + g * 8, pd);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_ready);
      // next chunk's row statistics: buffer (c + 1) & 1 was last read in chunk c - 1
      if (c + 1 < nchunks) {
        store_stats((c + 1) & 1, l_n, d_n);
        named_bar_sync(1, kMathThreads);
      }
      grp = grp_n;
      pb = pb_n;
    }
    // ---------------- epilogue: dK * scale, dV -> 16-bit -> global (one key row per thread, half the channels)
    if (nchunks > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
    }
    T* dkr = static_cast<T*>(a.dk) + b * a.sdk.b + kvh * a.sdk.h + static_cast<int64_t>(j) * a.sdk.n;
    T* dvr = static_cast<T*>(a.dv) + b * a.sdv.b + kvh * a.sdv.h + static_cast<int64_t>(j) * a.sdv.n;
#pragma unroll
    for (int cc = 0; cc < D / 2; cc += 16) {
      const int c0 = half * (D / 2) + cc;
      uint32_t kv_[16], vv_[16], pk[8], pv[8];
      if (nchunks > 0) {                                  // uniform: the tcgen05.ld stay warp-convergent
        tmem_ld16(tl + C::kColK + c0, kv_);
        tmem_ld16(tl + C::kColV + c0, vv_);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) kv_[e] = vv_[e] = 0u;
      }
      if (to_part) {                       // a split of a sink-holding tile: fp32 partial rows, summed by the reduce kernel
        float* pr_ = dkv_part_row(a, D, tile_x, kvh, b, split, kr);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          *reinterpret_cast<float4*>(pr_ + c0 + e) =
              make_float4(__uint_as_float(kv_[e]) * a.scale, __uint_as_float(kv_[e + 1]) * a.scale,
                          __uint_as_float(kv_[e + 2]) * a.scale, __uint_as_float(kv_[e + 3]) * a.scale);
          *reinterpret_cast<float4*>(pr_ + D + c0 + e) = make_float4(__uint_as_float(vv_[e]), __uint_as_float(vv_[e + 1]),
                                                                       __uint_as_float(vv_[e + 2]), __uint_as_float(vv_[e + 3]));
        }
        continue;
      }
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        pk[e >> 1] = pack16<T>(__uint_as_float(kv_[e]) * a.scale, __uint_as_float(kv_[e + 1]) * a.scale);
        pv[e >> 1] = pack16<T>(__uint_as_float(vv_[e]), __uint_as_float(vv_[e + 1]));
      }
      if (j < a.N) {
        if (c0 < a.Dl) {                 // head dims between 64 and 128 (80, 96, ...) run with zero-padded channels
          *reinterpret_cast<uint4*>(dkr + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(dvr + c0) = make_uint4(pv[0], pv[1], pv[2], pv[3]);
        }
        if (c0 + 8 < a.Dl) {
          *reinterpret_cast<uint4*>(dkr + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          *reinterpret_cast<uint4*>(dvr + c0 + 8) = make_uint4(pv[4], pv[5], pv[6], pv[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMathWarps + 1) tmem_dealloc(tmem, C::kTmemCols);
}

// ================================================================================== dK/dV kernel, head_dim 64
// Same schedule idea as dq64_kernel, KV-stationary: S^T two chunks ahead into a double-buffered
// region, dP^T one chunk ahead into a single region, two-phase math (exp first), and BOTH 16-bit
// operands of the accumulating UMMAs (P^T for dV, dS^T for dK) packed into the consumed S^T
// buffer, which frees the dP^T region as soon as the math has read it.  Two issuer warps share the
// UMMA work (a tcgen05.mma costs its issuing thread ~80 cycles): B issues S^T / dP^T, A issues dV / dK.
struct Dkv64Cfg {
  static constexpr int D = 64;
  static constexpr int kBK = 128;
  static constexpr int kQStages = 4;      // Q(c): S^T(c) is issued at chunk c-2, dK(c) at chunk c
  static constexpr int kOStages = 3;      // dO(c): dP^T(c) at chunk c-1, dV(c) at chunk c
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBK * D * 2;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS = 0;      // S^T buffers at 0 and 128
  static constexpr uint32_t kColP = 256;    // dP^T
  static constexpr uint32_t kColK = 384;    // dK accumulator [keys][64]
  static constexpr uint32_t kColV = 448;    // dV accumulator
  static constexpr int kMathW = 16;                      // 4 per scheduler: each thread owns one key row x 32 chunk rows
  static constexpr int kMathT = kMathW * 32;
  static constexpr int kThreads = kMathT + 96;           // + TMA producer, issuer B, issuer A
  static constexpr int kSmem = 1024 + 2 * kKVBytes + (kQStages + kOStages) * kQBytes + 2 * 2 * 128 * 4 + 512;
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

template <typename T>
__global__ void __launch_bounds__(Dkv64Cfg::kThreads, 1) dkdv64_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmdO,
                                                                       const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmV,
                                                                       const DkvArgs a) {
  using C = Dkv64Cfg;
  constexpr int D = C::D;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* k_s = smem;
  unsigned char* v_s = k_s + C::kKVBytes;
  unsigned char* q_s = v_s + C::kKVBytes;                      // [kQStages][kQBytes]
  unsigned char* do_s = q_s + C::kQStages * C::kQBytes;        // [kOStages][kQBytes]
  float* row_l2 = reinterpret_cast<float*>(do_s + C::kOStages * C::kQBytes);   // [2][128]  -lse*log2e per chunk row
  float* row_dl = row_l2 + 2 * 128;                                            // [2][128]  delta per chunk row
  uint64_t* bars = reinterpret_cast<uint64_t*>(row_dl + 2 * 128);
  uint64_t* kv_full = bars;
  uint64_t* q_full = kv_full + 1;              // [kQStages]
  uint64_t* do_full = q_full + C::kQStages;    // [kOStages]
  uint64_t* chunk_done = do_full + C::kOStages;   // [4]  dV(c), dK(c) complete: frees Q(c), dO(c) and S^T buffer c & 1
  uint64_t* s_full = chunk_done + 4;           // [2]
  uint64_t* dp_full = s_full + 2;
  uint64_t* p_full = dp_full + 1;              // [2]: both issuer warps wait on it, alternating barriers rule out a missed phase
  uint64_t* acc_done = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const trc = (SFA_TRACE && blockIdx.x == 20) ? a.trace : nullptr;
  auto tev = [&](int role, int& cnt, int code, int idx) {
    if (SFA_TRACE && trc != nullptr && cnt < 256) {
      trc[(role * 256 + cnt) * 2] = (static_cast<long long>(code) << 32) | static_cast<unsigned>(idx);
      trc[(role * 256 + cnt) * 2 + 1] = clock64();
      ++cnt;
    }
  };
  int tile_x, kvh, b, split;
  dkv_block(a, tile_x, kvh, b, split);
  const int j0 = tile_x * C::kBK;
  int pb_lo, pb_hi;
  chunk_range(a, b, j0, C::kBK, pb_lo, pb_hi);
  dkv_split_range(a, tile_x, split, pb_lo, pb_hi);
  const bool to_part = (a.n_split > 1 && tile_x < a.split_tiles);
  const int npb = max(pb_hi - pb_lo + 1, 0);
  const int gpk = a.groups_per_kv;
  const int nchunks = npb * gpk;        // chunk c -> (position block pb_lo + c / gpk, group c % gpk); walked with counters

  if (warp == C::kMathW && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < C::kQStages; ++s) mbar_init(q_full + s, 1);
    for (int s = 0; s < C::kOStages; ++s) mbar_init(do_full + s, 1);
    for (int s = 0; s < 4; ++s) mbar_init(chunk_done + s, 1);
    mbar_init(s_full, 1);
    mbar_init(s_full + 1, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_full, C::kMathT);
    mbar_init(p_full + 1, C::kMathT);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == C::kMathW + 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == C::kMathW) {
    // ------------------------------------------------------------------ TMA producer: Q(c+1) then dO(c), c = -1, 0, ...
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * C::kKVBytes);
      tma_tile(k_s, &tmK, kv_full, a.k_swap, 0, j0, kvh, b);
      tma_tile(v_s, &tmV, kv_full, a.v_swap, 0, j0, kvh, b);
      int pbq = pb_lo, gq = 0, pbo = pb_lo, go = 0;
      for (int c = -1; c < nchunks; ++c) {
        const int cq = c + 1;
        if (cq < nchunks) {
          const int qs = cq % C::kQStages;
          if (cq >= C::kQStages) mbar_wait(chunk_done + ((cq - C::kQStages) & 3), ((cq - C::kQStages) >> 2) & 1);
          mbar_expect_tx(q_full + qs, C::kQBytes);
          tma_tile(q_s + qs * C::kQBytes, &tmQ, q_full + qs, a.q_swap, 0, pbq * a.P, (kvh * gpk + gq) * a.G, b);
          if (++gq == gpk) { gq = 0; ++pbq; }
        }
        if (c >= 0) {
          const int os = c % C::kOStages;
          if (c >= C::kOStages) mbar_wait(chunk_done + ((c - C::kOStages) & 3), ((c - C::kOStages) >> 2) & 1);
          mbar_expect_tx(do_full + os, C::kQBytes);
          tma_tile(do_s + os * C::kQBytes, &tmdO, do_full + os, a.q_swap, 0, pbo * a.P, (kvh * gpk + go) * a.G, b);
          if (++go == gpk) { go = 0; ++pbo; }
        }
      }
    }
    __syncwarp();
  } else if (warp == C::kMathW + 1) {
    // ------------------------------------------------------------------ issuer B: S^T(c) = K Q(c)^T, dP^T(c) = V dO(c)^T
    if (lane == 0 && nchunks > 0) {
      const uint32_t idesc_s = make_idesc(a.fmt, 128, 128, 0, 0);
      const uint64_t kd = make_sdesc(smem_u32(k_s), 16, 1024), vd = make_sdesc(smem_u32(v_s), 16, 1024);
      int tc = 0;
      auto issue_s = [&](int c) {
        const int qs = c % C::kQStages;
        tev(1, tc, 1, c);
        mbar_wait(q_full + qs, (c / C::kQStages) & 1);
        tc_fence_after();
        tev(1, tc, 2, c);
        const uint64_t qd = make_sdesc(smem_u32(q_s + qs * C::kQBytes), 16, 1024);
        const uint32_t ts = tmem + C::kColS + (c & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(ts, kd + kk * 2, qd + kk * 2, idesc_s, kk != 0);
        umma_commit(s_full + (c & 1));
        tev(1, tc, 3, c);
      };
      auto issue_dp = [&](int c) {
        const int os = c % C::kOStages;
        tev(2, tc, 1, c);
        mbar_wait(do_full + os, (c / C::kOStages) & 1);
        tc_fence_after();
        tev(2, tc, 2, c);
        const uint64_t dod = make_sdesc(smem_u32(do_s + os * C::kQBytes), 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C::kColP, vd + kk * 2, dod + kk * 2, idesc_s, kk != 0);
        umma_commit(dp_full);
        tev(2, tc, 3, c);
      };
      mbar_wait(kv_full, 0);
      tc_fence_after();
      issue_s(0);
      if (nchunks > 1) issue_s(1);
      issue_dp(0);
      for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
          tev(2, tc, 4, c);
          mbar_wait(p_full + (c & 1), (c >> 1) & 1);   // the math has read dP^T(c)
          tc_fence_after();
          issue_dp(c + 1);
        }
        if (c + 2 < nchunks) {
          tev(1, tc, 4, c);
          mbar_wait(chunk_done + (c & 3), (c >> 2) & 1);   // dV(c), dK(c) have consumed S^T buffer c & 1
          tc_fence_after();
          issue_s(c + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp == C::kMathW + 2) {
    // ------------------------------------------------------------------ issuer A: dV += P^T dO, dK += dS^T Q
    if (lane == 0) {
      const uint32_t idesc_acc = make_idesc(a.fmt, 128, D, 0, 1);
      int tc = 0;
      for (int c = 0; c < nchunks; ++c) {
        tev(3, tc, 1, c);
        const int qs = c % C::kQStages, os = c % C::kOStages;
        const uint64_t qd = make_sdesc(smem_u32(q_s + qs * C::kQBytes), C::kQBytes, 1024);
        const uint64_t dod = make_sdesc(smem_u32(do_s + os * C::kQBytes), C::kQBytes, 1024);
        const uint32_t ts = tmem + C::kColS + (c & 1) * 128;
        mbar_wait(p_full + (c & 1), (c >> 1) & 1);
        tc_fence_after();
        tev(3, tc, 2, c);
        // packed operands inside the S^T buffer, per 32-row part p: P^T rows [32p, 32p+32) at columns 32p + [0,16),
        // dS^T rows of the same part at columns 32p + [16,32)  (8 columns per 16 rows)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tmem + C::kColV, ts + (kk >> 1) * 32 + (kk & 1) * 8, dod + kk * (2048 >> 4), idesc_acc, (c > 0 || kk > 0));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tmem + C::kColK, ts + (kk >> 1) * 32 + 16 + (kk & 1) * 8, qd + kk * (2048 >> 4), idesc_acc,
                  (c > 0 || kk > 0));
        umma_commit(chunk_done + (c & 3));
        tev(3, tc, 3, c);
      }
      umma_commit(acc_done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ element-wise math + epilogue
    const int quarter = warp & 3, part = warp >> 2;     // part: which 32 of the 128 chunk rows
    const int kr = quarter * 32 + lane;                 // key row == TMEM lane
    const int j = j0 + kr;
    const int jw_lo = j0 + quarter * 32, jw_hi = jw_lo + 31;          // this warp's keys
    // queries that attend key j: i in [j, i_hi]
    int i_hi = (j >= a.N) ? -1 : ((j < a.S) ? 0x7fffffff : ((a.W > 0) ? j + a.W - 1 : -1));
    if (a.seq_hi != nullptr && j < a.N) i_hi = min(i_hi, __ldg(a.seq_hi + b * a.seq_bs + j) - 1);   // packed sequences
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int tid = threadIdx.x;                        // 0..511
    const int sh_p = 31 - __clz(a.P), sh_g = 31 - __clz(a.G);          // P and G are powers of two

    // lse and delta of chunk row `tid` (threads 0..127): issued one chunk ahead and only consumed (converted to
    // -lse*log2e and stored to shared memory) at the end of the current chunk, so the load latency is hidden
    auto load_row = [&](int pb, int grp, float& l, float& dl) {
      l = INFINITY;                      // rows past N: P = exp2(s - inf) = 0
      dl = 0.f;
      if (tid < 128) {
        const int pr = a.q_swap ? (tid >> sh_g) : (tid & (a.P - 1));
        const int gr = a.q_swap ? (tid & (a.G - 1)) : (tid >> sh_p);
        const int i = pb * a.P + pr;
        if (i < a.N) {
          const int64_t row = (static_cast<int64_t>(b) * a.Hq + (kvh * gpk + grp) * a.G + gr) * a.N + i;
          asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(l) : "l"(a.lse + row));
          asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dl) : "l"(a.delta + row));
        }
      }
    };
    auto neg_l2_of = [](float l) { return (l == -INFINITY) ? -INFINITY : -l * kLog2e; };   // lse = -inf: nothing attended
    int pb = pb_lo, grp = 0, mtc = 0;
    float nl_n, dl_n;
    if (nchunks > 0) {
      load_row(pb, grp, nl_n, dl_n);
      if (tid < 128) {
        row_l2[tid] = neg_l2_of(nl_n);
        row_dl[tid] = dl_n;
      }
    }
    for (int c = 0; c < nchunks; ++c) {
      const int q0 = pb * a.P;
      int pbn = pb, gn = grp;
      if (++gn == gpk) { gn = 0; ++pbn; }
      if (tid == 0) tev(4, mtc, 1, c);
      named_bar_sync(1, C::kMathT);                     // rows of chunk c are in row_*[c & 1]
      if (tid == 0) tev(4, mtc, 5, c);
      if (c + 1 < nchunks) load_row(pbn, gn, nl_n, dl_n);   // in flight during this chunk's math
      const float* rl = row_l2 + (c & 1) * 128;
      const float* rd = row_dl + (c & 1) * 128;
      const uint32_t ts = tl + C::kColS + (c & 1) * 128;

      // per 16-column block: attended positions [i_lo, i_hi_b]; any / all lanes of the warp attended
      // Both phases are rolled loops over this thread's two 16-row blocks with ONE code path each, so the hot code
      // of the 16 math warps stays resident in the instruction cache (ncu: stall_no_inst dominated the unrolled
      // multi-variant version).  P^T goes to TMEM as 16-bit in phase 1 (its final place, already masked) and is
      // read back in phase 2, so nothing has to live in registers across the dP^T wait.
      // attended rows of key j inside block [c0, c0+16): bit e of the mask
      auto block_mask = [&](int c0, bool& any) -> uint32_t {
        int i_lo, i_up;
        if (a.q_swap) {
          i_lo = q0 + (c0 >> sh_g);
          i_up = q0 + ((c0 + 15) >> sh_g);
        } else if (a.P >= 16) {
          i_lo = q0 + (c0 & (a.P - 1));
          i_up = i_lo + 15;
        } else {
          i_lo = q0;
          i_up = q0 + a.P - 1;
        }
        // warp-uniform: does any key of this warp see any row of the block?
        any = (jw_lo <= i_up) && ((jw_lo < a.S) || (a.W > 0 && jw_hi >= i_lo - a.W + 1)) && (i_lo < a.N);
        if (!any) return 0u;
        if (!a.q_swap && a.P >= 16) {       // 16 consecutive positions: one run [lo, hi]
          const int lo = max(j - i_lo, 0), hi = min(i_hi - i_lo, 15);
          return (hi >= lo) ? ((0xffffu >> (15 - hi)) & (0xffffu << lo)) : 0u;
        }
        uint32_t m = 0;
#pragma unroll 1
        for (int e = 0; e < 16; ++e) {
          const int rr = c0 + e;
          const int i = q0 + (a.q_swap ? (rr >> sh_g) : (rr & (a.P - 1)));
          m |= ((i >= j) && (i <= i_hi)) ? (1u << e) : 0u;
        }
        return m;
      };
      mbar_wait(s_full + (c & 1), (c >> 1) & 1);
      tc_fence_after();
      if (tid == 0) tev(4, mtc, 2, c);
      // ---- phase 1: P^T = mask * exp2(S^T * c - lse[row]) -> 16-bit at S^T columns part*32 + [0,16)
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        const int c0 = part * 32 + jj * 16;
        bool any;
        const uint32_t mask = block_mask(c0, any);
        uint32_t pp[8];
        if (any) {
          uint32_t sv[16];
          tmem_ld16(ts + c0, sv);
          tmem_ld_wait();
          if (__all_sync(0xffffffffu, mask == 0xffffu)) {      // block fully inside the band for every key of the warp
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(rl + c0 + e);
              const float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, l4.x));
              const float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, l4.y));
              const float p2 = fast_exp2(fmaf(__uint_as_float(sv[e + 2]), a.sl2, l4.z));
              const float p3 = fast_exp2(fmaf(__uint_as_float(sv[e + 3]), a.sl2, l4.w));
              pp[e >> 1] = pack16_fast<T>(p0, p1);
              pp[(e >> 1) + 1] = pack16_fast<T>(p2, p3);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(rl + c0 + e);
              float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, l4.x));
              float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, l4.y));
              float p2 = fast_exp2(fmaf(__uint_as_float(sv[e + 2]), a.sl2, l4.z));
              float p3 = fast_exp2(fmaf(__uint_as_float(sv[e + 3]), a.sl2, l4.w));
              p0 = (mask & (1u << e)) ? p0 : 0.f;
              p1 = (mask & (2u << e)) ? p1 : 0.f;
              p2 = (mask & (4u << e)) ? p2 : 0.f;
              p3 = (mask & (8u << e)) ? p3 : 0.f;
              pp[e >> 1] = pack16_fast<T>(p0, p1);
              pp[(e >> 1) + 1] = pack16_fast<T>(p2, p3);
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) pp[e] = 0u;
        }
        tmem_st8(ts + part * 32 + jj * 8, pp);
      }
      tmem_st_wait();
      // ---- phase 2: dS^T = P^T * (dP^T - delta[row]) -> 16-bit at S^T columns part*32 + [16,32)
      if (tid == 0) tev(4, mtc, 6, c);
      mbar_wait(dp_full, c & 1);
      tc_fence_after();
      if (tid == 0) tev(4, mtc, 4, c);
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        const int c0 = part * 32 + jj * 16;
        bool any;
        (void)block_mask(c0, any);
        uint32_t pd[8];
        if (any) {
          uint32_t pq[8], dv[16];
          tmem_ld8(ts + part * 32 + jj * 8, pq);
          tmem_ld16(tl + C::kColP + c0, dv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(rd + c0 + e);
            float p0, p1, p2, p3;
            unpack16<T>(pq[e >> 1], p0, p1);
            unpack16<T>(pq[(e >> 1) + 1], p2, p3);
            pd[e >> 1] = pack16_fast<T>(p0 * (__uint_as_float(dv[e]) - d4.x), p1 * (__uint_as_float(dv[e + 1]) - d4.y));
            pd[(e >> 1) + 1] = pack16_fast<T>(p2 * (__uint_as_float(dv[e + 2]) - d4.z), p3 * (__uint_as_float(dv[e + 3]) - d4.w));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) pd[e] = 0u;
        }
        tmem_st8(ts + part * 32 + 16 + jj * 8, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + (c & 1));
      if (tid == 0) tev(4, mtc, 3, c);
      if (c + 1 < nchunks && tid < 128) {
        row_l2[((c + 1) & 1) * 128 + tid] = neg_l2_of(nl_n);
        row_dl[((c + 1) & 1) * 128 + tid] = dl_n;
      }
      pb = pbn;
      grp = gn;
    }
    // ---- epilogue: dK * scale, dV -> 16-bit -> global (one key row per thread, half the channels)
    if (nchunks > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
    }
    T* dkr = static_cast<T*>(a.dk) + b * a.sdk.b + kvh * a.sdk.h + static_cast<int64_t>(j) * a.sdk.n;
    T* dvr = static_cast<T*>(a.dv) + b * a.sdv.b + kvh * a.sdv.h + static_cast<int64_t>(j) * a.sdv.n;
    {
      const int c0 = part * 16;                           // 16 of the 64 channels per thread
      uint32_t kv_[16], vv_[16], pk[8], pv2[8];
      if (nchunks > 0) {                                  // uniform: the tcgen05.ld stay warp-convergent
        tmem_ld16(tl + C::kColK + c0, kv_);
        tmem_ld16(tl + C::kColV + c0, vv_);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) kv_[e] = vv_[e] = 0u;
      }
      if (to_part) {                       // a split of a sink-holding tile: fp32 partial rows, summed by the reduce kernel
        float* pr_ = dkv_part_row(a, D, tile_x, kvh, b, split, kr);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          *reinterpret_cast<float4*>(pr_ + c0 + e) =
              make_float4(__uint_as_float(kv_[e]) * a.scale, __uint_as_float(kv_[e + 1]) * a.scale,
                          __uint_as_float(kv_[e + 2]) * a.scale, __uint_as_float(kv_[e + 3]) * a.scale);
          *reinterpret_cast<float4*>(pr_ + D + c0 + e) = make_float4(__uint_as_float(vv_[e]), __uint_as_float(vv_[e + 1]),
                                                                       __uint_as_float(vv_[e + 2]), __uint_as_float(vv_[e + 3]));
        }
      }
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        pk[e >> 1] = pack16<T>(__uint_as_float(kv_[e]) * a.scale, __uint_as_float(kv_[e + 1]) * a.scale);
        pv2[e >> 1] = pack16<T>(__uint_as_float(vv_[e]), __uint_as_float(vv_[e + 1]));
      }
      if (j < a.N && !to_part) {
        *reinterpret_cast<uint4*>(dkr + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dkr + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        *reinterpret_cast<uint4*>(dvr + c0) = make_uint4(pv2[0], pv2[1], pv2[2], pv2[3]);
        *reinterpret_cast<uint4*>(dvr + c0 + 8) = make_uint4(pv2[4], pv2[5], pv2[6], pv2[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C::kMathW + 1) tmem_dealloc(tmem, C::kTmemCols);
}

}  // namespace
static long long* g_trace = nullptr;
void set_trace_buffer(long long* p) { g_trace = p; }
long long* trace_buffer() { return g_trace; }
namespace {

int sm_count() { return device_sm_count(); }

// Fused delta: possible when every packed tile has exactly one KV item (no sinks, the band of a tile fits one
// item).  OFF by default -- measured at the C1 shape it saves the 43 us preprocess pass but costs the dS warps
// (which sit on the dS -> dP -> dS critical chain) 35 us, and delta from 16-bit P puts ds_aux 1.4e-2 away from
// the fp32 path (the bar is 2e-3).  SFA_FUSE_DELTA=1 enables it for experiments.
inline bool fuses_delta(const AttnParams& p, int P, int BN) {
  static const bool enabled = SFA_DQ64_FUSE_DELTA_CODE && getenv("SFA_FUSE_DELTA") != nullptr;
  const int64_t span = (int64_t)(p.W < p.N ? p.W : p.N) + P - 1;
  return enabled && p.S == 0 && p.W > 0 && span <= BN;
}

// kWide: head_dim 64 on the D-generic kernels (dq_kernel / dkdv_kernel: items of 96 / chunks of 128 keys, built for long
// KV loops) instead of dq64_kernel / dkdv64_kernel (built around short bands)
#ifndef SFA_WIDE64_MIN_W
#define SFA_WIDE64_MIN_W 2048
#endif
template <int D, bool kWide> struct DqSel {
  static constexpr int kBNMax = DqCfg<D>::kBNMax, kSmem = DqCfg<D>::kSmem;
};
template <> struct DqSel<64, false> {
  static constexpr int kBNMax = Dq64Cfg::kBNMax, kSmem = Dq64Cfg::kSmem;
};

template <typename T, int D, bool kWide = false>
cudaError_t launch_bwd(const AttnParams& p, int dtype, int stages, cudaStream_t st) {
  const int group = p.Hq / p.Hkv;
  int G, P;
  pick_packing(p.Hq, p.Hkv, G, P);
  const int fmt = (dtype == SFA_DTYPE_BF16) ? 1 : 0;
  TileMap mq, mdo;
  if (!make_tile_map(&mq, p.q, dtype, p.D, p.N, p.Hq, p.B, p.sq, P, G)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mdo, p.dout, dtype, p.D, p.N, p.Hq, p.B, p.sdo, P, G)) return cudaErrorInvalidValue;
  if (mq.swap_nh != mdo.swap_nh) return cudaErrorInvalidValue;   // guarded by tc_bwd_supported

  if (stages & 2) {
    constexpr bool k64 = (D == 64) && !kWide;      // the head_dim-64 special kernels
    constexpr int kBNMax = DqSel<D, kWide>::kBNMax;
    constexpr int kSmemDq = DqSel<D, kWide>::kSmem;
    static std::atomic<unsigned long long> attr_done{0};
    {
      cudaError_t e;
      if constexpr (k64) e = ensure_dyn_smem(dq64_kernel<T>, kSmemDq, attr_done);
      else e = ensure_dyn_smem(dq_kernel<T, D>, kSmemDq, attr_done);
      if (e != cudaSuccess) return e;
    }
    const int BN = pick_bn(p.W, p.N, P, kBNMax);
    TileMap mk, mv, mdq;
    if (!make_tile_map(&mk, p.k, dtype, p.D, p.N, p.Hkv, p.B, p.sk, BN, 1)) return cudaErrorInvalidValue;
    if (!make_tile_map(&mv, p.v, dtype, p.D, p.N, p.Hkv, p.B, p.sv, BN, 1)) return cudaErrorInvalidValue;
    if (!make_tile_map(&mdq, p.dq, dtype, p.D, p.N, p.Hq, p.B, p.sdq, P, G)) return cudaErrorInvalidValue;
    BwdArgs a;
    a.B = p.B; a.N = p.N; a.S = p.S; a.W = p.W; a.Hq = p.Hq; a.G = G; a.P = P; a.BN = BN;
    a.groups_per_kv = group / G;
    a.ny = p.Hq / G;
    a.nblk = (p.N + P - 1) / P;
    a.total_tiles = a.nblk * a.ny * p.B;
    a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh; a.dq_swap = mdq.swap_nh;
    a.fmt = fmt;
    a.sl2 = p.scale * kLog2e;
    a.scale = p.scale;
    a.lse = p.lse;
    a.delta = p.delta;
    a.trace = trace_buffer();
    a.bn_mul = bn_magic(BN);
    a.q_off = 0; a.seq_lo = p.seq_lo; a.seq_bs = p.seq_bs;      // packed sequences (no chunk offset in these kernels)
    a.dbg_delay = debug_knob(0);
    a.fuse_delta = (k64 && fuses_delta(p, P, BN)) ? 1 : 0;
    a.delta_out = p.delta;
    a.dsrow = (p.s_aux != nullptr && p.ds_aux != nullptr) ? p.dsrow : nullptr;
    a.s_aux = p.s_aux;
    const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
    a.tiles_per_cta = 0;
    if constexpr (k64) dq64_kernel<T><<<grid, Dq64Cfg::kThreads, kSmemDq, st>>>(mq.map, mdo.map, mk.map, mv.map, mdq.map, a);
    else dq_kernel<T, D><<<grid, kThreads, kSmemDq, st>>>(mq.map, mdo.map, mk.map, mv.map, mdq.map, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (stages & 4) {
    constexpr int kBK = 128;
    constexpr bool k64 = (D == 64) && !kWide;
    constexpr int kSmemKv = k64 ? Dkv64Cfg::kSmem : DkvCfg<D>::kSmem;
    static std::atomic<unsigned long long> attr_done{0};
    {
      cudaError_t e;
      if constexpr (k64) e = ensure_dyn_smem(dkdv64_kernel<T>, kSmemKv, attr_done);
      else e = ensure_dyn_smem(dkdv_kernel<T, D>, kSmemKv, attr_done);
      if (e != cudaSuccess) return e;
    }
    TileMap mk, mv;
    if (!make_tile_map(&mk, p.k, dtype, p.D, p.N, p.Hkv, p.B, p.sk, kBK, 1)) return cudaErrorInvalidValue;
    if (!make_tile_map(&mv, p.v, dtype, p.D, p.N, p.Hkv, p.B, p.sv, kBK, 1)) return cudaErrorInvalidValue;
    DkvArgs a;
    a.B = p.B; a.N = p.N; a.S = p.S; a.W = p.W; a.Hq = p.Hq; a.Hkv = p.Hkv; a.G = G; a.P = P;
    a.groups_per_kv = group / G;
    a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh;
    a.fmt = fmt;
    a.sl2 = p.scale * kLog2e;
    a.scale = p.scale;
    a.lse = p.lse;
    a.delta = p.delta;
    a.dk = p.dk; a.dv = p.dv; a.sdk = p.sdk; a.sdv = p.sdv;
    a.Dl = p.D;
    a.seq_hi = p.seq_hi; a.seq_bs = p.seq_bs;
    a.dbg_delay = debug_knob(0);
    a.trace = trace_buffer();
    dim3 grid((p.N + kBK - 1) / kBK, p.Hkv, p.B);
    a.ntiles = static_cast<int>(grid.x);
    // split the position-block range of the tiles that hold sink keys (every later row visits them) over several CTAs
    a.split_tiles = 0;
    a.n_split = 1;
    a.split_part = p.kv_part;
    if (p.S > 0) {
      constexpr int Dk = k64 ? 64 : D;
      a.split_tiles = ((p.S < p.N ? p.S : p.N) + kBK - 1) / kBK;
      const int npb_all = (p.N + P - 1) / P;                                         // position blocks a sink tile sees
      const int npb_reg = ((p.W < p.N ? p.W : p.N) + kBK + P - 1) / P + 1;           // ... a window-only tile
      int want = (npb_all + npb_reg - 1) / npb_reg;
      const size_t slot_bytes = static_cast<size_t>(128) * 2 * Dk * sizeof(float);
      const size_t slots = (p.kv_part != nullptr) ? p.kv_part_bytes / slot_bytes : 0;
      const size_t tiles = static_cast<size_t>(a.split_tiles) * p.Hkv * p.B;
      const int cap = static_cast<int>(slots / (tiles ? tiles : 1));
      if (want > cap) want = cap;
      if (want > 64) want = 64;
      // only when the unsplit tile would be the tail of the launch: its position blocks against the blocks one SM gets
      // of everything else (split CTAs re-load K / V and add a reduce launch: 3 % slower where it is not needed)
      const double per_sm = static_cast<double>(a.ntiles) * p.Hkv * p.B * npb_reg / sm_count();
      if (npb_all < 0.75 * per_sm) want = 1;
      a.n_split = want >= 2 ? want : 1;
    }
    const int total_ctas = ((a.ntiles - a.split_tiles) + a.split_tiles * a.n_split) * p.Hkv * p.B;
    if constexpr (k64) dkdv64_kernel<T><<<total_ctas, Dkv64Cfg::kThreads, kSmemKv, st>>>(mq.map, mdo.map, mk.map, mv.map, a);
    else dkdv_kernel<T, D><<<total_ctas, kThreads, kSmemKv, st>>>(mq.map, mdo.map, mk.map, mv.map, a);   // 1-D: longest tiles first
    if (a.n_split > 1) {
      if (cudaError_t e2 = cudaGetLastError()) return e2;
      dkdv_split_reduce_kernel<T><<<dim3(a.split_tiles, p.Hkv, p.B), 256, 0, st>>>(a, k64 ? 64 : D);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace

bool tc_bwd_fuses_delta(const AttnParams& p, int dtype) {
  (void)dtype;
  if (p.D != 64) return false;
  int G, P;
  pick_packing(p.Hq, p.Hkv, G, P);
  return fuses_delta(p, P, pick_bn(p.W, p.N, P, Dq64Cfg::kBNMax));
}

bool tc_bwd_supported(const AttnParams& p, int dtype) {
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return false;
  // extended geometry: these kernels take packed sequences (without sink tokens); chunked prefill (q_off / a longer
  // key axis) is the fused kernel's or the CUDA-core path's
  if (p.q_off != 0 || p.Nkv != p.N) return false;
  if (p.seq_lo != nullptr && p.S > 0) return false;
  // head dims 72 .. 120 (multiples of 8: 80, 96, 112, ...) run on the head_dim-128 kernels: the TMA boxes cover 128
  // channels, the tensor only has D, so the tail is hardware zero fill on loads and clipped on stores
  if (p.D != 64 && !(p.D > 64 && p.D <= 128 && p.D % 8 == 0)) return false;
  if (p.N < 1) return false;
  if (p.S <= 0 && p.W <= 0) return false;     // nothing attended: the CUDA-core path writes the zeros
  if (!(tma_compatible(p.q, p.sq, p.B, p.Hq, p.N) && tma_compatible(p.k, p.sk, p.B, p.Hkv, p.N) && tma_compatible(p.v, p.sv, p.B, p.Hkv, p.N) &&
        tma_compatible(p.dout, p.sdo, p.B, p.Hq, p.N) && tma_compatible(p.dq, p.sdq, p.B, p.Hq, p.N)))
    return false;
  // dK/dV rows are written with 16-byte stores
  if (reinterpret_cast<uintptr_t>(p.dk) % 16 || reinterpret_cast<uintptr_t>(p.dv) % 16) return false;
  if (p.sdk.n % 8 || p.sdk.h % 8 || p.sdk.b % 8 || p.sdv.n % 8 || p.sdv.h % 8 || p.sdv.b % 8) return false;
  // Q and dO tiles must land in shared memory in the same row order
  const bool q_swap = (p.Hq > 1 && p.N > 1) ? (p.sq.h < p.sq.n) : false;
  const bool do_swap = (p.Hq > 1 && p.N > 1) ? (p.sdo.h < p.sdo.n) : false;
  return q_swap == do_swap;
}

// head_dim 64 with a long chunk loop per key tile (wide window / many sink-seeing rows): dK/dV on the D-generic kernel
// (ordered tensor pipe, 128-key tiles) -- 2 409 -> 2 033 us at the gpt-oss full-attention shape; the dQ side stays on
// dq64_kernel (the generic one is slower there: 1 728 vs 1 457 us)
static bool wide64_dkdv(const AttnParams& p) {
  static const char* env = getenv("SFA_WIDE64");       // diagnostics: 0 / 1 force the choice
  if (env != nullptr) return env[0] == '1';
  // measured at B=1 N=8192 Hq=64 Hkv=8 (tools/time_graph.py c1full, SFA_TG_W): window 1024: 700 vs 621 us, 2048: 998 vs
  // 1 019, 4096: 1 487 vs 1 669, 8192: 2 033 vs 2 409.  With sink tokens the key tile that holds them visits every row
  // of the sequence in ONE CTA; at head_dim 64 that CTA outlasts the rest of the launch in either kernel (a known tail:
  // 4 sink tokens cost ~0.9 ms at this shape), and dkdv64_kernel's shorter chunks make it the lesser evil.
  return p.S == 0 && (p.W < p.N ? p.W : p.N) >= SFA_WIDE64_MIN_W;
}

template <typename T>
cudaError_t tc_bwd_t(const AttnParams& p, int dtype, int stages, cudaStream_t st) {
  if (p.D != 64) return launch_bwd<T, 128>(p, dtype, stages, st);       // 64 < D <= 128 -> <128>
  if ((stages & 4) && wide64_dkdv(p) && !tc_bwd_fuses_delta(p, dtype)) {
    if (stages & 2)
      if (cudaError_t e = launch_bwd<T, 64>(p, dtype, stages & ~4, st)) return e;
    return launch_bwd<T, 64, true>(p, dtype, 4, st);
  }
  return launch_bwd<T, 64>(p, dtype, stages, st);
}

cudaError_t tc_bwd(const AttnParams& p, int dtype, int stages, cudaStream_t st) {
  return dtype == SFA_DTYPE_BF16 ? tc_bwd_t<__nv_bfloat16>(p, dtype, stages, st) : tc_bwd_t<__half>(p, dtype, stages, st);
}

}  // namespace sfa
