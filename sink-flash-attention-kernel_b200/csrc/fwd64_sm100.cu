// Persistent, warp-specialised forward for head_dim 64 (reference kernel being replaced:
// _sink_flash_attn_fwd_kernel, sink_flash_attention.py:93-194).  Same packed tile (128 MMA rows = G q-heads x
// P positions of one KV head) and two-range KV walk as fwd_sm100.cu, restructured around what the
// round-1 probes measured on B200: a tcgen05.mma costs its issuing thread ~80 cycles, an mbarrier hop on a
// single-thread role 200-300 cycles, MUFU.EX2 runs at 16/clk/SM, and the old kernel paid ~18 TMEM round trips
// plus a full prologue per tile.  Here:
//
//   warps 0-15   softmax           two groups of eight warps on alternate tiles (ping-pong); in a group two warps
//                                  per TMEM lane quarter, each owning half of the key columns of its rows; two
//                                  rolled passes (max, exp) over 16-column chunks -- small loop bodies, the
//                                  kernels are instruction-cache bound otherwise; P written as 16-bit over the
//                                  consumed S columns
//   warps 16-19  epilogue          O / l -> 16-bit -> swizzled smem -> TMA store, LSE = m ln2 + log l
//   warp 20      TMA producer      Q ring (3 tiles), K and V rings (3 items of up to 144 keys each)
//   warp 21      UMMA issuer S     S(n) = Q K^T -> S buffer n % 3, up to two items ahead of the softmax
//   warp 22      UMMA issuer PV    O += P(n) V(n)   (TS form, P read from the S buffer; one O accumulator)
//
// Online softmax in exp2 units seeded with (m, l) = (s_aux, 1) (:139-146); O is rescaled lazily (only when a
// row max moves by more than 2^8), which never happens for single-item tiles such as window 128.
#include "attn_common.cuh"
#include "tmap.cuh"

namespace sfa {
namespace {

struct Fwd64Args {
  int B, N, S, W, Hq, G, P, BN, groups_per_kv, ny, nblk, total_tiles, tiles_per_cta;
  unsigned long long bn_mul;
  int q_swap, k_swap, v_swap, o_swap;
  int fmt;
  float sl2;
  const float* s_aux;
  float* lse;
  long long* trace;
  int sp_n;      // > 0: O tiles are ALSO stored into the peer buffer of the rank owning positions [k*sp_n, (k+1)*sp_n)
  // extended geometry (sfa_fwd_ex): query row iq sits at absolute key position iq + q_off (chunked prefill, halo keys
  // of the sequence-chunk parallel mode); seq_lo[b][iq] = first key of the packed sequence row iq belongs to
  int q_off;
  const int* seq_lo;
  int64_t seq_bs;
  int single_item;   // every tile has exactly one KV item (host-known): the softmax groups hop straight to their tiles
};
struct PeerMaps {
  CUtensorMap m[8];
};
#ifndef SFA_TRACE
#define SFA_TRACE 0
#endif
__device__ __forceinline__ void tev(long long* trace, int role, int& cnt, int code, int idx) {
  if (SFA_TRACE && trace != nullptr && blockIdx.x == 0 && cnt < 256) {
    trace[(role * 256 + cnt) * 2] = (static_cast<long long>(code) << 32) | static_cast<unsigned>(idx);
    trace[(role * 256 + cnt) * 2 + 1] = clock64();
    ++cnt;
  }
}

struct Fwd64Cfg {
  // (the full-width fast path of the PV issuer assumes kParts == 2: parts split 9 chunks as 5 + 4)
  static constexpr int D = 64;
  static constexpr int kBNMax = 144;
  static constexpr int kQStages = 3, kKStages = 3, kVStages = 3;
  static constexpr int kQBytes = 128 * D * 2;
  static constexpr int kKVBytes = kBNMax * D * 2;
  static constexpr uint32_t kTmemCols = 512;
  // THREE S buffers (item n uses buffer n % 3) and ONE O accumulator.  With two S buffers S(n + 2) had to wait for
  // PV(n) -- P(n) aliases the S buffer -- and that chain (P written -> nine PV UMMAs issued at ~140 cycles each -> S
  // issued) kept each softmax group idle for a third of its period (timeline r1u_trace_fwd64_pingpong.log).  Now
  // S(n + 2) only waits for PV(n - 1).  The price is a single O accumulator: PV of the next tile starts after the
  // epilogue has read the previous tile's O (a few hundred cycles, hidden behind the other group's softmax).
  static constexpr int kSBufs = 3;
  static constexpr uint32_t kColS = 0;               // S buffers at 0, kBNMax, 2 * kBNMax
  static constexpr uint32_t kColO = kSBufs * kBNMax; // the O accumulator
  static constexpr int kMaxCh = (kBNMax / 16 + 1) / 2;
  // Two softmax GROUPS take alternate tiles (ping-pong): the warps of one group move in lock-step through a tile
  // (S wait, max exchange, P hand-over), so with a single group every scheduler's warps stall on the same latency
  // at the same time -- the timeline showed 4 200 cycles per tile against 1 470 of MUFU work.  With two groups half
  // a tile apart each scheduler always has a group in a math pass.  Per group: kParts warps per TMEM lane quarter,
  // each owning 1 / kParts of the key columns of its rows.  Measured at the gpt-oss shape: one group x 3 parts
  // 62.2 us, two groups x 2 parts 54.7 us, two groups x 3 parts (992 threads) 60.3 us.  A separate (single) TMEM
  // region for P, so that S(n + 2) need not wait for PV(n), serialises the two groups on that region: 63.3 us.
#ifndef SFA_FWD_GROUPS
#define SFA_FWD_GROUPS 2
#endif
  static constexpr int kGroups = SFA_FWD_GROUPS;
  static constexpr int kParts = 2;
  static constexpr int kGroupWarps = 4 * kParts;
  static constexpr int kSoftWarps = kGroups * kGroupWarps;
  static constexpr int kThreads = (kSoftWarps + 4 + 3) * 32;
  static constexpr int kXch = kGroups * 2 * kParts * 128;     // row-max exchange: two buffers (item parity) per group
  static constexpr int kStatFloats = 2 * 128 + 2 * kParts * 128 + kXch + 64;   // row_m, row_l, xch, s_aux (<= 64 heads cached)
  static constexpr int kSmem = 1024 + (kQStages + 1) * kQBytes + (kKStages + kVStages) * kKVBytes + kStatFloats * 4 + 512;
  static_assert(kSBufs * kBNMax + D <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

// chunk ranges of the column parts: nch 16-column chunks split as evenly as possible
__device__ __forceinline__ void part_starts(int nch, int (&cs)[Fwd64Cfg::kParts + 1]) {
  const int base = nch / Fwd64Cfg::kParts, rem = nch % Fwd64Cfg::kParts;
  cs[0] = 0;
#pragma unroll
  for (int p = 0; p < Fwd64Cfg::kParts; ++p) cs[p + 1] = cs[p] + base + (p < rem ? 1 : 0);
}

// kSingle: every tile has exactly one KV item (no sink tokens, the band fits one tile: the narrow-window training shape).
// A separate instantiation, not a run-time switch: the walker's general planning, the skipped-tile waits and the lazy O
// rescale drop out of the code -- the roles of this kernel share one instruction cache (see bwdf_sm100.cu: there 10 KB
// of cold code cost 3 us).
// kExt: chunk offset / packed-sequence bounds compiled in (same reason).
template <typename T, bool kSingle, bool kExt>
__global__ void __launch_bounds__(Fwd64Cfg::kThreads, 1) fwd64_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV,
                                                                      const __grid_constant__ CUtensorMap tmO,
                                                                      const __grid_constant__ PeerMaps pm,
                                                                      const Fwd64Args a) {
  using C = Fwd64Cfg;
  using Walk = ItemWalkT<Fwd64Args, kExt>;
  constexpr int D = C::D;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = smem;                                   // [kQStages][kQBytes]
  unsigned char* stage_s = q_s + C::kQStages * C::kQBytes;     // O staging
  unsigned char* k_s = stage_s + C::kQBytes;                   // [kKStages][kKVBytes]
  unsigned char* v_s = k_s + C::kKStages * C::kKVBytes;        // [kVStages][kKVBytes]
  float* row_m = reinterpret_cast<float*>(v_s + C::kVStages * C::kKVBytes);   // [2][128]     final running max (log2 units) per tile parity
  float* row_l = row_m + 2 * 128;                                             // [2][kParts][128]  partial row sums of the column parts
  float* xch = row_l + 2 * C::kParts * 128;                                   // [kGroups][2][kParts][128]  per-item row-max exchange between the parts
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + C::kXch + 64);
  uint64_t* q_full = bars;
  uint64_t* q_empty = q_full + C::kQStages;
  uint64_t* k_full = q_empty + C::kQStages;
  uint64_t* k_empty = k_full + C::kKStages;
  uint64_t* v_full = k_empty + C::kKStages;
  uint64_t* v_empty = v_full + C::kVStages;
  uint64_t* s_full = v_empty + C::kVStages;        // [3]  S(n) complete                      (issuer S -> softmax)
  uint64_t* p_full = s_full + C::kSBufs;           // [3]  P(n) written                       (softmax -> issuer PV)
  uint64_t* sbuf_free = p_full + C::kSBufs;        // [3]  PV(n) complete: S buffer n % 3 free (issuer PV -> issuer S, softmax rescale)
  uint64_t* o_done = sbuf_free + C::kSBufs;        // [2]  tile's O complete                  (issuer PV -> epilogue)
  uint64_t* o_free = o_done + 2;                   // [2]  O accumulator + row stats read     (epilogue -> issuer PV, softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  constexpr int kWEpi = C::kSoftWarps, kWProd = kWEpi + 4, kWIssS = kWProd + 1, kWIssPV = kWProd + 2;
  if (warp == kWProd && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < C::kQStages; ++s) { mbar_init(q_full + s, 1); mbar_init(q_empty + s, 1); }
    for (int s = 0; s < C::kKStages; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < C::kVStages; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
    for (int s = 0; s < C::kSBufs; ++s) {
      mbar_init(s_full + s, 1);
      mbar_init(p_full + s, C::kGroupWarps * 32);     // an item is handled by ONE group
      mbar_init(sbuf_free + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(o_done + s, 1);
      mbar_init(o_free + s, 128);
    }
    fence_barrier_init();
  }
  if (warp == kWIssS) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
#ifdef SFA_DEBUG_HANG
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("bars at smem 0x%x: q_full +0, q_empty +%d, k_full +%d, k_empty +%d, v_full +%d, v_empty +%d, s_full +%d, p_full +%d, sbuf_free +%d, o_done +%d, o_free +%d (bytes)\n",
           smem_u32(bars), (int)((q_empty - bars) * 8), (int)((k_full - bars) * 8), (int)((k_empty - bars) * 8), (int)((v_full - bars) * 8),
           (int)((v_empty - bars) * 8), (int)((s_full - bars) * 8), (int)((p_full - bars) * 8), (int)((sbuf_free - bars) * 8),
           (int)((o_done - bars) * 8), (int)((o_free - bars) * 8));
#endif

  if (warp == kWProd) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Walk w(a);
      while (kSingle ? w.next_single(1) : w.next()) {
        const int hq0 = w.y * a.G, kvh = w.y / a.groups_per_kv;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const int kst = w.n % C::kKStages, vst = w.n % C::kVStages;
        if (kSingle || w.t == 0) {
          const int qs = w.it % C::kQStages;
          mbar_wait(q_empty + qs, ((w.it / C::kQStages) & 1) ^ 1);
          mbar_expect_tx(q_full + qs, C::kQBytes);
          tma_tile(q_s + qs * C::kQBytes, &tmQ, q_full + qs, a.q_swap, 0, w.q0, hq0, w.b);
        }
        mbar_wait(k_empty + kst, ((w.n / C::kKStages) & 1) ^ 1);
        mbar_expect_tx(k_full + kst, a.BN * D * 2);
        tma_tile(k_s + kst * C::kKVBytes, &tmK, k_full + kst, a.k_swap, 0, kstart, kvh, w.b);
        mbar_wait(v_empty + vst, ((w.n / C::kVStages) & 1) ^ 1);
        mbar_expect_tx(v_full + vst, a.BN * D * 2);
        tma_tile(v_s + vst * C::kKVBytes, &tmV, v_full + vst, a.v_swap, 0, kstart, kvh, w.b);
      }
    }
    __syncwarp();
  } else if (warp == kWIssS) {
    // ------------------------------------------------------------------ UMMA issuer S
    if (lane == 0) {
      Walk w(a);
      int tc = 0;
      while (kSingle ? w.next_single(1) : w.next()) {
        tev(a.trace, 1, tc, 1, w.n);
        const int qs = w.it % C::kQStages, kst = w.n % C::kKStages, sb = w.n % C::kSBufs;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const uint32_t idesc_s = make_idesc(a.fmt, 128, cols, 0, 0);
        const uint64_t qd = make_sdesc(smem_u32(q_s + qs * C::kQBytes), 16, 1024);
        const uint64_t kd = make_sdesc(smem_u32(k_s + kst * C::kKVBytes), 16, 1024);
        if (kSingle || w.t == 0) mbar_wait(q_full + qs, (w.it / C::kQStages) & 1);
        mbar_wait(k_full + kst, (w.n / C::kKStages) & 1);
        if (w.n >= C::kSBufs) mbar_wait(sbuf_free + sb, ((w.n / C::kSBufs) - 1) & 1);     // PV(n-3) has consumed P(n-3)
        tc_fence_after();
        tev(a.trace, 1, tc, 2, w.n);
        const uint32_t ts = tmem + C::kColS + sb * C::kBNMax;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(ts, qd + kk * 2, kd + kk * 2, idesc_s, kk != 0);
        umma_commit(s_full + sb);
        umma_commit(k_empty + kst);
        if (kSingle || w.last_of_tile()) umma_commit(q_empty + qs);
        tev(a.trace, 1, tc, 3, w.n);
      }
    }
    __syncwarp();
  } else if (warp == kWIssPV) {
    // ------------------------------------------------------------------ UMMA issuer PV
    if (lane == 0) {
      const uint32_t idesc_pv = make_idesc(a.fmt, 128, D, 0, 1);
      Walk w(a);
      int tc = 0;
      while (kSingle ? w.next_single(1) : w.next()) {
        tev(a.trace, 3, tc, 1, w.n);
        const int vst = w.n % C::kVStages, sb = w.n % C::kSBufs, tb = w.it & 1;
        int kstart, cols; bool is_sink;
        w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
        const uint64_t vd = make_sdesc(smem_u32(v_s + vst * C::kKVBytes), C::kKVBytes, 1024);
        const uint32_t ts = tmem + C::kColS + sb * C::kBNMax;
        const int nk = cols >> 4;
        // P of column part p (chunks [cs_p, cs_{p+1})) is packed from S column 16*cs_p on: 8 columns per chunk
        int cs[C::kParts + 1];
        part_starts(nk, cs);
        mbar_wait(v_full + vst, (w.n / C::kVStages) & 1);
        mbar_wait(p_full + sb, (w.n / C::kSBufs) & 1);
        // ONE O accumulator: the epilogue must have read the previous tile's O before this tile overwrites it
        if ((kSingle || w.t == 0) && w.it >= 1) mbar_wait(o_free + (tb ^ 1), ((w.it - 1) >> 1) & 1);
        tc_fence_after();
        tev(a.trace, 3, tc, 2, w.n);
        if (nk == C::kBNMax / 16) {
          // full-width item (every tile of a 128-wide window): no per-UMMA predicates or selects -- this thread's
          // issue time gates S(n + 2), which reuses the S buffer P(n) sits in
          const uint32_t dO_ = tmem + C::kColO;
#pragma unroll
          for (int kk = 0; kk < C::kBNMax / 16; ++kk) {
            constexpr int kSplit = (C::kBNMax / 16 + C::kParts - 1) / C::kParts;      // cs[1] for kParts == 2
            const int pstart = (C::kParts == 2) ? (kk >= kSplit ? kSplit : 0) : 0;
            umma_ts(dO_, ts + pstart * 16 + (kk - pstart) * 8, vd + kk * (2048 >> 4), idesc_pv, ((!kSingle && w.t > 0) || kk > 0));
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < C::kBNMax / 16; ++kk)
            if (kk < nk) {
              int pstart = 0;
#pragma unroll
              for (int pp = 1; pp < C::kParts; ++pp) pstart = (kk >= cs[pp]) ? cs[pp] : pstart;
              umma_ts(tmem + C::kColO, ts + pstart * 16 + (kk - pstart) * 8, vd + kk * (2048 >> 4), idesc_pv,
                      ((!kSingle && w.t > 0) || kk > 0));
            }
        }
        umma_commit(sbuf_free + sb);
        umma_commit(v_empty + vst);
        if (kSingle || w.last_of_tile()) umma_commit(o_done + tb);
        tev(a.trace, 3, tc, 3, w.n);
      }
    }
    __syncwarp();
  } else if (warp < C::kSoftWarps) {
    // ------------------------------------------------------------------ softmax: row == TMEM lane, one part of the columns
    const int quarter = warp & 3, part = (warp >> 2) % C::kParts, grp = warp / C::kGroupWarps;
    float* saux_s = xch + C::kXch;          // s_aux * log2e of the first 64 heads (a global load per tile cost ~400 cycles)
    if (a.s_aux != nullptr && threadIdx.x < 64 && threadIdx.x < a.Hq) saux_s[threadIdx.x] = a.s_aux[threadIdx.x] * kLog2e;
    named_bar_sync(15, C::kSoftWarps * 32);
    const int r = quarter * 32 + lane;
    const int pr = a.q_swap ? (r / a.G) : (r % a.P);
    const int gr = a.q_swap ? (r % a.G) : (r / a.P);
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    float m_used = -INFINITY, l = 0.f;
    int i = 0, mtc = 0, row_lo = 0;
    Walk w(a);
    bool first_hop = true;
    while (true) {
      if (kSingle) {
        if (!w.next_single(first_hop ? grp + 1 : C::kGroups)) break;
        first_hop = false;
      } else {
        if (!w.next()) break;
        if ((w.it % C::kGroups) != grp) {
          // Another group's tile.  Its items are still WAITED for, one by one: a parity wait is only valid while the
          // waiter is at most one phase behind the barrier, and a tile of a wide window has many items per S buffer.
          // (Without this a group that skipped a long tile took S(n - 3)'s completion for S(n)'s: wrong results,
          // then a protocol deadlock.)
          mbar_wait(s_full + (w.n % C::kSBufs), (w.n / C::kSBufs) & 1);
          continue;
        }
      }
      const int sb = w.n % C::kSBufs, tb = w.it & 1;
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 1, w.n);
      if (kSingle || w.t == 0) {
        i = w.q0 + pr;
        if (kExt && a.seq_lo != nullptr) row_lo = (i < a.N) ? __ldg(a.seq_lo + w.b * a.seq_bs + i) : 0;
        const int h = w.y * a.G + gr;
        m_used = a.s_aux ? (h < 64 ? saux_s[h] : __ldg(a.s_aux + h) * kLog2e) : -INFINITY;
        l = (a.s_aux && part == 0) ? 1.f : 0.f;
      }
      const uint32_t ts = tl + C::kColS + sb * C::kBNMax;
      int kstart, cols; bool is_sink;
      w.pl.tile(w.t, a.BN, kstart, cols, is_sink);
      int c_lo, c_hi;
      row_range(!kSingle && is_sink, i + (kExt ? a.q_off : 0), kstart, cols, a.S, a.W, c_lo, c_hi);
      if (kExt && a.seq_lo != nullptr) c_lo = max(c_lo, row_lo - kstart);      // never across a packed-sequence boundary
      if (i >= a.N) c_hi = -1;
      const int nch = cols >> 4;
      const int pbase = nch / C::kParts, prem = nch % C::kParts;       // same split as part_starts()
      const int ch0 = part * pbase + min(part, prem), ch1 = ch0 + pbase + (part < prem ? 1 : 0);

      mbar_wait(s_full + sb, (w.n / C::kSBufs) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 2, w.n);
      auto wtrace = [&](int code) {
        if (SFA_TRACE && a.trace != nullptr && blockIdx.x == 0 && lane == 0 && (w.n == 14 || w.n == 15)) {
          const int slot = 5 * 256 + warp * 8 + (w.n - 14) * 4 + code;
          a.trace[slot * 2] = (static_cast<long long>(code + 1) << 32) | static_cast<unsigned>(w.n * 100 + warp);
          a.trace[slot * 2 + 1] = clock64();
        }
      };
      wtrace(0);
      // Two rolled passes over this half's 16-column chunks (max, then exp), one code path with the mask always on.
      // ncu showed the unrolled three-variant version starved for instructions (stall_no_inst 50-80 % of the
      // samples in the max / exp code): the hot loops must stay resident in the instruction cache that the
      // softmax, epilogue and issuer warps of a scheduler share.  S is re-read from TMEM in the second pass
      // (a tcgen05.ld round trip is ~60 cycles, probe_tmem).
      // ---- pass 1: row max over the attended columns of this half
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};    // four independent chains
#pragma unroll 1
      for (int cb = ch0; cb < ch1; ++cb) {
        uint32_t sv[16];
        tmem_ld16(ts + cb * 16, sv);
        tmem_ld_wait();
        const int lo = c_lo - cb * 16, hi = c_hi - cb * 16;
        if (__all_sync(0xffffffffu, lo <= 0 && hi >= 15)) {      // interior chunk: no mask (most chunks of a band)
#pragma unroll
          for (int e = 0; e < 16; ++e) mxa[e & 3] = fmaxf(mxa[e & 3], __uint_as_float(sv[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            mxa[e & 3] = fmaxf(mxa[e & 3], (e >= lo && e <= hi) ? __uint_as_float(sv[e]) : -INFINITY);
        }
      }
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 5, w.n);
      wtrace(1);
      float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      // ---- the parts of a row agree on its max through shared memory
      // (one pair of buffers per group: with several items per tile the two groups can be on items of equal parity)
      float* xb = xch + (grp * 2 + (w.n & 1)) * (C::kParts * 128);
      xb[part * 128 + r] = mx;
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 6, w.n);
      named_bar_sync(1 + quarter + 4 * grp, C::kParts * 32);      // ids 1-12 (up to three groups); 13-15 are taken
#pragma unroll
      for (int pp = 0; pp < C::kParts; ++pp) mx = fmaxf(mx, xb[pp * 128 + r]);
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 4, w.n);
      wtrace(2);
      const float m_new = fmaxf(m_used, mx * a.sl2);
      const bool need = (m_new - m_used) > 8.0f;        // also -inf -> finite; false for NaN (-inf - -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? exp2f(m_used - m_new) : 1.f;
        if (need) {
          l *= alpha;
          m_used = m_new;
        }
        if (!kSingle && w.t > 0) {
          // lazy rescale of this half's O columns: PV(n-1) must have completed
          mbar_wait(sbuf_free + ((w.n - 1) % C::kSBufs), ((w.n - 1) / C::kSBufs) & 1);
          tc_fence_after();
          if (part < 2) {        // warp-uniform: parts 0 and 1 rescale 32 of the 64 O columns each
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t v[16];
              const uint32_t oa = tl + C::kColO + part * 32 + cc * 16;
              tmem_ld16(oa, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
              tmem_st16(oa, v);
            }
          }
        }
      }
      // ---- pass 2: P = exp2(s*c - m) -> 16-bit over the consumed S columns of this half, partial row sum
      const float neg_m = -m_used;
      float lsum = 0.f;
#pragma unroll 1
      for (int cb = ch0; cb < ch1; ++cb) {
        uint32_t sv[16], pk[8];
        const int c0 = cb * 16;
        tmem_ld16(ts + c0, sv);
        tmem_ld_wait();
        const int lo = c_lo - c0, hi = c_hi - c0;
        if (__all_sync(0xffffffffu, lo <= 0 && hi >= 15)) {      // interior chunk: no mask
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, neg_m));
            const float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, neg_m));
            lsum += p0 + p1;
            pk[e >> 1] = pack16_fast<T>(p0, p1);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, neg_m));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, neg_m));
            p0 = (e >= lo && e <= hi) ? p0 : 0.f;
            p1 = (e + 1 >= lo && e + 1 <= hi) ? p1 : 0.f;
            lsum += p0 + p1;
            pk[e >> 1] = pack16_fast<T>(p0, p1);
          }
        }
        tmem_st8(ts + ch0 * 16 + ((c0 - ch0 * 16) >> 1), pk);
      }
      l += lsum;
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 7, w.n);
      tmem_st_wait();
      if (kSingle || w.last_of_tile()) {
        // hand the row statistics to the epilogue warps (ordered by the p_full -> PV -> o_done chain)
        if (w.it >= 2) mbar_wait(o_free + tb, ((w.it - 2) >> 1) & 1);
        if (part == 0) row_m[tb * 128 + r] = m_used;
        row_l[(tb * C::kParts + part) * 128 + r] = l;
      }
      tc_fence_before();
      mbar_arrive(p_full + sb);
      wtrace(3);
      if (threadIdx.x == 0) tev(a.trace, 4, mtc, 3, w.n);
    }
  } else if (warp < kWProd) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int pr = a.q_swap ? (r / a.G) : (r % a.P);
    const int gr = a.q_swap ? (r % a.G) : (r / a.P);
    const int ro = a.o_swap ? (pr * a.G + gr) : (gr * a.P + pr);      // row in O's box order
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int et = threadIdx.x - kWEpi * 32;
    Walk w(a);
    int mtc = 0;
    while (kSingle ? w.next_single(1) : w.next()) {
      if (!kSingle && !w.last_of_tile()) continue;
      const int tb = w.it & 1;
      if (et == 0) tev(a.trace, 6, mtc, 1, w.n);
      if (et == 0) tma_store_wait_read0();     // the previous store has finished reading the staging buffer
      named_bar_sync(14, 128);
      mbar_wait(o_done + tb, (w.it >> 1) & 1);
      tc_fence_after();
      if (et == 0) tev(a.trace, 6, mtc, 2, w.n);
      const float m_fin = row_m[tb * 128 + r];
      float l_fin = 0.f;
#pragma unroll
      for (int pp = 0; pp < C::kParts; ++pp) l_fin += row_l[(tb * C::kParts + pp) * 128 + r];
      const float inv = (l_fin > 0.f) ? 1.f / l_fin : 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[16], pk[8];
        tmem_ld16(tl + C::kColO + cc * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2)
          pk[e >> 1] = pack16<T>(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv);
        *reinterpret_cast<uint4*>(stage_s + sw128_off(ro, cc * 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(stage_s + sw128_off(ro, cc * 2 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      tc_fence_before();
      mbar_arrive(o_free + tb);
      const int i = w.q0 + pr;
      if (i < a.N)
        a.lse[(static_cast<int64_t>(w.b) * a.Hq + w.y * a.G + gr) * a.N + i] =
            (l_fin > 0.f) ? m_fin * kLn2 + logf(l_fin) : -INFINITY;
      fence_proxy_async_smem();
      named_bar_sync(13, 128);
      if (et == 0) {
        tma_tile_store(&tmO, stage_s, a.o_swap, 0, w.q0, w.y * a.G, w.b);
        if (a.sp_n > 0) {       // Ulysses: the same staged tile goes to the sequence owner over NVLink (TMA store to a peer mapping)
          const int seg = w.q0 / a.sp_n;
          tma_tile_store(&pm.m[seg], stage_s, a.o_swap, 0, w.q0 - seg * a.sp_n, w.y * a.G, w.b);
        }
        tma_store_commit();
        tev(a.trace, 6, mtc, 3, w.n);
      }
    }
    if (et == 0) {
      tma_store_wait_all0();
      // routed tiles went to another GPU: order them before everything that follows this kernel at system scope
      // (the host side's cross-rank barrier signals the peer right after the kernel)
      if (a.sp_n > 0) __threadfence_system();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWIssS) tmem_dealloc(tmem, C::kTmemCols);
}

int sm_count_fwd() { return device_sm_count(); }

template <typename T>
cudaError_t launch_fwd64(const AttnParams& p, int dtype, cudaStream_t st) {
  using C = Fwd64Cfg;
  constexpr int D = 64;
  static std::atomic<unsigned long long> attr_done[4] = {};
  if (cudaError_t e = ensure_dyn_smem(fwd64_kernel<T, false, false>, C::kSmem, attr_done[0])) return e;
  if (cudaError_t e = ensure_dyn_smem(fwd64_kernel<T, false, true>, C::kSmem, attr_done[1])) return e;
  if (cudaError_t e = ensure_dyn_smem(fwd64_kernel<T, true, false>, C::kSmem, attr_done[2])) return e;
  if (cudaError_t e = ensure_dyn_smem(fwd64_kernel<T, true, true>, C::kSmem, attr_done[3])) return e;
  const int group = p.Hq / p.Hkv;
  int G, P;
  pick_packing(p.Hq, p.Hkv, G, P);
  const int BN = pick_bn(p.W, p.Nkv, P, C::kBNMax);
  TileMap mq, mk, mv, mo;
  if (!make_tile_map(&mq, p.q, dtype, D, p.N, p.Hq, p.B, p.sq, P, G)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mk, p.k, dtype, D, p.Nkv, p.Hkv, p.B, p.sk, BN, 1)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mv, p.v, dtype, D, p.Nkv, p.Hkv, p.B, p.sv, BN, 1)) return cudaErrorInvalidValue;
  if (!make_tile_map(&mo, p.o, dtype, D, p.N, p.Hq, p.B, p.so, P, G)) return cudaErrorInvalidValue;
  Fwd64Args a;
  a.B = p.B; a.N = p.N; a.S = p.S; a.W = p.W; a.Hq = p.Hq; a.G = G; a.P = P; a.BN = BN;
  a.groups_per_kv = group / G;
  a.ny = p.Hq / G;
  a.nblk = (p.N + P - 1) / P;
  a.total_tiles = a.nblk * a.ny * p.B;
  a.tiles_per_cta = 0;
  a.bn_mul = bn_magic(BN);
  a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh; a.o_swap = mo.swap_nh;
  a.fmt = (dtype == SFA_DTYPE_BF16) ? 1 : 0;
  a.sl2 = p.scale * kLog2e;
  a.s_aux = p.s_aux;
  a.lse = p.lse;
  a.trace = trace_buffer();
  a.q_off = p.q_off; a.seq_lo = p.seq_lo; a.seq_bs = p.seq_bs;
  {
    const int64_t span = static_cast<int64_t>(p.W < p.Nkv ? p.W : p.Nkv) + P - 1;
    a.single_item = (p.S == 0 && p.W > 0 && span <= C::kBNMax) ? 1 : 0;
  }
  a.sp_n = 0;
  PeerMaps pm;
  for (int r = 0; r < 8; ++r) pm.m[r] = mo.map;
  if (p.o_route != nullptr) {
    const SpRoute& rt = *p.o_route;
    if (rt.P < 1 || rt.P > 8 || rt.n_local % P != 0 || static_cast<int64_t>(rt.n_local) * rt.P != p.N)
      return cudaErrorInvalidValue;
    const Strides4 sp{static_cast<int64_t>(rt.n_local) * rt.heads_total * D, D, static_cast<int64_t>(rt.heads_total) * D};
    const int es = 2;
    for (int r = 0; r < rt.P; ++r) {
      TileMap mp;
      void* base = static_cast<char*>(rt.peer[r]) + static_cast<int64_t>(rt.head_off) * D * es;
      if (!make_tile_map(&mp, base, dtype, D, rt.n_local, p.Hq, p.B, sp, P, G)) return cudaErrorInvalidValue;
      if (mp.swap_nh != mo.swap_nh) return cudaErrorInvalidValue;      // one staged tile serves both stores
      pm.m[r] = mp.map;
    }
    a.sp_n = rt.n_local;
  }
  const int grid = a.total_tiles < sm_count_fwd() ? a.total_tiles : sm_count_fwd();
  const bool ext = p.has_ext();
#define SFA_FWD64_LAUNCH(S_, E_) fwd64_kernel<T, S_, E_><<<grid, C::kThreads, C::kSmem, st>>>(mq.map, mk.map, mv.map, mo.map, pm, a)
  if (a.single_item) { if (ext) SFA_FWD64_LAUNCH(true, true); else SFA_FWD64_LAUNCH(true, false); }
  else { if (ext) SFA_FWD64_LAUNCH(false, true); else SFA_FWD64_LAUNCH(false, false); }
#undef SFA_FWD64_LAUNCH
  return cudaGetLastError();
}

}  // namespace

bool tc_fwd64_supported(const AttnParams& p, int dtype) {
  if (p.seq_lo != nullptr && p.S > 0) return false;      // per-sequence sink tokens: CUDA-core path
  if (p.has_ext() && p.o_route != nullptr) return false;
  return p.D == 64 && (p.S > 0 || p.W > 0);     // nothing attended at all: the one-tile-per-CTA kernel writes the O = 0 rows
}

// the routed store reuses the staged tile of the local store: the local O must be in HF order too ([B, N, H, D]
// strides, as the peer buffers are) and a tile must not straddle two owners
bool tc_fwd64_route_supported(const AttnParams& p, int dtype) {
  if (!tc_fwd64_supported(p, dtype) || p.o_route == nullptr) return false;
  int G, P;
  pick_packing(p.Hq, p.Hkv, G, P);
  const SpRoute& rt = *p.o_route;
  const bool hf = (p.Hq > 1 && p.N > 1) ? (p.so.h < p.so.n) : true;
  return hf && rt.P >= 1 && rt.P <= 8 && rt.n_local % P == 0 && static_cast<int64_t>(rt.n_local) * rt.P == p.N;
}

cudaError_t tc_fwd64(const AttnParams& p, int dtype, cudaStream_t st) {
  return dtype == SFA_DTYPE_BF16 ? launch_fwd64<__nv_bfloat16>(p, dtype, st) : launch_fwd64<__half>(p, dtype, st);
}

}  // namespace sfa
