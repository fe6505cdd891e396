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
// order, so each k-step of dV^T / dK^T is ONE UMMA over the whole ring (every UMMA costs >= 55 cycles whatever
// its N, tools/probe_mma_desc.py); the image columns of the one slot outside the window are zero.  After the
// tile that last touches a block, the epilogue warps read its slot (dV^T and dK^T with one tcgen05.ld) and write
// dK/dV -- the GQA group sum happened inside the contraction over the rows.  The slot is cleared for its next
// block by a UMMA against a zero operand, i.e. in tensor-pipe order (a tcgen05.st would race the read-modify-
// write of the in-flight ring UMMAs).
//
// A CTA owns a contiguous run of tiles.  Key blocks shared with the neighbouring CTA (the nb - 1 blocks before
// its first tile and the last nb - 1 blocks of its run) are written as fp32 partials and summed by a small
// fix-up kernel: no atomics, no inter-CTA waits, deterministic.
//
// Warp roles (25 warps): 0-11 math (lane quarter = warp & 3, a third of the 16-column chunks each), 12-15 and
// 16-19 two epilogue groups taking alternate tiles (dQ store, ring drain), 20 TMA producer, 21 UMMA issuer
// S / dP, 22 UMMA issuer dV^T, 23 UMMA issuer dK^T, 24 UMMA issuer dQ.  An issuing thread spends 100-150 cycles
// per tcgen05.mma here (9 instructions on the uniform datapath, competing with six busy warps per scheduler)
// while the tensor pipe needs 55-80: one issuer for all five products ran at ~8000 cycles per tile.
//
// delta = rowsum(dO o O) (sink_flash_attention.py:582) is computed IN the kernel -- no separate streaming pass over
// O and dO (26 us at the gpt-oss shape, with dO read a second time): the epilogue group of tile n computes
// delta(n + 3) in the idle time after its epilogue, from the dO tile the TMA already put into shared memory and
// an O tile loaded beside it, into a 128-float buffer the math warps read before pass 2.  The O buffer's 16 KB
// come from V, which is single-buffered: only dP(n) reads V(n), so its buffer is free again half a tile before
// V(n + 1) is needed.  The O loads are issued by the group that just consumed the buffer (not by the producer
// warp: its in-order waits would tie every other load to the delta computation).  What did not work: O and dO
// rows as 16-byte global loads in the epilogue groups (round 1: 152 us, 8 rows in flight per warp); three
// dedicated delta warps with the O rows in registers (152 us: two exposed memory round trips per tile, no
// registers for a second batch in flight); three delta warps reading both tiles from shared memory (149 us:
// ~4900 cycles per tile for ~500 instructions -- a 26th-28th warp gets an issue slot every ~10 cycles here).
#include <stdio.h>
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
  static constexpr int kThreads = 25 * 32;
  static constexpr int kStageBytes = 8 * 1024;         // dQ store transpose, [32 rows][32 B] per epilogue warp
  static constexpr int kZeroBytes = 1024;              // zero B operand of the slot-clearing UMMAs (P <= 32 columns)
  static constexpr int kDeltaBytes = 2 * 128 * 4;      // delta rows of two tiles in flight
  // no alignment slack: the dynamic shared memory is declared 1024-byte aligned (checked at kernel entry)
  static constexpr int kSmem = 5 * kQBytes + 3 * kKVBytes + 2 * kPBytes + kStageBytes + kZeroBytes + kDeltaBytes + 512;
  static constexpr int kPartKeys = 128;                // keys per side of a CTA's fp32 partials
  static constexpr int kMaxCtas = 160;
  static_assert(kColR + kRingCols <= 512, "TMEM budget");
  static_assert(kSmem <= 227 * 1024, "smem budget");
};

struct FusedArgs {
  int B, N, W, Hq, Hkv, G, P, lgP, nb, R, cols, nch, nblk, total_tiles, tiles_per_cta;
  int q_swap, k_swap, v_swap;
  int fuse_delta;  // 1: the epilogue groups compute delta in the kernel; 0: a preprocess kernel wrote it before
  int write_delta; // fuse_delta only: ds_aux wanted -- per (head, tile[, quarter]) partials of -sum exp(s_aux - lse) * delta
                   // (sink_flash_attention.py:653-665) go to the workspace, the fix-up launch sums them in a fixed order
  const float* s_aux;
  int dbg_delay;   // test knob (sfa_set_debug): the part-1 math warps sleep this many ns before pass 2, the epilogue
                   // groups before their dQ stores -- widens every cross-warp window of the pipeline
  int dbg_norace;  // test knob: 1 drops the per-quarter barrier that orders the P-image reads of pass 2(n) before the
                   // writes of pass 1(n + 1) (reproduces the round-1 run-to-run difference in dQ / dK on one GPU)
  int fmt;       // 0 f16, 1 bf16
  float sl2;     // scale * log2(e)
  float scale;
  const float* lse;
  float* delta;          // [B,Hq,N] workspace

  const void* k;         // (unused since the L2-prefetch experiment left the kernel)
  const void* v;
  Strides4 sk, sv;
  void* dq;
  void* dk;
  void* dv;
  Strides4 sdq, sdk, sdv;
  void* dq_peer[8];   // Ulysses routing of dQ: rows of positions [k*dq_seg_n, (k+1)*dq_seg_n) go to dq_peer[k] (sdq = peer strides)
  int dq_seg_n;       // 0: off, dq / sdq describe the local tensor
  // extended geometry (sfa_bwd_ex): query tile pb covers key blocks pb + qb - nb + 1 .. pb + qb (qb = q_off / P: the
  // queries are the LAST rows of a longer key axis of Nkv rows -- chunked prefill, halo keys); seq_lo: packed sequences
  int qb, Nkv;
  const int* seq_lo;
  int64_t seq_bs;
  float* part;   // [grid][2 sides][kPartKeys][2 (dV, dK)][64] fp32
  long long* trace;   // optional timeline buffer (sfa_set_trace_buffer, -DSFA_TRACE=1 builds); nullptr in production
};

// Timeline probe for performance work (tools/trace_fused.py): CTA 0 appends (role, code, tile, clock64) records.
#ifndef SFA_TRACE
#define SFA_TRACE 0
#endif
__device__ __forceinline__ void ftrace(long long* trace, int role, int& cnt, int code, int idx) {
  // trace[2 * 9 * 256]: the CTA to record (sfa_set_debug knob 2; timeline builds only)
  if (SFA_TRACE && trace != nullptr && static_cast<long long>(blockIdx.x) == trace[2 * 9 * 256] && cnt < 256) {
    trace[(role * 256 + cnt) * 2] = (static_cast<long long>(code) << 32) | static_cast<unsigned>(idx);
    trace[(role * 256 + cnt) * 2 + 1] = clock64();
    ++cnt;
  }
}

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
// (the start coordinates need two integer divisions: computed ONCE per thread before the role dispatch -- every
// role builds its own walker, and the kernel is instruction-cache bound)
struct WalkInit {
  int tile, end, pb, y, b;
  __device__ __forceinline__ explicit WalkInit(const FusedArgs& a) {
    tile = static_cast<int>(blockIdx.x) * a.tiles_per_cta;
    end = min(tile + a.tiles_per_cta, a.total_tiles);
    pb = tile % a.nblk;
    const int r = tile / a.nblk;
    y = r % a.Hkv;
    b = r / a.Hkv;
  }
};
struct FusedWalk {
  const FusedArgs& a;
  int tile, end, pb, y, b, it;
  bool seg_first;     // first tile of a sequence segment inside this CTA
  __device__ __forceinline__ FusedWalk(const FusedArgs& a_, const WalkInit& wi)
      : a(a_), tile(wi.tile - 1), end(wi.end), pb(wi.pb - 1), y(wi.y), b(wi.b), it(-1), seg_first(false) {}
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

// kExt: extended geometry (packed sequences / chunk offset) compiled in.  A separate instantiation, not a run-time
// switch: the two extra live values per math thread pushed the plain kernel over its 72-register budget (36 B of
// spills -- with 227 KB of shared memory every spill reload is an L2 round trip) and cost 4.5 us at the gpt-oss shape.
// kLean: the production launch -- delta inside the kernel, no routed dQ, no debug knobs: their branches are compiled out
// (the size of this kernel's code is a first-order parameter).
template <typename T, bool kExt, bool kLean>
__global__ void __launch_bounds__(FusedCfg::kThreads, 1) bwd_fused64_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                             const __grid_constant__ CUtensorMap tmdO,
                                                                             const __grid_constant__ CUtensorMap tmK,
                                                                             const __grid_constant__ CUtensorMap tmV,
                                                                             const __grid_constant__ CUtensorMap tmO,
                                                                             const FusedArgs a) {
  using C = FusedCfg;
  extern __shared__ __align__(1024) unsigned char smem_al[];
  unsigned char* smem = smem_al;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();       // SWIZZLE_128B tiles need 1024-byte aligned slabs
  unsigned char* q_s = smem;                          // [2][kQBytes]
  unsigned char* do_s = q_s + 2 * C::kQBytes;         // [2][kQBytes]
  unsigned char* k_s = do_s + 2 * C::kQBytes;         // [2][kKVBytes]
  unsigned char* v_s = k_s + 2 * C::kKVBytes;         // [kKVBytes]
  unsigned char* o_s = v_s + C::kKVBytes;             // [kQBytes]  O tile (delta)
  unsigned char* p_s = o_s + C::kQBytes;              // P image
  unsigned char* ds_s = p_s + C::kPBytes;             // dS image
  unsigned char* stage_s = ds_s + C::kPBytes;         // [8 warps][32 rows][32 B]: dQ store transpose
  unsigned char* z_s = stage_s + C::kStageBytes;      // zeros: B operand of the slot-clearing UMMAs
  float* delta_s = reinterpret_cast<float*>(z_s + C::kZeroBytes);   // [2][128]: delta of the rows of tiles n, n + 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(z_s + C::kZeroBytes + C::kDeltaBytes);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = q_full + 2;     // [2]  dK^T(n) complete
  uint64_t* k_full = q_empty + 2;
  uint64_t* k_empty = k_full + 2;     //      dQ(n) complete
  uint64_t* do_full = k_empty + 2;
  uint64_t* do_empty = do_full + 2;   //      dV^T(n) complete (+ delta(n) computed: the four warps of epilogue group (n & 1) ^ 1)
  uint64_t* v_full = do_empty + 2;    // [0]: V(n), phase n & 1;  [1]: unused
  uint64_t* v_empty = v_full + 2;     // [0]: dP(n) complete;     [1]: unused
  uint64_t* s_full = v_empty + 2;     // S(n) complete                          (issuer A -> math)
  uint64_t* s_free = s_full + 1;      // S(n) read                              (math -> issuer A)
  uint64_t* dp_full = s_free + 1;     // dP(n) complete                         (issuer A -> math)
  uint64_t* dp_free = dp_full + 1;    // dP(n) read                             (math -> issuer A)
  uint64_t* p_ready = dp_free + 1;    // P(n) in shared memory                  (math -> issuer V)
  uint64_t* p_free = p_ready + 1;     // dV^T(n) complete: P image reusable     (issuer V -> math)
  uint64_t* ds_ready = p_free + 1;    // dS(n) in shared memory                 (math -> issuer K)
  uint64_t* ds_free = ds_ready + 1;   // dK^T(n), dQ(n) complete                (issuer K -> math)
  uint64_t* dq_done = ds_free + 1;    // [2] all UMMAs of tile n complete       (issuers V + K -> epilogue group n & 1)
  uint64_t* dq_free = dq_done + 2;    // [2] dQ(n) read                         (epilogue group n & 1 -> issuer K)
  uint64_t* drain_done = dq_free + 2; // [2] ring drain of tile n finished      (epilogue group n & 1 -> issuers V, K)
  uint64_t* delta_ready = drain_done + 2;   // [2] delta rows of tile n in delta_s[n & 1]   (the OTHER epilogue group, (n & 1) ^ 1 -> math)
  uint64_t* delta_free = delta_ready + 2;   // [2] ... read by every math thread            (math -> epilogue group)
  uint64_t* kq_issued = delta_free + 2;    // unused (the dP-ordering experiment it served is gone); keeps the layout
  uint64_t* o_full = kq_issued + 1;         // [8 blocks of 16 rows][2]: O rows of tile n in block j: barrier [j][n & 1], phase
                                            // (n >> 1) & 1 -- per tile parity because the epilogue groups take alternate
                                            // tiles (a parity wait cannot tell phase n from phase n - 2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WalkInit wi(a);

  if (warp == 20 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if ((kLean || a.fuse_delta)) tma_prefetch_desc(&tmO);
    // with delta in the kernel: do_empty[s] (bars 10, 11) = dV^T(n) complete + the four warps that computed delta(n)
    for (int s = 0; s < 16; ++s) mbar_init(bars + s, ((kLean || a.fuse_delta) && (s == 10 || s == 11)) ? 5 : 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, C::kMathWarps * 32);       // per-thread arrivals and waits in the math warps: measured faster than
                                                 // one polling / arriving lane per warp + __syncwarp (fwd64: 62 vs 67 us)
    mbar_init(dp_full, 1);
    mbar_init(dp_free, C::kMathWarps * 32);
    mbar_init(p_ready, C::kMathWarps * 32);
    mbar_init(p_free, 1);
    mbar_init(ds_ready, C::kMathWarps * 32);
    mbar_init(ds_free, 2);
    for (int g = 0; g < 2; ++g) {
      mbar_init(dq_done + g, 3);
      mbar_init(dq_free + g, 4);
      mbar_init(drain_done + g, 4);
      mbar_init(delta_ready + g, 4);
      mbar_init(delta_free + g, C::kMathWarps * 32);
    }
    mbar_init(kq_issued, 2);
    for (int j = 0; j < 16; ++j) mbar_init(o_full + j, 1);
    fence_barrier_init();
  }
  if (warp == 21) tmem_alloc(tmem_slot, C::kTmemCols);
  for (int i = threadIdx.x; i < C::kZeroBytes / 4; i += C::kThreads) reinterpret_cast<uint32_t*>(z_s)[i] = 0u;
  fence_proxy_async_smem();
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
  const int qb = kExt ? a.qb : 0, nkv = kExt ? a.Nkv : a.N;
  const int* const seq_lo = kExt ? a.seq_lo : nullptr;

  if (warp >= 20) {
  if (warp == 20) {
    // ------------------------------------------------------------------ TMA producer (lane 0) + L2 prefetch (warp)
    {
      // (An L2 prefetch of the next tiles ahead of the TMA loads, and dP(n + 1) queued behind dK^T(n) / dQ(n), were both
      // measured slower in round 1; their code is gone: the size of this kernel's code is a first-order parameter.)
      FusedWalk w(a, wi);
      const uint32_t kv_bytes = a.cols * C::D * 2;
      int tc = 0;
      if ((kLean || a.fuse_delta) && lane == 0 && wi.tile < wi.end) {        // O(0), as eight 16-row blocks; the epilogue warps load the rest
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const int n0 = wi.pb * P + (a.q_swap ? ((16 * j) >> (7 - a.lgP)) : ((16 * j) & (P - 1)));
          const int h0 = wi.y * a.G + (a.q_swap ? 0 : ((16 * j) >> a.lgP));
          mbar_expect_tx(o_full + j * 2, 2048);
          tma_load_4d(o_s + j * 2048, &tmO, o_full + j * 2, 0, a.q_swap ? h0 : n0, a.q_swap ? n0 : h0, wi.b);
        }
      }
      while (w.next()) {
        if (lane == 0) {
          ftrace(a.trace, 0, tc, 1, w.it);
          const int s = w.it & 1;
          const uint32_t eph = ((w.it >> 1) & 1) ^ 1;
          const int q0 = w.pb * P, hq0 = w.y * a.G, kstart = (w.pb + qb - nb + 1) * P;
          // dO first: its buffer is released by dV^T(n - 2), half a tile before dK^T(n - 2) / dQ(n - 2) release Q and
          // K -- and the delta of tile n (epilogue groups) wants dO(n) as early as it can get it
          mbar_wait(do_empty + s, eph);
          mbar_expect_tx(do_full + s, C::kQBytes);
          tma_tile(do_s + s * C::kQBytes, &tmdO, do_full + s, a.q_swap, 0, q0, hq0, w.b);
          mbar_wait(q_empty + s, eph);
          mbar_expect_tx(q_full + s, C::kQBytes);
          tma_tile(q_s + s * C::kQBytes, &tmQ, q_full + s, a.q_swap, 0, q0, hq0, w.b);
          mbar_wait(k_empty + s, eph);
          mbar_expect_tx(k_full + s, kv_bytes);
          tma_tile(k_s + s * C::kKVBytes, &tmK, k_full + s, a.k_swap, 0, kstart, w.y, w.b);
          mbar_wait(v_empty, (w.it & 1) ^ 1);             // single buffer: dP(n - 1) complete
          mbar_expect_tx(v_full, kv_bytes);
          tma_tile(v_s, &tmV, v_full, a.v_swap, 0, kstart, w.y, w.b);
          ftrace(a.trace, 0, tc, 3, w.it);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == 21) {
    // ------------------------------------------------------------------ UMMA issuer A: S = Q K^T, dP = dO V^T
    if (lane == 0) {
      const uint32_t idesc = make_idesc(a.fmt, 128, a.cols, 0, 0);
      const uint64_t qd0 = make_sdesc(smem_u32(q_s), 16, 1024), kd0 = make_sdesc(smem_u32(k_s), 16, 1024);
      const uint64_t dod0 = make_sdesc(smem_u32(do_s), 16, 1024), vd0 = make_sdesc(smem_u32(v_s), 16, 1024);
      FusedWalk w(a, wi);
      int tc = 0;
      while (w.next()) {
        ftrace(a.trace, 1, tc, 1, w.it);
        const int s = w.it & 1;
        const uint32_t fph = (w.it >> 1) & 1;
        const uint64_t qd = qd0 + s * (C::kQBytes >> 4), kd = kd0 + s * (C::kKVBytes >> 4);
        const uint64_t dod = dod0 + s * (C::kQBytes >> 4), vd = vd0;
        mbar_wait(q_full + s, fph);
        mbar_wait(k_full + s, fph);
        if (w.it >= 1) mbar_wait(s_free, (w.it - 1) & 1);
        tc_fence_after();
        ftrace(a.trace, 1, tc, 2, w.it);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C::kColS, qd + kk * 2, kd + kk * 2, idesc, kk != 0);
        umma_commit(s_full);
        ftrace(a.trace, 1, tc, 3, w.it);
        mbar_wait(do_full + s, fph);
        mbar_wait(v_full, w.it & 1);
        if (w.it >= 1) {
          mbar_wait(dp_free, (w.it - 1) & 1);
        }
        tc_fence_after();
        ftrace(a.trace, 1, tc, 4, w.it);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C::kColP, dod + kk * 2, vd + kk * 2, idesc, kk != 0);
        umma_commit(dp_full);
        umma_commit(v_empty);
        ftrace(a.trace, 1, tc, 5, w.it);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ UMMA issuers V (dV^T), K (dK^T), Q (dQ)
    // Each k-step covers the WHOLE ring with one UMMA: the image columns of the slot whose block left the window
    // with the previous tile are zero (written by the math warps), so that slot -- possibly being drained right
    // now -- only gets +0.  The slot the tile's newest block enters is cleared first by a one-k-step UMMA against
    // a zero B operand (accumulate off): the clearing is ordered with the other UMMAs, not with the drain.
    if (lane == 0 && warp == 24) {
      // dQ = dS K: A = dS image (K-major, un-swizzled), B = K tile (MN-major)
      const uint32_t idesc_dq = make_idesc(a.fmt, 128, C::D, 0, 1);
      const uint64_t dsa0 = make_sdesc_ns(smem_u32(ds_s), 2048, 128);
      const uint64_t kb0 = make_sdesc(smem_u32(k_s), C::kKVBytes, 1024);
      const uint32_t wrap16 = static_cast<uint32_t>(R * P) * 16u;
      FusedWalk w(a, wi);
      SlotTrack st;
      st.slot0 = 0;
      int tc = 0;
      while (w.next()) {
        ftrace(a.trace, 7, tc, 1, w.it);
        st.step(w, R);
        const int s = w.it & 1;
        const uint32_t fph = (w.it >> 1) & 1;
        const uint64_t kb = kb0 + s * (C::kKVBytes >> 4);
        mbar_wait(ds_ready, w.it & 1);
        mbar_wait(k_full + s, fph);
        if (w.it >= 1) mbar_wait(dq_free + ((w.it - 1) & 1), ((w.it - 1) >> 1) & 1);
        tc_fence_after();
        ftrace(a.trace, 7, tc, 2, w.it);
        // natural chunk kk of the tile sits at ring column (slot0 * P + 16 kk) mod (R * P); one image column
        // group of 8 is 2048 B = 128 descriptor units, i.e. 16 units per column
        uint32_t rc16 = static_cast<uint32_t>(st.slot0 * P) * 16u;
#pragma unroll 1
        for (int kk = 0; kk < a.nch; ++kk) {
          umma_ss(tmem + C::kColQ, dsa0 + rc16, kb + kk * (2048 >> 4), idesc_dq, kk > 0);
          rc16 += 256u;
          if (rc16 >= wrap16) rc16 -= wrap16;
        }
        umma_commit(ds_free);
        umma_commit(k_empty + s);
        umma_commit(dq_done + s);
        ftrace(a.trace, 7, tc, 3, w.it);
      }
    } else if (lane == 0) {
      const bool isK = (warp == 23);
      const uint32_t idesc_ring = make_idesc(a.fmt, 64, R * P, 1, 1);
      const uint32_t idesc_zero = make_idesc(a.fmt, 64, P, 1, 1);
      const uint64_t img = make_sdesc_ns(smem_u32(isK ? ds_s : p_s), 128, 2048);          // MN-major B: ring columns
      const uint64_t zero_b = make_sdesc_ns(smem_u32(z_s), 128, 256);
      const uint64_t a0 = make_sdesc(smem_u32(isK ? q_s : do_s), 16384, 1024);            // MN-major A: Q^T / dO^T
      const uint32_t ring = tmem + C::kColR + (isK ? (16u << 16) : 0u);
      uint64_t* const in_ready = isK ? ds_ready : p_ready;
      uint64_t* const a_full = isK ? q_full : do_full;
      uint64_t* const a_empty = isK ? q_empty : do_empty;
      const int role = isK ? 5 : 2;
      FusedWalk w(a, wi);
      SlotTrack st;
      st.slot0 = 0;
      int tc = 0;
      while (w.next()) {
        ftrace(a.trace, role, tc, 1, w.it);
        st.step(w, R);
        const int s = w.it & 1;
        const uint32_t fph = (w.it >> 1) & 1;
        int newest = st.slot0 + nb - 1;
        if (newest >= R) newest -= R;
        const uint64_t ad = a0 + s * (C::kQBytes >> 4);
        mbar_wait(in_ready, w.it & 1);
        mbar_wait(a_full + s, fph);
        // the slot the tile's newest block enters must have been drained
        if (w.seg_first) {
          if (w.it >= 1) mbar_wait(drain_done + ((w.it - 1) & 1), ((w.it - 1) >> 1) & 1);
        } else if (w.it >= 2) {
          mbar_wait(drain_done + (w.it & 1), ((w.it - 2) >> 1) & 1);
        }
        tc_fence_after();
        ftrace(a.trace, role, tc, 2, w.it);
        umma_ss(ring + newest * P, ad, zero_b, idesc_zero, 0);
        // fully unrolled: this thread's issue rate is on the pipeline's critical cycle (the P / dS image is free for
        // the next tile only when these UMMAs have completed); `unroll 2` cost 4.6 us on the kernel
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss(ring, ad + kk * (2048 >> 4), img + kk * (256 >> 4), idesc_ring, 1);
        umma_commit(a_empty + s);
        umma_commit(isK ? ds_free : p_free);
        umma_commit(dq_done + s);
        ftrace(a.trace, role, tc, 3, w.it);
      }
    }
    __syncwarp();
  }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                     // TMEM lane == MMA row
    const int pr = a.q_swap ? (r / a.G) : (r & (P - 1));
    const int gr = a.q_swap ? (r & (a.G - 1)) : (r >> a.lgP);
    const uint32_t tl = tmem + (static_cast<uint32_t>(quarter * 32) << 16);

    if (warp < C::kMathWarps) {
      // ---------------------------------------------------------------- math warps
      const int part = warp >> 2;                          // chunks part, part + 3, part + 6
      const uint32_t rowoff = static_cast<uint32_t>((r >> 3) * 128 + (r & 7) * 16);
      const uint32_t p_a = smem_u32(p_s) + rowoff, ds_a = smem_u32(ds_s) + rowoff;
      const int RP = R * P;
      auto load_row = [&](const float* src, const FusedWalk& t, bool valid, float dflt) {
        float v = dflt;
        const int i = t.pb * P + pr;
        if (valid && i < a.N) {
          const int64_t row = (static_cast<int64_t>(t.b) * a.Hq + t.y * a.G + gr) * a.N + i;
          asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(src + row));
        }
        return v;
      };
      // delta of tile t from the preprocess pass (fuse_delta == 0); with fuse_delta it comes from delta_s before pass 2
      auto load_delta = [&](const FusedWalk& t, bool valid) {
        float v = 0.f;
        if (valid && !(kLean || a.fuse_delta)) {
          const int i = t.pb * P + pr;
          if (i < a.N) {
            const int64_t row = (static_cast<int64_t>(t.b) * a.Hq + t.y * a.G + gr) * a.N + i;
            asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(a.delta + row) : "memory");
          }
        }
        return v;
      };
      const uint64_t scale2 = pack_f32x2(a.scale, a.scale), sl2_2 = pack_f32x2(a.sl2, a.sl2);
      FusedWalk w(a, wi), wn(a, wi);
      SlotTrack st;
      st.slot0 = 0;
      bool has_next = wn.next();
      // first key of the row's packed sequence, in query-row units (absolute - q_off); unused when not packed
      auto load_lo = [&](const FusedWalk& t, bool valid) {
        int v = 0;
        if (kExt && seq_lo != nullptr && valid) {
          const int i = t.pb * P + pr;
          if (i < a.N) v = __ldg(seq_lo + t.b * a.seq_bs + i) - qb * P;
        }
        return v;
      };
      float l_next = load_row(a.lse, wn, has_next, INFINITY), d_next = load_delta(wn, has_next);
      int lo_next = load_lo(wn, has_next);
      int tc = 0;
      const bool tr = SFA_TRACE && (threadIdx.x == 0);
      while (w.next()) {
        if (tr) ftrace(a.trace, 3, tc, 1, w.it);
        st.step(w, R);
        const float lse_i = l_next;
        float dsc = d_next * a.scale;
        const int lo_row = lo_next;
        has_next = wn.next();
        l_next = load_row(a.lse, wn, has_next, INFINITY);
        d_next = load_delta(wn, has_next);
        lo_next = load_lo(wn, has_next);
        const float neg_l2 = (lse_i == -INFINITY) ? -INFINITY : -lse_i * kLog2e;   // lse = +-inf: P = 0
        const uint64_t negl2_2 = pack_f32x2(neg_l2, neg_l2);
        const int i = w.pb * P + pr;
        const int kstart = (w.pb - nb + 1) * P;            // in QUERY-row units: absolute key = this + q_off
        // attended keys of the row: [max(i - W + 1, first key of its sequence), i]; keys start at absolute 0
        // (kept as two plain steps: the one-expression form with a -2^30 "no sequence" sentinel folded into a
        // three-way max came out of ptxas 12.9 with the sign of the q_off term lost -- caught by the chunked-prefill tests)
        int key_lo = -qb * P;                              // absolute key 0
        if (kExt && seq_lo != nullptr) key_lo = max(key_lo, lo_row);
        const int c_lo = max(i - a.W + 1, key_lo) - kstart;
        const int c_hi = (i < a.N) ? (i - kstart) : -1;
        int xs = st.slot0 + nb;                            // slot whose block left the window: zero image columns
        if (xs >= R) xs -= R;
        const uint32_t zoff = static_cast<uint32_t>(xs * P) * 256u;
        // ---- pass 1: P = exp2(S * c - lse), masked, 16-bit -> P image (ring-column order).  The loops over this
        // warp's chunks are ROLLED: the kernel's roles run concurrently and share a ~32 KB instruction cache (the
        // fully unrolled version, 73 KB of code, ran 2x slower in every role).
        mbar_wait(s_full, w.it & 1);
        tc_fence_after();
        if (w.it >= 1) mbar_wait(p_free, (w.it - 1) & 1);
        if (tr) ftrace(a.trace, 3, tc, 2, w.it);
        // Pass 2(n) reads P back from the image, and the ring-column order shifts by one key block per tile: the cell
        // [row r, ring chunk c] this thread writes now for tile n + 1 was natural chunk cb + 1 of tile n, owned -- and
        // read in pass 2(n) -- by the warp of the NEXT part of the same lane quarter (and the zeroed slot by part 0).
        // p_free only orders the tensor pipe's reads; the three warps of a quarter (same rows, all chunks) must agree
        // that pass 2(n) is over before any of them overwrites the image.  Without this barrier a warp that ran a full
        // pass ahead of its neighbour fed P(n + 1) into dS(n): dQ and dK changed from run to run, dV never (round 1).
        if (w.it >= 1 && !(!kLean && a.dbg_norace)) named_bar_sync(1 + quarter, 96);
        if (part == 1)
          for (int g8 = 0; g8 < (P >> 3); ++g8) st_shared_v4(p_a + zoff + g8 * 2048u, 0u, 0u, 0u, 0u);
        const int rc0 = st.slot0 * P;
#pragma unroll 1
        for (int cb = part; cb < a.nch; cb += 3) {
          uint32_t sv[16], pk[8];
          tmem_ld16(tl + C::kColS + cb * 16, sv);
          tmem_ld_wait();
          int rc = rc0 + cb * 16;                          // ring column of natural chunk cb, 256 B per column
          if (rc >= RP) rc -= RP;
          const uint32_t off = static_cast<uint32_t>(rc) * 256u;
          const int lo = c_lo - cb * 16, hi = c_hi - cb * 16;      // attended elements of this chunk: [lo, hi]
          if (__all_sync(0xffffffffu, lo <= 0 && hi >= 15)) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float x0, x1;
              unpack_f32x2(fma_f32x2(sv[e], sv[e + 1], sl2_2, negl2_2), x0, x1);
              pk[e >> 1] = pack16<T>(fast_exp2(x0), fast_exp2(x1));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), a.sl2, neg_l2));
              float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), a.sl2, neg_l2));
              p0 = (e >= lo && e <= hi) ? p0 : 0.f;
              p1 = (e + 1 >= lo && e + 1 <= hi) ? p1 : 0.f;
              pk[e >> 1] = pack16<T>(p0, p1);
            }
          }
          st_shared_v4(p_a + off, pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(p_a + off + 2048u, pk[4], pk[5], pk[6], pk[7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(s_free);
        mbar_arrive(p_ready);
        if (tr) ftrace(a.trace, 3, tc, 3, w.it);
        // ---- pass 2: dS = scale * P o (dP - delta), 16-bit -> dS image (dQ and dK both carry the scale); P comes
        // back from the image (this thread's own row)
        mbar_wait(dp_full, w.it & 1);
        tc_fence_after();
        if (w.it >= 1) mbar_wait(ds_free, (w.it - 1) & 1);
        if ((kLean || a.fuse_delta)) {                                // this row's delta from the delta warps
          mbar_wait(delta_ready + (w.it & 1), (w.it >> 1) & 1);
          dsc = delta_s[(w.it & 1) * 128 + r] * a.scale;
          mbar_arrive(delta_free + (w.it & 1));
          if (tr) ftrace(a.trace, 3, tc, 6, w.it);
        }
        const uint64_t ndsc2 = pack_f32x2(-dsc, -dsc);
        if (!kLean && a.dbg_delay && part == 1) __nanosleep(a.dbg_delay);
        if (tr) ftrace(a.trace, 3, tc, 4, w.it);
        if (part == 1)
          for (int g8 = 0; g8 < (P >> 3); ++g8) st_shared_v4(ds_a + zoff + g8 * 2048u, 0u, 0u, 0u, 0u);
#pragma unroll 1
        for (int cb = part; cb < a.nch; cb += 3) {
          uint32_t dv[16], pk[8], dk[8];
          tmem_ld16(tl + C::kColP + cb * 16, dv);
          int rc = rc0 + cb * 16;
          if (rc >= RP) rc -= RP;
          const uint32_t off = static_cast<uint32_t>(rc) * 256u;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3]) : "r"(p_a + off) : "memory");
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(pk[4]), "=r"(pk[5]), "=r"(pk[6]), "=r"(pk[7]) : "r"(p_a + off + 2048u) : "memory");
          tmem_ld_wait();
          // t = scale * (dP - delta) as one FFMA2 per pair, rounded to 16 bit, then dS = P * t as one HMUL2 per
          // pair on the packed P straight from the image: 3 issued instructions per pair instead of 7 (masked P is
          // exactly 0 and dP is finite: dS = 0 there)
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float t0, t1;
            unpack_f32x2(fma_f32x2(dv[e], dv[e + 1], scale2, ndsc2), t0, t1);
            dk[e >> 1] = mul16x2<T>(pk[e >> 1], pack16<T>(t0, t1));
          }
          st_shared_v4(ds_a + off, dk[0], dk[1], dk[2], dk[3]);
          st_shared_v4(ds_a + off + 2048u, dk[4], dk[5], dk[6], dk[7]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(dp_free);
        mbar_arrive(ds_ready);
        if (tr) ftrace(a.trace, 3, tc, 5, w.it);
      }
    } else {
      // ---------------------------------------------------------------- epilogue warps: dQ store, ring drain
      // two groups of four warps (one per lane quarter); group g takes the tiles with (it & 1) == g
      const int grp = (warp - C::kMathWarps) >> 2;
      const int which = lane >> 4;                          // 0: dV^T (lanes 0-15), 1: dK^T (lanes 16-31)
      const int dch = quarter * 16 + (lane & 15);           // channel of this lane's ring row
      T* const okv = static_cast<T*>(which ? a.dk : a.dv);
      const Strides4 skv = which ? a.sdk : a.sdv;
      float* const part_cta = a.part + static_cast<size_t>(blockIdx.x) * 2 * C::kPartKeys * 128 + which * 64 + dch;
      // dQ rows go through a per-warp [32 rows][32 B] transpose buffer: 2 lanes write 32 contiguous bytes of a row
      const uint32_t stg = smem_u32(stage_s) + (warp - C::kMathWarps) * 1024;
      const uint32_t st_wr = stg + lane * 32, st_key = (lane >> 2) & 1;
      int64_t row_off[2];                                   // this lane stores chunk (lane & 1) of rows 16 t + lane / 2
      int row_pr[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int r2 = quarter * 32 + 16 * t + (lane >> 1);
        const int pr2 = a.q_swap ? (r2 / a.G) : (r2 & (P - 1));
        const int gr2 = a.q_swap ? (r2 & (a.G - 1)) : (r2 >> a.lgP);
        row_pr[t] = pr2;
        row_off[t] = static_cast<int64_t>(gr2) * a.sdq.h + static_cast<int64_t>(pr2) * a.sdq.n + (lane & 1) * 8;
      }
      int tc = 0;
      const bool tr = SFA_TRACE && (lane == 0) && (quarter == 0);
      const int trole = grp ? 6 : 4;
      FusedWalk w(a, wi);
      SlotTrack st;
      st.slot0 = 0;
      int pa = 0;
      // ---- delta(n + 3) after the epilogue of tile n (i.e. the tiles of the OTHER group): O(n + 3) and dO(n + 3) land
      // right when this group has finished epilogue(n), a tile before the UMMAs of its next tile n + 2 complete and more
      // than a tile before pass 2(n + 3).  (delta(n + 2) in front of epilogue(n): the group sat in the wait for the
      // delta inputs when dQ(n) was ready -- dq_free, the ring drain and with them every issuer ran late: 171 us.)
      // This warp owns rows 32 * quarter .. + 31.
      // (the delta tile's coordinates come from two divisions per tile, not from a third walker: registers)
      int dit = (kLean || a.fuse_delta) ? (grp ^ 1) : (1 << 30);         // index of the group's next delta tile in this CTA's run
      // O rows arrive as 16-row blocks (2 KB, one TMA box each); the warp that has consumed block j of tile k loads block j
      // of tile k + 1 into the same place -- for the same quarter's warp of the other group.  (One 16 KB box per tile,
      // loaded when the whole group had finished: delta(k) -> load -> delta(k + 1) was a chain of ~4600 + ~2000 cycles
      // per tile, longer than the tile period.)
      // (O(0) comes from the producer warp.  Code size matters here: the roles of this kernel run concurrently out of
      // one instruction cache -- "no instruction" is among its top stall reasons -- so the loops below are rolled and
      // the tile coordinates take two divisions, not seven.)
      auto delta_tile = [&]() {
        if (wi.tile + dit >= wi.end) return;
        if (tr) ftrace(a.trace, trole, tc, 7, dit);
        const int s = dit & 1;
        // coordinates first (integer divisions), under the wait for the inputs
        const int dhead = (wi.tile + dit) / a.nblk, dpb = wi.tile + dit - dhead * a.nblk;   // dhead = b * Hkv + y
        const int dbt = dhead / a.Hkv, dy = dhead - dbt * a.Hkv;
        const int pmax = a.N - 1 - dpb * P;
        // ds_aux: this row's lse and its head's s_aux, requested now, used after the row sum
        float lse_r = -INFINITY, sx = 0.f;
        const int my_pr = a.q_swap ? (lane >> (7 - a.lgP)) + quarter * (32 >> (7 - a.lgP)) : ((quarter * 32 + lane) & (P - 1));
        const int my_gr = a.q_swap ? (lane & (a.G - 1)) : ((quarter * 32 + lane) >> a.lgP);
        if (a.write_delta) {
          if (my_pr <= pmax)
            lse_r = __ldg(a.lse + (static_cast<int64_t>(dhead) * a.G + my_gr) * a.N + dpb * P + my_pr);
          sx = __ldg(a.s_aux + dy * a.G + my_gr);
        }
        int on = -1, ohb = 0;                                 // next tile (lane 0 loads its O blocks): position,
                                                              // head | batch << 12 (registers)
        if (lane == 0 && wi.tile + dit + 1 < wi.end) {
          const bool wrap = dpb + 1 == a.nblk, wrap_b = wrap && dy + 1 == a.Hkv;
          on = wrap ? 0 : (dpb + 1) * P;
          ohb = ((wrap_b ? 0 : dy + (wrap ? 1 : 0)) * a.G) | ((dbt + (wrap_b ? 1 : 0)) << 12);
        }
        const uint32_t ob = smem_u32(o_s), dob = smem_u32(do_s) + s * C::kQBytes;
        float* const dl = delta_s + s * 128;
        mbar_wait(do_full + s, (dit >> 1) & 1);               // dO tile
        if (dit >= 2) mbar_wait(delta_free + s, ((dit - 2) >> 1) & 1);
        // One lane per row (row 32 * quarter + lane; the 128B swizzle keeps the eight rows of a quarter-warp on distinct
        // banks): 16 LDS.128 + 1 STS per warp and no shuffles.  The math warps keep the MIO queue (LDS / STS / SHFL /
        // MUFU / tcgen05.ld) busy, and every MIO instruction of this role waits in it: 8 lanes per row with three
        // shuffles per row group (48 MIO instructions per warp) took ~4600 cycles per tile; warp-level MMAs
        // (delta = diag(O dO^T), mma.sync m16n8k16) queue in the tensor pipe between the UMMAs, which is what bounds
        // this kernel: +1500 cycles per tile.
        mbar_wait(o_full + (2 * quarter) * 2 + s, (dit >> 1) & 1);
        mbar_wait(o_full + (2 * quarter + 1) * 2 + s, (dit >> 1) & 1);
        if (tr) ftrace(a.trace, trole, tc, 8, dit);
        {
          const int r2 = quarter * 32 + lane;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
          for (int c = 0; c < 8; c += 1) {
            uint4 ov[1], gv[1];
#pragma unroll
            for (int u = 0; u < 1; ++u) {
              const uint32_t off = sw128_off(r2, c + u);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(ov[u].x), "=r"(ov[u].y), "=r"(ov[u].z), "=r"(ov[u].w) : "r"(ob + off) : "memory");
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(gv[u].x), "=r"(gv[u].y), "=r"(gv[u].z), "=r"(gv[u].w) : "r"(dob + off) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 1; ++u) {
              const uint32_t ou[4] = {ov[u].x, ov[u].y, ov[u].z, ov[u].w}, gu[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float o0, o1, g0, g1;
                unpack16f<T>(ou[e], o0, o1);
                unpack16f<T>(gu[e], g0, g1);
                s0 = fmaf(o0, g0, s0);
                s1 = fmaf(o1, g1, s1);
              }
            }
          }
          const float sum = s0 + s1;
          dl[r2] = sum;
          if (a.write_delta) {
            // the rows of one head in this warp: P consecutive lanes (P = 16: two heads per warp), or -- swapped layout --
            // the lanes with equal lane % G; xor-shuffle tree over them, result in the lowest lane of each head
            float c = (lse_r == -INFINITY) ? 0.f : -__expf(sx - lse_r) * sum;
#pragma unroll 1
            for (int o = a.q_swap ? a.G : 1, hi = a.q_swap ? 32 : min(P, 32); o < hi; o <<= 1)
              c += __shfl_xor_sync(0xffffffffu, c, o);
            const bool first = a.q_swap ? (lane < a.G) : ((lane & (min(P, 32) - 1)) == 0);
            // slot (head, tile) -- in the swapped layout every quarter holds rows of every head: (head, tile, quarter)
            const int nslot = a.q_swap ? 4 : 1;
            if (first)
              a.delta[((static_cast<int64_t>(dhead) * a.G + my_gr) * a.nblk + dpb) * nslot + (a.q_swap ? quarter : 0)] = c;
          }
        }
        __syncwarp();                                         // every lane has read its O row
        if (on >= 0) {
#pragma unroll 1
          for (int j = 2 * quarter; j < 2 * quarter + 2; ++j) {
            // block j of the packed tile: one head x 16 positions, or (swapped layout) 16 / G positions x G heads
            const int n0 = on + (a.q_swap ? ((16 * j) >> (7 - a.lgP)) : ((16 * j) & (P - 1)));
            const int h0 = (ohb & 4095) + (a.q_swap ? 0 : ((16 * j) >> a.lgP));
            uint64_t* const bar = o_full + j * 2 + (s ^ 1);
            mbar_expect_tx(bar, 2048);
            tma_load_4d(o_s + j * 2048, &tmO, bar, 0, a.q_swap ? h0 : n0, a.q_swap ? n0 : h0, ohb >> 12);
          }
        }
        if (lane == 0) {                                      // this warp's rows are in delta_s, its reads of dO are done
          mbar_arrive(delta_ready + s);
          mbar_arrive(do_empty + s);
        }
        if (tr) ftrace(a.trace, trole, tc, 6, dit);
        dit += 2;
      };
      // lead-in rounds: delta(0), delta(2) by group 1, delta(1) by group 0; then every round is epilogue(n), delta(n + 3)
      int lead = (kLean || a.fuse_delta) ? 1 + grp : 0;
      while (true) {
        if (lead > 0) {
          --lead;
        } else {
        bool ok;
        do {
          ok = w.next();
          if (ok) {
            st.step(w, R);
            if (w.seg_first) pa = w.pb;
          }
        } while (ok && (w.it & 1) != grp);
        if (!ok) break;
        if (tr) ftrace(a.trace, trole, tc, 1, w.it);
        mbar_wait_warp(dq_done + grp, (w.it >> 1) & 1);
        tc_fence_after();
        if (tr) ftrace(a.trace, trole, tc, 2, w.it);
        {
          T* dq_base = static_cast<T*>(a.dq);
          int il0 = w.pb * P;
          if (!kLean && a.dq_seg_n > 0) {       // straight into the sequence owner's buffer over NVLink
            const int seg = il0 / a.dq_seg_n;
            dq_base = static_cast<T*>(a.dq_peer[seg]);
            il0 -= seg * a.dq_seg_n;
          }
          T* const tile_dq = dq_base + static_cast<int64_t>(w.b) * a.sdq.b +
                             static_cast<int64_t>(w.y * a.G) * a.sdq.h + static_cast<int64_t>(il0) * a.sdq.n;
          const bool ok0 = w.pb * P + row_pr[0] < a.N, ok1 = w.pb * P + row_pr[1] < a.N;
          if (!kLean && a.dbg_delay && grp == 0) __nanosleep(a.dbg_delay * 4);
#pragma unroll 1
          for (int hq = 0; hq < 4; ++hq) {                  // channels 16 hq .. 16 hq + 15
            uint32_t x[16];
            tmem_ld16(tl + C::kColQ + hq * 16, x);
            tmem_ld_wait();
            if (hq == 3) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(dq_free + grp);
              if (tr) ftrace(a.trace, trole, tc, 3, w.it);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c)
              st_shared_v4(st_wr + ((c ^ st_key) << 4),
                           pack16<T>(__uint_as_float(x[8 * c + 0]), __uint_as_float(x[8 * c + 1])),
                           pack16<T>(__uint_as_float(x[8 * c + 2]), __uint_as_float(x[8 * c + 3])),
                           pack16<T>(__uint_as_float(x[8 * c + 4]), __uint_as_float(x[8 * c + 5])),
                           pack16<T>(__uint_as_float(x[8 * c + 6]), __uint_as_float(x[8 * c + 7])));
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int rr = 16 * t + (lane >> 1);
              uint4 val;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                           : "r"(stg + rr * 32 + (((lane & 1) ^ ((rr >> 2) & 1)) << 4))
                           : "memory");
              if (t == 0 ? ok0 : ok1) *reinterpret_cast<uint4*>(tile_dq + row_off[t] + hq * 16) = val;
            }
            __syncwarp();
          }
        }
        if (tr) ftrace(a.trace, trole, tc, 4, w.it);
        // ring drain: the block this tile touched last -- or, at the end of a segment, every live block (then
        // the whole ring is cleared with tcgen05.st: no ring UMMA is in flight until this drain is signalled)
        const bool last = w.seg_last();
        const bool tails = last && !w.seq_end();            // the run ends inside a sequence: blocks jb >= 1 are partial
        const int nd = last ? nb : 1;
        const int nh = P >> 4;
#pragma unroll 1
        for (int jb = 0; jb < nd; ++jb) {
          const int j = w.pb - nb + 1 + jb;                 // key block in query-tile units; absolute block j + qb
          if (j + qb < 0) continue;
          int slot = st.slot0 + jb;
          if (slot >= R) slot -= R;
#pragma unroll 1
          for (int h = 0; h < nh; ++h) {
            uint32_t x[16];
            tmem_ld16(tl + C::kColR + slot * P + h * 16, x);
            tmem_ld_wait();
            // blocks before the first tile of the run are shared with the previous CTA -- unless the run starts a
            // sequence (pa == 0): then they are halo keys (qb > 0) that only this CTA touches
            const bool head = pa > 0 && j < pa, tail = tails && jb >= 1;
            if (head || tail) {
              const int idx = head ? (j - (pa - nb + 1)) : (jb - 1);
              float* dst = part_cta + (static_cast<size_t>(tail ? 1 : 0) * C::kPartKeys + idx * P + h * 16) * 128;
#pragma unroll
              for (int e = 0; e < 16; ++e) dst[e * 128] = __uint_as_float(x[e]);
            } else {
              const int key0 = (j + qb) * P + h * 16;
              T* dst = okv + static_cast<int64_t>(w.b) * skv.b + static_cast<int64_t>(w.y) * skv.h +
                       static_cast<int64_t>(key0) * skv.n + dch;
              const int nv = nkv - key0;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (e < nv) *dst = from_f<T>(__uint_as_float(x[e]));
                dst += skv.n;
              }
            }
          }
        }
        if (last) {
          // The clear below covers EVERY slot, also the one the other epilogue group may still be reading for tile
          // n - 1 (the block that left the window with it): wait for that drain first.  (The groups only meet here:
          // everywhere else a slot is recycled by the issuers, which wait for its drain themselves.)
          if (w.it >= 1) mbar_wait_warp(drain_done + (grp ^ 1), ((w.it - 1) >> 1) & 1);
          uint32_t z[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) z[e] = 0u;
          for (int c = 0; c < R * P; c += 16) tmem_st16(tl + C::kColR + c, z);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(drain_done + grp);
        if (tr) ftrace(a.trace, trole, tc, 5, w.it);
        }
        delta_tile();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 21) tmem_dealloc(tmem, C::kTmemCols);
}

// Sums the fp32 partials of the key blocks shared by two neighbouring CTAs (tail of c - 1, head of c).
// grid (boundaries, float4 groups / 256): one float4 per thread -- every load is independent and in flight at once
// (the first version looped 64 dependent-latency iterations per thread and took 40 us for 2.4 MB).
// Blocks x >= nbound (y == 0) reduce ds_aux of head x - nbound instead (fixed order: deterministic;
// sink_flash_attention.py:653-665) -- the block partials of the preprocess pass, or the per-tile partials the fused
// kernel's epilogue groups left -- one launch less on the backward's critical path.
template <typename T>
__global__ void __launch_bounds__(256) bwd_fused_fixup_kernel(const FusedArgs a, const int vec_ok, const int nbound,
                                                              const float* __restrict__ ds_partial, float* ds_aux,
                                                              const int ds_nblk) {
  using C = FusedCfg;
  if (static_cast<int>(blockIdx.x) >= nbound) {
    if (blockIdx.y != 0) return;
    __shared__ float red[256];
    const int h = blockIdx.x - nbound;
    float s = 0.f;
    for (int b = 0; b < a.B; ++b)
      for (int t = threadIdx.x; t < ds_nblk; t += 256) s += ds_partial[(static_cast<int64_t>(b) * a.Hq + h) * ds_nblk + t];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) ds_aux[h] = red[0];
    return;
  }
  const int c = blockIdx.x + 1;
  const int t0 = c * a.tiles_per_cta;
  const int pa = t0 % a.nblk;
  if (pa == 0) return;                       // the boundary coincides with a sequence start: nothing shared
  // four float4 groups per thread, all eight loads in flight before the first add (one group per thread: 2 352 blocks in
  // two waves, one memory round trip each: 6.7 us for 19 MB)
  const int seq = t0 / a.nblk, y = seq % a.Hkv, b = seq / a.Hkv;
  const int nelem = (a.nb - 1) * a.P * 128;
  const float* const tails = a.part + (static_cast<size_t>(c - 1) * 2 + 1) * C::kPartKeys * 128;
  const float* const heads = a.part + (static_cast<size_t>(c) * 2 + 0) * C::kPartKeys * 128;
  float4 tl[4], hd[4];
  int ee[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ee[k] = ((blockIdx.y * 4 + k) * 256 + threadIdx.x) * 4;
    const int key = (pa + a.qb - a.nb + 1) * a.P + (ee[k] >> 7);
    if (ee[k] >= nelem || key < 0 || key >= a.Nkv) ee[k] = -1;
    if (ee[k] >= 0) {
      tl[k] = *reinterpret_cast<const float4*>(tails + ee[k]);
      hd[k] = *reinterpret_cast<const float4*>(heads + ee[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = ee[k];
    if (e < 0) continue;
    const int key = (pa + a.qb - a.nb + 1) * a.P + (e >> 7);
    const int which = (e >> 6) & 1, d = e & 63;
    T* o = static_cast<T*>(which ? a.dk : a.dv);
    const Strides4& s = which ? a.sdk : a.sdv;
    o += static_cast<int64_t>(b) * s.b + static_cast<int64_t>(y) * s.h + static_cast<int64_t>(key) * s.n + d;
    const float v0 = tl[k].x + hd[k].x, v1 = tl[k].y + hd[k].y, v2 = tl[k].z + hd[k].z, v3 = tl[k].w + hd[k].w;
    if (vec_ok) {
      *reinterpret_cast<uint2*>(o) = make_uint2(pack16<T>(v0, v1), pack16<T>(v2, v3));
    } else {
      o[0] = from_f<T>(v0);
      o[1] = from_f<T>(v1);
      o[2] = from_f<T>(v2);
      o[3] = from_f<T>(v3);
    }
  }
}

int fused_sm_count() { return device_sm_count(); }

// geometry of the fused path; false when the problem does not fit it
bool fused_geometry(const AttnParams& p, int& G, int& P, int& nb) {
  if (p.D != 64 || p.S != 0 || p.W < 1 || p.N < 1) return false;
  pick_packing(p.Hq, p.Hkv, G, P);
  if ((p.Hq / p.Hkv) != G) return false;              // one packed tile must hold the whole GQA group
  if (P != 16 && P != 32) return false;
  if (p.Hq > 4096 || p.B >= (1 << 19)) return false;   // the O-tile loader packs head | batch << 12 into one register
  if (p.q_off % P != 0) return false;                 // the key-block ring advances in whole query tiles
  const int64_t weff = p.W < p.Nkv ? p.W : p.Nkv;
  const int64_t nb64 = (weff - 1 + P - 1) / P + 1;
  if (nb64 * P > FusedCfg::kColsMax || (nb64 + 1) * P > FusedCfg::kRingCols) return false;
  nb = static_cast<int>(nb64);
  if ((nb - 1) * P > FusedCfg::kPartKeys) return false;
  return true;
}

template <typename T>
cudaError_t launch_fused(const AttnParams& p, int dtype, float* part, const float* ds_partial, int ds_nblk,
                         cudaStream_t st) {
  // with the in-kernel delta: ds_partial == p.delta (the kernel leaves per-tile partials there) when ds_aux is wanted
  using C = FusedCfg;
  int G, P, nb;
  if (!fused_geometry(p, G, P, nb)) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 1, sfa_last_error()); return cudaErrorInvalidValue; }
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
  if (!make_tile_map(&mq, p.q, dtype, C::D, p.N, p.Hq, p.B, p.sq, P, G)) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 2, sfa_last_error()); return cudaErrorInvalidValue; }
  if (!make_tile_map(&mdo, p.dout, dtype, C::D, p.N, p.Hq, p.B, p.sdo, P, G)) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 3, sfa_last_error()); return cudaErrorInvalidValue; }
  if (!make_tile_map(&mk, p.k, dtype, C::D, p.Nkv, p.Hkv, p.B, p.sk, a.cols, 1)) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 4, sfa_last_error()); return cudaErrorInvalidValue; }
  if (!make_tile_map(&mv, p.v, dtype, C::D, p.Nkv, p.Hkv, p.B, p.sv, a.cols, 1)) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 5, sfa_last_error()); return cudaErrorInvalidValue; }
  if (mq.swap_nh != mdo.swap_nh) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 6, sfa_last_error()); return cudaErrorInvalidValue; }
  a.fuse_delta = tc_bwd_fused_computes_delta(p) ? 1 : 0;
  TileMap mo = mq;
  if (a.fuse_delta) {
    // 16-row blocks of the packed tile: one head x 16 positions, or (swapped layout) 16 / G positions x G heads
    const bool o_swap = (p.Hq > 1 && p.N > 1) ? (p.so.h < p.so.n) : false;
    if (!make_tile_map(&mo, p.o, dtype, C::D, p.N, p.Hq, p.B, p.so, o_swap ? 16 / G : 16, o_swap ? G : 1))
      { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 7, sfa_last_error()); return cudaErrorInvalidValue; }
    if (mo.swap_nh != mq.swap_nh) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 8, sfa_last_error()); return cudaErrorInvalidValue; }
  }
  a.q_swap = mq.swap_nh; a.k_swap = mk.swap_nh; a.v_swap = mv.swap_nh;
  a.fmt = (dtype == SFA_DTYPE_BF16) ? 1 : 0;
  a.write_delta = (a.fuse_delta && ds_partial != nullptr) ? 1 : 0;
  a.s_aux = p.s_aux;
  if (a.write_delta) ds_nblk = a.nblk * (a.q_swap ? 4 : 1);
  a.qb = p.q_off / P; a.Nkv = p.Nkv; a.seq_lo = p.seq_lo; a.seq_bs = p.seq_bs;
  a.dbg_delay = debug_knob(0);
  a.dbg_norace = debug_knob(1);
  a.sl2 = p.scale * kLog2e;
  a.scale = p.scale;
  a.lse = p.lse;
  a.delta = p.delta;
  a.k = p.k; a.v = p.v; a.sk = p.sk; a.sv = p.sv;
  a.dq = p.dq; a.dk = p.dk; a.dv = p.dv;
  a.sdq = p.sdq; a.sdk = p.sdk; a.sdv = p.sdv;
  a.dq_seg_n = 0;
  for (int r = 0; r < 8; ++r) a.dq_peer[r] = nullptr;
  if (p.dq_route != nullptr) {
    const SpRoute& rt = *p.dq_route;
    if (rt.P < 1 || rt.P > 8 || rt.n_local % P != 0 || static_cast<int64_t>(rt.n_local) * rt.P != p.N)
      { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 9, sfa_last_error()); return cudaErrorInvalidValue; }
    a.sdq = Strides4{static_cast<int64_t>(rt.n_local) * rt.heads_total * C::D, C::D, static_cast<int64_t>(rt.heads_total) * C::D};
    for (int r = 0; r < rt.P; ++r) {
      if (rt.peer[r] == nullptr || reinterpret_cast<uintptr_t>(rt.peer[r]) % 16 != 0) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: reject #%d: %s\n", 10, sfa_last_error()); return cudaErrorInvalidValue; }
      a.dq_peer[r] = static_cast<char*>(rt.peer[r]) + static_cast<int64_t>(rt.head_off) * C::D * 2;
    }
    a.dq_seg_n = rt.n_local;
  }
  a.part = part;
  a.trace = trace_buffer();
  cudaError_t e;
  const bool lean = a.fuse_delta && a.dq_seg_n == 0 && a.dbg_delay == 0 && a.dbg_norace == 0 && a.trace == nullptr;
#define SFA_FUSED_LAUNCH(E_, L_)                                                                                       \
  {                                                                                                                    \
    static std::atomic<unsigned long long> attr_done{0};                                                               \
    if ((e = ensure_dyn_smem(bwd_fused64_kernel<T, E_, L_>, C::kSmem, attr_done)) != cudaSuccess) return e;            \
    bwd_fused64_kernel<T, E_, L_><<<grid, C::kThreads, C::kSmem, st>>>(mq.map, mdo.map, mk.map, mv.map, mo.map, a);    \
  }
  if (p.has_ext()) { if (lean) SFA_FUSED_LAUNCH(true, true) else SFA_FUSED_LAUNCH(true, false) }
  else { if (lean) SFA_FUSED_LAUNCH(false, true) else SFA_FUSED_LAUNCH(false, false) }
#undef SFA_FUSED_LAUNCH
  e = cudaGetLastError();
  if (e != cudaSuccess) { if (getenv("SFA_DEBUG_LAUNCH")) fprintf(stderr, "launch_fused: kernel launch error %d grid %d\n", (int)e, grid); return e; }
  const int nbound = (grid > 1 && nb > 1) ? grid - 1 : 0;   // nb == 1 (window <= one block): no shared key block
  const int nred = (ds_partial != nullptr && ds_nblk > 0) ? p.Hq : 0;
  if (nbound + nred > 0) {
    auto al8 = [](const void* ptr, const Strides4& sd) {
      return reinterpret_cast<uintptr_t>(ptr) % 8 == 0 && sd.n % 4 == 0 && sd.h % 4 == 0 && sd.b % 4 == 0;
    };
    const int vec_ok = al8(p.dk, p.sdk) && al8(p.dv, p.sdv);
    const int groups = nbound ? (nb - 1) * P * 128 / 4 : 1;
    bwd_fused_fixup_kernel<T><<<dim3(nbound + nred, (groups + 1023) / 1024), 256, 0, st>>>(
        a, vec_ok, nbound, ds_partial, p.ds_aux, ds_nblk);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace

// delta inside the kernel (three delta warps; the round-1 attempt on the epilogue groups, 8 rows in flight per warp,
// cost 152 us against 98 + 26 us).  SFA_FUSED_DELTA=0 goes back to the separate preprocess pass.
bool tc_bwd_fused_computes_delta(const AttnParams& p) {
  static const bool on = !(getenv("SFA_FUSED_DELTA") != nullptr && atoi(getenv("SFA_FUSED_DELTA")) == 0);
  if (!on || !tma_compatible(p.o, p.so, p.B, p.Hq, p.N)) return false;
  const bool q_swap = (p.Hq > 1 && p.N > 1) ? (p.sq.h < p.sq.n) : false;
  const bool o_swap = (p.Hq > 1 && p.N > 1) ? (p.so.h < p.so.n) : false;
  return q_swap == o_swap;
}

size_t tc_bwd_fused_workspace_bytes() {
  return static_cast<size_t>(FusedCfg::kMaxCtas) * 2 * FusedCfg::kPartKeys * 128 * sizeof(float);
}

bool tc_bwd_fused_supported(const AttnParams& p, int dtype) {
  static const bool disabled = getenv("SFA_NO_FUSED_BWD") != nullptr;
  if (disabled) return false;
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16) return false;
  int G, P, nb;
  if (!fused_geometry(p, G, P, nb)) return false;
  if (p.has_ext() && p.dq_route != nullptr) return false;
  if (!(tma_compatible(p.q, p.sq, p.B, p.Hq, p.N) && tma_compatible(p.k, p.sk, p.B, p.Hkv, p.Nkv) && tma_compatible(p.v, p.sv, p.B, p.Hkv, p.Nkv) &&
        tma_compatible(p.dout, p.sdo, p.B, p.Hq, p.N)))
    return false;
  // dQ rows are written with 16-byte stores; O and dO rows are read with 16-byte loads (delta)
  if (p.dq_route == nullptr &&
      (reinterpret_cast<uintptr_t>(p.dq) % 16 || p.sdq.n % 8 || p.sdq.h % 8 || p.sdq.b % 8)) return false;
  if (p.dq_route != nullptr && (P > 0) && (p.dq_route->n_local % P != 0)) return false;
  if (reinterpret_cast<uintptr_t>(p.o) % 16 || p.so.n % 8 || p.so.h % 8 || p.so.b % 8) return false;
  if (reinterpret_cast<uintptr_t>(p.dout) % 16 || p.sdo.n % 8 || p.sdo.h % 8 || p.sdo.b % 8) return false;
  const bool q_swap = (p.Hq > 1 && p.N > 1) ? (p.sq.h < p.sq.n) : false;
  const bool do_swap = (p.Hq > 1 && p.N > 1) ? (p.sdo.h < p.sdo.n) : false;
  return q_swap == do_swap;
}

cudaError_t tc_bwd_fused(const AttnParams& p, int dtype, float* part, const float* ds_partial, int ds_nblk,
                         cudaStream_t st) {
  if (dtype == SFA_DTYPE_BF16) return launch_fused<__nv_bfloat16>(p, dtype, part, ds_partial, ds_nblk, st);
  return launch_fused<__half>(p, dtype, part, ds_partial, ds_nblk, st);
}

}  // namespace sfa
