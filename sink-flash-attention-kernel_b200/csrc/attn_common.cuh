// Device helpers shared by the tcgen05 forward and backward kernels: the two-range KV tile plan,
// 16-bit packing, and TMA tile loads that honour a per-tensor (position, head) dim order.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace sfa {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// KV tiles seen by the packed query tile [q0, q0+P): sink tiles over [0, min(S, q_hi+1)) first,
// then BN-row tiles over the causal band [max(q0-W+1, S, 0), q_hi].  Same walk as the reference's
// two ranges (sink_flash_attention.py:151-180) but with run-time, unaligned tile starts.
struct TilePlan {
  int n_sink, n_tiles, w_lo, q_hi, s_eff;
  __device__ __forceinline__ void tile(int t, int BN, int& kstart, int& cols, bool& is_sink) const {
    if (t < n_sink) {
      is_sink = true;
      kstart = t * BN;
      cols = min(BN, ((s_eff - kstart + 15) >> 4) << 4);
    } else {
      is_sink = false;
      kstart = w_lo + (t - n_sink) * BN;
      cols = min(BN, ((q_hi + 1 - kstart + 15) >> 4) << 4);
    }
  }
};

// x / BN without an integer division (a ~150-cycle dependent chain on the single-thread TMA / MMA roles):
// q = (x * ceil(2^40 / BN)) >> 40, exact for 0 <= x < 2^24 and BN <= 256.
__host__ __device__ __forceinline__ unsigned long long bn_magic(int BN) {
  return ((1ull << 40) + static_cast<unsigned long long>(BN) - 1) / static_cast<unsigned long long>(BN);
}
__device__ __forceinline__ int div_bn(int x, unsigned long long magic) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(x)) * magic) >> 40);
}

__device__ __forceinline__ TilePlan make_plan(int q0, int P, int N, int S, int W, int BN, unsigned long long magic) {
  TilePlan pl;
  pl.q_hi = min(q0 + P, N) - 1;
  pl.s_eff = min(S, pl.q_hi + 1);
  pl.n_sink = (pl.s_eff > 0) ? div_bn(pl.s_eff + BN - 1, magic) : 0;
  pl.w_lo = max(max(q0 - W + 1, S), 0);
  const int n_win = (W > 0 && pl.w_lo <= pl.q_hi) ? div_bn(pl.q_hi - pl.w_lo + BN, magic) : 0;
  pl.n_tiles = pl.n_sink + n_win;
  return pl;
}
__device__ __forceinline__ TilePlan make_plan(int q0, int P, int N, int S, int W, int BN) {
  return make_plan(q0, P, N, S, W, BN, bn_magic(BN));
}
// Packed (varlen) sequences without sink tokens: the band of a tile never starts before the first key of the
// sequence its FIRST row belongs to (sequence starts are non-decreasing along the rows), so whole KV tiles of earlier
// sequences are skipped, not just masked -- a full-attention layer over a packed batch stays O(sum len^2).
__device__ __forceinline__ void clamp_plan(TilePlan& pl, int seq_lo_q0, int W, int BN, unsigned long long magic) {
  pl.w_lo = max(pl.w_lo, seq_lo_q0);
  const int n_win = (W > 0 && pl.w_lo <= pl.q_hi) ? div_bn(pl.q_hi - pl.w_lo + BN, magic) : 0;
  pl.n_tiles = pl.n_sink + n_win;
}

// Persistent-kernel work walker.  Args type A provides N, S, W, P, BN, ny, nblk, total_tiles, tiles_per_cta,
// bn_mul, and the extended geometry q_off (absolute position of query row 0), seq_lo / seq_bs (packed sequences).  Tile id = (b * ny + y) * nblk + pb (pb: position block of P positions, y: packed head group).
template <class A>
__device__ __forceinline__ void decode_tile(const A& a, int tile, int& pb, int& y, int& b) {
  pb = tile % a.nblk;
  const int r = tile / a.nblk;
  y = r % a.ny;
  b = r / a.ny;
}

// Walks this CTA's work items: tiles blockIdx.x, +gridDim.x, ... (or a contiguous range) and the KV tiles of
// each, with the tile coordinates kept incrementally (no integer divisions per tile: a division is a
// ~150-cycle dependent chain on the single-thread TMA / UMMA roles).
// kExt = false: the extended geometry (chunk offset, packed-sequence bounds) is compiled out.
template <class A, bool kExt = true>
struct ItemWalkT {
  const A& a;
  int tile, it, t, n;        // tile id, tile iteration, KV tile inside the tile, running item count
  int step, end;
  int pb, y, b, q0;
  TilePlan pl;
  __device__ __forceinline__ explicit ItemWalkT(const A& a_) : a(a_), it(-1), t(0), n(-1) {
    int first;
    if (a.tiles_per_cta > 0) {
      step = 1;
      first = static_cast<int>(blockIdx.x) * a.tiles_per_cta;
      end = min(first + a.tiles_per_cta, a.total_tiles);
    } else {
      step = gridDim.x;
      first = static_cast<int>(blockIdx.x);
      end = a.total_tiles;
    }
    decode_tile(a, first, pb, y, b);     // the only divisions: once per role
    tile = first - step;
    pb -= step;                          // next() adds it back
    pl.n_tiles = 0;
    q0 = 0;
  }
  __device__ __forceinline__ static void advance(const A& a, int step, int& pb, int& y, int& b) {
    pb += step;
    while (pb >= a.nblk) {
      pb -= a.nblk;
      if (++y == a.ny) {
        y = 0;
        ++b;
      }
    }
  }
  __device__ __forceinline__ bool next() {
    ++t;
    while (t >= pl.n_tiles) {
      tile += step;
      ++it;
      if (tile >= end) return false;
      advance(a, step, pb, y, b);
      q0 = pb * a.P;
      // the plan lives in ABSOLUTE key positions: row iq of the tile sits at iq + q_off
      pl = make_plan(q0 + (kExt ? a.q_off : 0), a.P, a.N + (kExt ? a.q_off : 0), a.S, a.W, a.BN, a.bn_mul);
      if (kExt && a.seq_lo != nullptr) clamp_plan(pl, a.seq_lo[b * a.seq_bs + q0], a.W, a.BN, a.bn_mul);
      t = 0;
    }
    ++n;
    return true;
  }
  __device__ __forceinline__ bool last_of_tile() const { return t == pl.n_tiles - 1; }
  // Launches in which EVERY tile has exactly one KV item (no sink tokens, the band fits one tile -- the narrow-window
  // training shape): hop over `hop` tiles and plan only the one landed on.  Lets a role that owns every second tile
  // (the forward's ping-pong softmax groups) skip the other group's tile without paying for its plan; the item
  // counter stays in step with the roles that visit every tile (n == it).
  __device__ __forceinline__ bool next_single(int hop) {
    tile += step * hop;
    it += hop;
    n += hop;
    if (tile >= end) return false;
    advance(a, step * hop, pb, y, b);
    q0 = pb * a.P;
    pl = make_plan(q0 + (kExt ? a.q_off : 0), a.P, a.N + (kExt ? a.q_off : 0), a.S, a.W, a.BN, a.bn_mul);
    if (kExt && a.seq_lo != nullptr) clamp_plan(pl, a.seq_lo[b * a.seq_bs + q0], a.W, a.BN, a.bn_mul);
    t = 0;
    return true;
  }
};

// attended columns [c_lo, c_hi] of query position i inside a tile that starts at key `kstart`
__device__ __forceinline__ void row_range(bool is_sink, int i, int kstart, int cols, int S, int W, int& c_lo, int& c_hi) {
  if (is_sink) {
    c_lo = 0;
    c_hi = min(S, i + 1) - kstart - 1;
  } else {
    c_lo = max(max(i - W + 1, S) - kstart, 0);
    c_hi = i - kstart;
  }
  c_hi = min(c_hi, cols - 1);
}

// fp32 pair -> packed 16-bit pair (round to nearest even)
template <typename T> __device__ __forceinline__ uint32_t pack16(float a, float b);
template <> __device__ __forceinline__ uint32_t pack16<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack16<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Measured (tools/time_graph.py): replacing the cvt by integer rounding (IADD + PRMT) made every kernel SLOWER
// (dQ 94 -> 112 us): the math warps are bound by TMEM round-trip latency and issue slots, not by the XU pipe.
// ---- packed math (fewer issued instructions per element: the tcgen05 kernels are bound by instruction issue)
// two fp32 lanes in one 64-bit register; fma.rn.f32x2 is one FFMA2 on sm_100
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint32_t a_lo, uint32_t a_hi, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("{\n\t.reg .b64 ra;\n\tmov.b64 ra, {%1, %2};\n\tfma.rn.f32x2 %0, ra, %3, %4;\n\t}"
      : "=l"(d) : "r"(a_lo), "r"(a_hi), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// 16-bit pair product: a (packed pair) * b (packed pair), rounded once to the 16-bit type (HMUL2)
template <typename T> __device__ __forceinline__ uint32_t mul16x2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t mul16x2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <> __device__ __forceinline__ uint32_t mul16x2<__half>(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <typename T> __device__ __forceinline__ uint32_t pack16_fast(float a, float b) { return pack16<T>(a, b); }

__device__ __forceinline__ void tma_tile(void* dst, const CUtensorMap* m, uint64_t* bar, int swap, int d, int n, int h,
                                         int b) {
  if (swap) tma_load_4d(dst, m, bar, d, h, n, b);
  else tma_load_4d(dst, m, bar, d, n, h, b);
}
__device__ __forceinline__ void tma_tile_prefetch(const CUtensorMap* m, int swap, int d, int n, int h, int b) {
  if (swap) tma_prefetch_4d(m, d, h, n, b);
  else tma_prefetch_4d(m, d, n, h, b);
}
__device__ __forceinline__ void tma_tile_store(const CUtensorMap* m, const void* src, int swap, int d, int n, int h, int b) {
  if (swap) tma_store_4d(m, src, d, h, n, b);
  else tma_store_4d(m, src, d, n, h, b);
}

// GQA packing: G heads x P positions = 128 MMA rows (G = largest power of two dividing the group, <= 16)
inline void pick_packing(int Hq, int Hkv, int& G, int& P) {
  const int group = Hq / Hkv;
  G = 1;
  while (G < 16 && group % (G * 2) == 0) G *= 2;
  P = 128 / G;
}

// KV tile rows.  A band of min(W,N) + P - 1 keys that fits one tile gets a tile of exactly that size (rounded
// to the UMMA N granularity of 16); a wider band is cut into full bn_max tiles -- the last, partial tile of a
// row of tiles only issues UMMAs over its own (16-rounded) column count.
inline int pick_bn(int W, int N, int P, int bn_max) {
  int64_t span = (int64_t)(W < N ? W : N) + P - 1;
  if (span < 16) span = 16;
  if (span > (int64_t)N + P) span = (int64_t)N + P;
  if (span > bn_max) return bn_max;
  return (int)((span + 15) / 16 * 16);
}

}  // namespace sfa
