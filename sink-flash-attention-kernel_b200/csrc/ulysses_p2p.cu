// Ulysses all-to-all as ONE peer-memory scatter kernel per tensor (replaces, for the sequence-parallel layout of
// BASELINE configs[4], the pack copy + NCCL all_to_all_single + unpack copy of sp_utils._QKVSeqToHead /
// _HeadToSeq; the reference leaves this exchange to verl, verl_patch.py:15-20).  Every rank reads its local
// tensor once with 16-byte loads and stores each row straight into the final position of the destination
// rank's receive buffer over NVLink (the buffers are CUDA peer mappings handed in as raw pointers: the C ABI
// does not know how they were exchanged -- the Python host uses torch symmetric memory).  A cross-rank barrier
// (host side, on the same stream) follows the launch.
//
//   mode 0, sequence -> heads:  src [B, n, H, D] (this rank's positions, all heads; any strides, unit channel
//           stride).  Head h goes to rank r = h / hl as local head h % hl:
//           dst_r [B, P*n, dst_heads, D] contiguous, row (b, rank*n + i, head_off + h % hl).
//   mode 1, heads -> sequence:  src [B, P*n, hl, D] (all positions, this rank's heads).  Position i goes to
//           rank s = i / n:  dst_s [B, n, dst_heads, D] contiguous, row (b, i % n, head_off + rank*hl + h).
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace sfa {
namespace {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  void* p[kMaxPeers];
};

struct ScatterArgs {
  const char* src;
  int64_t sb, sn, sh;      // source strides in BYTES for (batch, position, head)
  int B, L, H;             // source extents: batch, positions, heads
  int n, hl;               // positions per rank, heads per rank
  int P, rank, mode;
  int dst_heads, head_off;
  int row16;               // 16-byte chunks per row (D * elem_size / 16)
};

__device__ __forceinline__ char* scatter_dst(const ScatterArgs& a, const PeerPtrs& peers, int64_t idx, const char*& src) {
  const int c = static_cast<int>(idx % a.row16);
  int64_t t = idx / a.row16;
  const int h = static_cast<int>(t % a.H);
  t /= a.H;
  const int i = static_cast<int>(t % a.L);
  const int b = static_cast<int>(t / a.L);
  src = a.src + b * a.sb + i * a.sn + h * a.sh + c * 16;
  int dst_rank;
  int64_t drow;
  if (a.mode == 0) {
    dst_rank = h / a.hl;
    drow = (static_cast<int64_t>(b) * (a.P * a.n) + a.rank * a.n + i) * a.dst_heads + a.head_off + (h - dst_rank * a.hl);
  } else {
    dst_rank = i / a.n;
    drow = (static_cast<int64_t>(b) * a.n + (i - dst_rank * a.n)) * a.dst_heads + a.head_off + a.rank * a.hl + h;
  }
  return static_cast<char*>(peers.p[dst_rank]) + (drow * a.row16 + c) * 16;
}

// kScatterUnroll independent 16-byte loads in flight per thread before the first (possibly remote) store.
// Measured at N = 2 with both directions busy (tools/tune_scatter.py): 507-580 GB/s per direction for every grid
// size (4-32 blocks per SM) and unroll (1-8), and 560 GB/s for a cp.async.bulk variant moving 4 KB rows through
// shared memory -- the link, not the kernel, sets the rate, so the simple kernel stays.
template <int kScatterUnroll>
__global__ void __launch_bounds__(256) ulysses_scatter_kernel(const ScatterArgs a, const PeerPtrs peers) {
  const int64_t total = static_cast<int64_t>(a.B) * a.L * a.H * a.row16;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; idx + (kScatterUnroll - 1) * step < total; idx += kScatterUnroll * step) {
    uint4 v[kScatterUnroll];
    char* d[kScatterUnroll];
#pragma unroll
    for (int u = 0; u < kScatterUnroll; ++u) {
      const char* sp;
      d[u] = scatter_dst(a, peers, idx + u * step, sp);
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(sp));
    }
#pragma unroll
    for (int u = 0; u < kScatterUnroll; ++u) *reinterpret_cast<uint4*>(d[u]) = v[u];
  }
  for (; idx < total; idx += step) {
    const char* sp;
    char* d = scatter_dst(a, peers, idx, sp);
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(sp);
  }
}

}  // namespace

cudaError_t ulysses_scatter(const void* src, void* const* peer_dst, int P, int rank, int mode, int B, int L, int H,
                            int D, int elem_size, const int64_t src_strides[3], int dst_heads, int head_off,
                            cudaStream_t st) {
  if (P < 1 || P > kMaxPeers || rank < 0 || rank >= P || (mode != 0 && mode != 1)) return cudaErrorInvalidValue;
  if ((D * elem_size) % 16 != 0 || reinterpret_cast<uintptr_t>(src) % 16 != 0) return cudaErrorInvalidValue;
  for (int s = 0; s < 3; ++s)
    if ((src_strides[s] * elem_size) % 16 != 0) return cudaErrorInvalidValue;
  ScatterArgs a;
  a.src = static_cast<const char*>(src);
  a.sb = src_strides[0] * elem_size;
  a.sn = src_strides[1] * elem_size;
  a.sh = src_strides[2] * elem_size;
  a.B = B; a.L = L; a.H = H; a.P = P; a.rank = rank; a.mode = mode;
  if (mode == 0) {
    if (H % P != 0) return cudaErrorInvalidValue;
    a.n = L;
    a.hl = H / P;
  } else {
    if (L % P != 0) return cudaErrorInvalidValue;
    a.n = L / P;
    a.hl = H;
  }
  a.dst_heads = dst_heads;
  a.head_off = head_off;
  a.row16 = D * elem_size / 16;
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) pp.p[r] = r < P ? peer_dst[r] : nullptr;
  for (int r = 0; r < P; ++r)
    if (pp.p[r] == nullptr || reinterpret_cast<uintptr_t>(pp.p[r]) % 16 != 0) return cudaErrorInvalidValue;
  const int64_t total = static_cast<int64_t>(B) * L * H * a.row16;
  if (total == 0) return cudaSuccess;
  int64_t blocks = (total + 255) / 256;
  // tuning knobs (tools/dev_p2p.py): blocks per SM of the grid-stride loop, loads in flight per thread
  static const int env_blocks = getenv("SFA_SCATTER_BLOCKS") ? atoi(getenv("SFA_SCATTER_BLOCKS")) : 16;
  static const int env_unroll = getenv("SFA_SCATTER_UNROLL") ? atoi(getenv("SFA_SCATTER_UNROLL")) : 4;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * env_blocks;
  if (blocks > cap) blocks = cap;
  const int unroll = env_unroll;
  if (unroll >= 8) ulysses_scatter_kernel<8><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, pp);
  else if (unroll >= 4) ulysses_scatter_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, pp);
  else if (unroll >= 2) ulysses_scatter_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, pp);
  else ulysses_scatter_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, pp);
  return cudaGetLastError();
}

}  // namespace sfa
