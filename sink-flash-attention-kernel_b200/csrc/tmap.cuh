// Host-side construction of 4-D TMA tensor maps over [B,H,N,D]-shaped tensors with arbitrary
// (batch, head, position) element strides.  libcuda is not linked: cuTensorMapEncodeTiled is
// resolved at run time through cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>

#include "common.cuh"

namespace sfa {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct TileMap {
  CUtensorMap map;
  int swap_nh;  // 1: the map's dim order is (D, H, N, B) -- coordinates must be passed (d, h, n, b)
};

// Tensor viewed as (D, N, H, B) with element strides (1, sn, sh, sb); 16-bit elements.
// Box = 64 channels x box_n positions x box_h heads x 1 batch, SWIZZLE_128B, OOB -> zero.
// The two middle dims are ordered by increasing stride, so a multi-head box lands in shared
// memory as rows (h*box_n + n) when swap_nh == 0 and as rows (n*box_h + h) when swap_nh == 1.
// Encoded maps are cached per thread, keyed on everything the encoding depends on (base pointer, dtype, extents,
// strides, box): a training loop calls the operator with the same tensors' geometry (and, with the caching allocator,
// mostly the same addresses) every step, and each launch needs 4 - 12 maps -- the driver call per map was a visible
// part of the eager step (0.26 ms of host time around 0.18 ms of kernels in round 1).
struct TileMapKey {
  const void* ptr;
  int64_t sb, sh, sn;
  int dtype, D, N, H, B, box_n, box_h;
  bool operator==(const TileMapKey& o) const {
    return ptr == o.ptr && sb == o.sb && sh == o.sh && sn == o.sn && dtype == o.dtype && D == o.D && N == o.N && H == o.H &&
           B == o.B && box_n == o.box_n && box_h == o.box_h;
  }
};
struct TileMapCache {
  static constexpr int kSlots = 64;
  TileMapKey key[kSlots];
  TileMap val[kSlots];
  bool used[kSlots] = {};
  static unsigned slot_of(const TileMapKey& k) {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= static_cast<uint64_t>(k.box_n) * 0xC2B2AE3D27D4EB4Full + static_cast<uint64_t>(k.sn) * 0x165667B19E3779F9ull +
         static_cast<uint64_t>(k.N) * 31 + static_cast<uint64_t>(k.box_h);
    return static_cast<unsigned>(h >> 40) % kSlots;
  }
};

inline bool make_tile_map_uncached(TileMap* out, const void* ptr, int dtype, int D, int N, int H, int B, const Strides4& s,
                                   int box_n, int box_h);

inline bool make_tile_map(TileMap* out, const void* ptr, int dtype, int D, int N, int H, int B, const Strides4& s,
                          int box_n, int box_h) {
  static thread_local TileMapCache cache;
  const TileMapKey k{ptr, s.b, s.h, s.n, dtype, D, N, H, B, box_n, box_h};
  const unsigned slot = TileMapCache::slot_of(k);
  if (cache.used[slot] && cache.key[slot] == k) {
    *out = cache.val[slot];
    return true;
  }
  if (!make_tile_map_uncached(out, ptr, dtype, D, N, H, B, s, box_n, box_h)) return false;
  cache.key[slot] = k;
  cache.val[slot] = *out;
  cache.used[slot] = true;
  return true;
}

inline bool make_tile_map_uncached(TileMap* out, const void* ptr, int dtype, int D, int N, int H, int B, const Strides4& s,
                                   int box_n, int box_h) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  const CUtensorMapDataType dt = (dtype == SFA_DTYPE_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const bool swap = (H > 1 && N > 1) ? (s.h < s.n) : false;
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  dims[0] = (cuuint64_t)D;
  box[0] = 64;
  if (!swap) {
    dims[1] = N; dims[2] = H;
    strides[0] = (cuuint64_t)s.n * 2; strides[1] = (cuuint64_t)s.h * 2;
    box[1] = box_n; box[2] = box_h;
  } else {
    dims[1] = H; dims[2] = N;
    strides[0] = (cuuint64_t)s.h * 2; strides[1] = (cuuint64_t)s.n * 2;
    box[1] = box_h; box[2] = box_n;
  }
  dims[3] = B;
  strides[2] = (cuuint64_t)s.b * 2;
  box[3] = 1;
  // size-1 dims may carry arbitrary strides in torch (0 included); any 16-B multiple is acceptable for the encoder.
  // A zero stride on a dim of extent > 1 (expand()ed / broadcast tensor) is NOT patched: TMA cannot express it and
  // tma_compatible() keeps such tensors off this path.
  for (int i = 0; i < 3; ++i)
    if (dims[i + 1] == 1) strides[i] = (cuuint64_t)16 * ((dims[0] * 2 + 15) / 16);
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) {
      set_error("TMA tensor map: zero stride on a dimension of extent %llu (broadcast views need a copy)",
                (unsigned long long)dims[i + 1]);
      return false;
    }
  if (D < 64) box[0] = D;
  CUresult r = enc(&out->map, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    // The encoder is a DRIVER entry point: it fails on a thread that has not made a runtime call yet -- autograd's
    // backward thread, when the first thing the backward does is build a tensor map (it used to launch the delta
    // preprocess kernel first).  cudaFree(nullptr) binds the device's primary context to this thread; retried once.
    // (Not done up front: cudaFree is not allowed while a stream capture is running, and a capturing thread has a
    // context.)
    cudaFree(nullptr);
    r = enc(&out->map, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u,%u,%u)",
              (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
              (unsigned long long)strides[2], box[0], box[1], box[2], box[3]);
    return false;
  }
  out->swap_nh = swap ? 1 : 0;
  return true;
}

// TMA needs a 16-B aligned base and, on every dim of extent > 1, a non-zero 16-B multiple stride (a zero stride is
// what expand() produces: k.expand(B, ...), or dO = out.mean(dim=2) expanded back -- the tensor map cannot walk it;
// those tensors take the CUDA-core path, which honours the true strides).  Extent-1 dims carry no constraint.
inline bool tma_compatible(const void* ptr, const Strides4& s, int B, int H, int N) {
  auto ok = [](int64_t st, int extent) { return extent <= 1 || (st != 0 && st % 8 == 0); };
  return (reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && ok(s.n, N) && ok(s.h, H) && ok(s.b, B);
}

}  // namespace sfa
