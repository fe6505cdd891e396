// Device-side append of one decoded token into the ring window buffers of a SinkCacheLayer (reference:
// SinkCacheLayer._decode, cache.py:129-147 -- two strided torch copies per layer and step, followed by the full
// linearisation of get_kv(), :185-216).  ONE launch writes the token's K and V rows of every (batch, kv head) into
// ring slot `write_pos`; together with sfa_decode_ring, which attends the sink and ring buffers in place, a decode
// step moves 2 * B * Hkv * D elements instead of copying the whole cache.
#include "common.cuh"

namespace sfa {
namespace {

struct AppendArgs {
  const char* k_new;
  const char* v_new;
  char* win_k;
  char* win_v;
  int64_t sn_b, sn_h;      // byte strides of the new rows (batch, head)
  int64_t sw_b, sw_h;      // byte strides of the ring buffers (batch, head)
  int64_t slot_off;        // write_pos * position stride, bytes
  int B, H, row_bytes, vec;
};

__global__ void __launch_bounds__(256) cache_append_kernel(const AppendArgs a) {
  const int per_row = a.vec ? a.row_bytes / 16 : a.row_bytes / 2;
  const int64_t total = static_cast<int64_t>(2) * a.B * a.H * per_row;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % per_row);
    int64_t r = idx / per_row;
    const int h = static_cast<int>(r % a.H);
    r /= a.H;
    const int b = static_cast<int>(r % a.B);
    const int which = static_cast<int>(r / a.B);
    const char* src = (which ? a.v_new : a.k_new) + b * a.sn_b + h * a.sn_h;
    char* dst = (which ? a.win_v : a.win_k) + b * a.sw_b + h * a.sw_h + a.slot_off;
    if (a.vec) reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
    else reinterpret_cast<uint16_t*>(dst)[c] = reinterpret_cast<const uint16_t*>(src)[c];
  }
}

}  // namespace

cudaError_t cache_append(const void* k_new, const void* v_new, void* win_k, void* win_v, int B, int H, int D,
                         int elem_size, const int64_t new_strides[2], const int64_t win_strides[3], int write_pos,
                         cudaStream_t st) {
  AppendArgs a;
  a.k_new = static_cast<const char*>(k_new);
  a.v_new = static_cast<const char*>(v_new);
  a.win_k = static_cast<char*>(win_k);
  a.win_v = static_cast<char*>(win_v);
  a.sn_b = new_strides[0] * elem_size; a.sn_h = new_strides[1] * elem_size;
  a.sw_b = win_strides[0] * elem_size; a.sw_h = win_strides[1] * elem_size;
  a.slot_off = static_cast<int64_t>(write_pos) * win_strides[2] * elem_size;
  a.B = B; a.H = H;
  a.row_bytes = D * elem_size;
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  a.vec = (a.row_bytes % 16 == 0 && al16(k_new) && al16(v_new) && al16(win_k) && al16(win_v) && a.sn_b % 16 == 0 &&
           a.sn_h % 16 == 0 && a.sw_b % 16 == 0 && a.sw_h % 16 == 0 && a.slot_off % 16 == 0) ? 1 : 0;
  if (!a.vec && a.row_bytes % 2 != 0) return cudaErrorInvalidValue;
  const int64_t total = static_cast<int64_t>(2) * B * H * (a.vec ? a.row_bytes / 16 : a.row_bytes / 2);
  if (total == 0) return cudaSuccess;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  cache_append_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace sfa
