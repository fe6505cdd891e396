// Shared declarations for libsinkfa: problem descriptors, dtype helpers, error plumbing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>

#include "../../include/sinkfa.h"

namespace sfa {

struct Strides4 {  // element strides for (batch, head, position); channel stride is 1
  int64_t b, h, n;
};

// Sequence-parallel output routing (Ulysses): instead of (fwd: in addition to) the local tensor, rows of position i
// go to peer[i / n_local], a [B, n_local, heads_total, D] contiguous buffer of that rank, as head head_off + h.
struct SpRoute {
  int P, n_local, heads_total, head_off;
  void* peer[8];
};

// Prefill / training problem.  Mask: sink_flash_attention.py:30-39 of the reference.
struct AttnParams {
  const void* q;
  const void* k;
  const void* v;
  void* o;          // fwd: output; bwd: input
  float* lse;       // [B,Hq,N]
  const float* s_aux;
  const void* dout;
  void* dq;
  void* dk;
  void* dv;
  float* ds_aux;
  float* delta;     // workspace [B,Hq,N]
  float* dsrow;     // workspace [B,Hq,N]: per-row ds_aux contributions when the dQ kernel computes delta itself
  float* kv_part;   // workspace: fp32 dK / dV partials of split sink-holding key tiles (dQ + dK/dV kernel pair), or nullptr
  size_t kv_part_bytes;
  Strides4 sq, sk, sv, so, sdo, sdq, sdk, sdv;
  int B, Hq, Hkv, N, D, S, W;      // N: number of QUERY rows (== keys unless Nkv says otherwise)
  float scale;
  const SpRoute* o_route;    // host pointers, nullptr = off; only the head_dim-64 tcgen05 kernels route
  const SpRoute* dq_route;
  // ---- extended geometry (sfa_fwd_ex / sfa_bwd_ex): packed (varlen) sequences and chunked prefill / halo keys.
  // Query row iq sits at absolute position i = iq + q_off of a key axis of length Nkv (k, v, dk, dv have Nkv rows);
  // it attends key j iff  lo(iq) <= j <= i  and  (j - lo(iq) < S  or  j >= i - W + 1),  lo = seq_lo ? seq_lo[b][iq] : 0
  // -- with q_off = 0, Nkv = N and no seq_lo this is the reference predicate (sink_flash_attention.py:30-39).
  const int* seq_lo;         // device [B or 1][N]: first key position of the sequence query row iq belongs to
  const int* seq_hi;         // device [B or 1][Nkv]: one past the last absolute QUERY position that may attend key j
  int64_t seq_bs;            // elements between the batch rows of seq_lo / seq_hi (0: one row shared by the batch)
  int Nkv, q_off;
  __host__ __device__ bool has_ext() const { return seq_lo != nullptr || q_off != 0 || Nkv != N; }
};

struct DecodeParams {
  const void* q;
  const void* k[2];   // up to two KV segments (sink buffer, window buffer); segment 1 may be empty
  const void* v[2];
  int len[2];
  Strides4 sk[2];     // (b, h, n)
  Strides4 sv[2];
  void* o;
  const float* s_aux;
  int64_t sq_b, sq_h, so_b, so_h;
  int B, Hq, Hkv, D;
  float scale;
  float* part_ml;     // workspace: [B,Hq,splits,2]
  float* part_o;      // workspace: [B,Hq,splits,D]
  int splits;
  // per-batch cache lengths and paged KV (sfa_decode_paged; the reference shares one length across the batch and keeps
  // the cache contiguous, cache.py:11-13): k[0] / v[0] are the page pools, sk[0] / sv[0] = (page, head, position) strides,
  // len[0] = the planning length (max_len rounded up to whole pages), len[1] = 0.
  int paged;                 // 1: this launch uses the fields below
  const int* block_table;    // [B][bt_stride] physical page of logical page j of batch row b; nullptr: contiguous cache
                             // (page == batch row, sk[0].b = batch stride)
  int64_t bt_stride;
  const int* seq_lens;       // [B] keys cached for batch row b (device); nullptr: len[0] for every row
  int page_size, lg_page;    // keys per page (a power of two, a multiple of 32); contiguous cache: unused
};

void set_error(const char* fmt, ...);
void set_impl_name(const char* name);
int debug_knob(int which);     // test / diagnostics knobs set through sfa_set_debug (0 when unset)

// Per-DEVICE one-time state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to a device,
// not to the process: a second GPU in the same process (device_map="auto", autograd's per-device threads) must get
// its own attribute call.  One bit per device ordinal; devices >= 64 simply repeat the (idempotent) call.
template <class K>
inline cudaError_t ensure_dyn_smem(K* kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  return e;
}
int device_sm_count();         // SM count of the CURRENT device (cached per device ordinal)

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// attended-set predicate, shared by every kernel (reference: sink_flash_attention.py:36-39)
__host__ __device__ __forceinline__ bool attended(int i, int j, int S, int W) {
  return (j <= i) && ((j < S) || (j >= i - W + 1));
}

// launchers implemented in the .cu files; every one returns a cudaError_t
cudaError_t simt_fwd(const AttnParams& p, int dtype, cudaStream_t st);
cudaError_t simt_bwd(const AttnParams& p, int dtype, int stages, cudaStream_t st);
cudaError_t simt_decode(const DecodeParams& p, int dtype, cudaStream_t st);
// defer_reduce_nblk != nullptr: do not launch the ds_aux reduce; *defer_reduce_nblk = partials per (b, head)
cudaError_t bwd_preprocess(const AttnParams& p, int dtype, float* ds_partial, cudaStream_t st,
                           int* defer_reduce_nblk = nullptr);

bool tc_fwd_supported(const AttnParams& p, int dtype);
cudaError_t tc_fwd(const AttnParams& p, int dtype, cudaStream_t st);
bool tc_fwd64_supported(const AttnParams& p, int dtype);   // persistent warp-specialised forward, head_dim 64
cudaError_t tc_fwd64(const AttnParams& p, int dtype, cudaStream_t st);
bool tc_fwd64_route_supported(const AttnParams& p, int dtype);
bool tc_fwd128_supported(const AttnParams& p, int dtype);  // persistent two-tile forward, 64 < head_dim <= 128
cudaError_t tc_fwd128(const AttnParams& p, int dtype, cudaStream_t st);
bool tc_bwd_supported(const AttnParams& p, int dtype);
bool tc_bwd_fuses_delta(const AttnParams& p, int dtype);   // the dQ kernel derives delta (and ds_aux rows) itself: no preprocess pass
cudaError_t ds_aux_reduce(const float* partial, float* ds_aux, int B, int Hq, int nblk, cudaStream_t st);
bool tc_bwd_fused_computes_delta(const AttnParams& p);   // delta by the fused kernel's own delta warps
cudaError_t ds_aux_from_delta(const float* delta, const float* lse, const float* s_aux, float* ds_aux, int B, int Hq,
                              int N, cudaStream_t st);
cudaError_t tc_bwd(const AttnParams& p, int dtype, int stages, cudaStream_t st);
// fused dQ + dK + dV kernel for narrow windows without sink tokens (bwdf_sm100.cu); `part` = fp32 partials workspace
bool tc_bwd_fused_supported(const AttnParams& p, int dtype);
size_t tc_bwd_fused_workspace_bytes();
// ds_partial / ds_nblk: block partials of ds_aux left by bwd_preprocess, reduced by extra blocks of the fix-up launch
// (nullptr / 0: nothing to reduce)
cudaError_t tc_bwd_fused(const AttnParams& p, int dtype, float* part, const float* ds_partial, int ds_nblk,
                         cudaStream_t st);

// one decoded token's K and V rows -> ring slot write_pos of the window buffers (cache_ops.cu)
cudaError_t cache_append(const void* k_new, const void* v_new, void* win_k, void* win_v, int B, int H, int D,
                         int elem_size, const int64_t new_strides[2], const int64_t win_strides[3], int write_pos,
                         cudaStream_t st);

bool mma_decode_supported(const DecodeParams& p, int dtype);
int mma_decode_splits(int B, int Hq, int Hkv, int total_len, int align = 1);   // workspace slots per q head (align: page size)
cudaError_t mma_decode(const DecodeParams& p, int dtype, cudaStream_t st);

void set_trace_buffer(long long* p);   // performance-debug timeline (device buffer, 3*256*2 int64) or nullptr
long long* trace_buffer();

// Ulysses exchange by peer-memory stores (ulysses_p2p.cu)
cudaError_t ulysses_scatter(const void* src, void* const* peer_dst, int P, int rank, int mode, int B, int L, int H,
                            int D, int elem_size, const int64_t src_strides[3], int dst_heads, int head_off,
                            cudaStream_t st);
}  // namespace sfa
