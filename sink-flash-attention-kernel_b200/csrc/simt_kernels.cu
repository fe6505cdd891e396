// CUDA-core (fp32 math) sink-attention kernels: the fp32-I/O path of the library (the tcgen05
// tensor path is 16-bit only), the path for head dims the tensor kernels do not tile
// (D not in {64,128}), and the on-device cross-check used by the parity tests at sizes the
// CPU oracle cannot reach.  One warp owns one query row (fwd, dQ) or one key row (dK/dV);
// lanes split the 32 keys of a chunk for the scores and the channels for the accumulators.
//
// Reference semantics: sink_flash_attention.py:30-39 (mask), :139-146 (s_aux seed),
// :183-194 (normalise, LSE), :242-251 (P, dV, dP, dS, dK), :449,481 (dQ), :582 (delta),
// :653-665 (ds_aux); decode_kernel.py:201-226 (s_aux as a virtual split).
#include "common.cuh"

namespace sfa {
namespace {

constexpr int kMaxD = 256;
constexpr int kDPL = kMaxD / 32;  // channels per lane

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <typename T>
__device__ __forceinline__ float dot_row(const T* row, const float* __restrict__ qs, int D) {   // row: no __restrict__ (no LDG.NC: K / V may sit in a peer-written buffer)
  float s = 0.f;
  for (int d = 0; d < D; ++d) s = fmaf(to_f<T>(row[d]), qs[d], s);
  return s;
}

// ------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(128) simt_fwd_kernel(AttnParams p) {
  __shared__ float qs[4][kMaxD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= p.N) return;
  const int D = p.D, S = p.S, W = p.W;
  const int kvh = h / (p.Hq / p.Hkv);
  const T* q = static_cast<const T*>(p.q) + b * p.sq.b + h * p.sq.h + (int64_t)i * p.sq.n;
  const T* kb = static_cast<const T*>(p.k) + b * p.sk.b + kvh * p.sk.h;
  const T* vb = static_cast<const T*>(p.v) + b * p.sv.b + kvh * p.sv.h;
  for (int d = lane; d < D; d += 32) qs[warp][d] = to_f<T>(q[d]);
  __syncwarp();

  float m = p.s_aux ? p.s_aux[h] : -INFINITY;
  float l = p.s_aux ? 1.f : 0.f;
  float acc[kDPL];
#pragma unroll
  for (int t = 0; t < kDPL; ++t) acc[t] = 0.f;

  // two disjoint key ranges: sinks [s0, min(s0 + S, ia + 1)) and the window [max(s0 + S, ia - W + 1), ia], with
  // ia = i + q_off the row's absolute position and s0 the start of its (packed) sequence
  const int ia = i + p.q_off;
  const int s0 = p.seq_lo ? p.seq_lo[b * p.seq_bs + i] : 0;
  int lo[2] = {s0, max(s0 + S, ia - W + 1)};
  int hi[2] = {min(s0 + S, ia + 1), (W > 0) ? ia + 1 : 0};
  for (int r = 0; r < 2; ++r) {
    for (int j0 = lo[r]; j0 < hi[r]; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < hi[r];
      float s = valid ? dot_row<T>(kb + (int64_t)j * p.sk.n, qs[warp], D) * p.scale : -INFINITY;
      const float m_new = fmaxf(m, warp_max(s));
      const float alpha = (m == -INFINITY) ? 0.f : expf(m - m_new);
      const float pj = valid ? expf(s - m_new) : 0.f;
      l = l * alpha + warp_sum(pj);
#pragma unroll
      for (int t = 0; t < kDPL; ++t) acc[t] *= alpha;
      const int cnt = min(32, hi[r] - j0);
      for (int jj = 0; jj < cnt; ++jj) {
        const float pjj = __shfl_sync(0xffffffffu, pj, jj);
        const T* vr = vb + (int64_t)(j0 + jj) * p.sv.n;
#pragma unroll
        for (int t = 0; t < kDPL; ++t) {
          const int d = lane + 32 * t;
          if (d < D) acc[t] = fmaf(pjj, to_f<T>(vr[d]), acc[t]);
        }
      }
      m = m_new;
    }
  }
  T* o = static_cast<T*>(p.o) + b * p.so.b + h * p.so.h + (int64_t)i * p.so.n;
  const float inv = (l == 0.f) ? 0.f : 1.f / l;
#pragma unroll
  for (int t = 0; t < kDPL; ++t) {
    const int d = lane + 32 * t;
    if (d < D) o[d] = from_f<T>(acc[t] * inv);
  }
  if (lane == 0) p.lse[((int64_t)b * p.Hq + h) * p.N + i] = (l == 0.f) ? -INFINITY : m + logf(l);
}

// ------------------------------------------------------------------------------------ dQ
template <typename T>
__global__ void __launch_bounds__(128) simt_dq_kernel(AttnParams p) {
  __shared__ float qs[4][kMaxD];
  __shared__ float dos[4][kMaxD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= p.N) return;
  const int D = p.D, S = p.S, W = p.W;
  const int kvh = h / (p.Hq / p.Hkv);
  const T* q = static_cast<const T*>(p.q) + b * p.sq.b + h * p.sq.h + (int64_t)i * p.sq.n;
  const T* dO = static_cast<const T*>(p.dout) + b * p.sdo.b + h * p.sdo.h + (int64_t)i * p.sdo.n;
  const T* kb = static_cast<const T*>(p.k) + b * p.sk.b + kvh * p.sk.h;
  const T* vb = static_cast<const T*>(p.v) + b * p.sv.b + kvh * p.sv.h;
  for (int d = lane; d < D; d += 32) {
    qs[warp][d] = to_f<T>(q[d]);
    dos[warp][d] = to_f<T>(dO[d]);
  }
  __syncwarp();
  const int64_t row = ((int64_t)b * p.Hq + h) * p.N + i;
  const float lse = p.lse[row], delta = p.delta[row];
  float acc[kDPL];
#pragma unroll
  for (int t = 0; t < kDPL; ++t) acc[t] = 0.f;
  const int ia = i + p.q_off;
  const int s0 = p.seq_lo ? p.seq_lo[b * p.seq_bs + i] : 0;
  int lo[2] = {s0, max(s0 + S, ia - W + 1)};
  int hi[2] = {min(s0 + S, ia + 1), (W > 0) ? ia + 1 : 0};
  for (int r = 0; r < 2; ++r) {
    for (int j0 = lo[r]; j0 < hi[r]; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = (j < hi[r]) && (lse != -INFINITY);
      float ds = 0.f;
      if (valid) {
        const float s = dot_row<T>(kb + (int64_t)j * p.sk.n, qs[warp], D) * p.scale;
        const float pr = expf(s - lse);
        const float dp = dot_row<T>(vb + (int64_t)j * p.sv.n, dos[warp], D);
        ds = pr * (dp - delta);
      }
      const int cnt = min(32, hi[r] - j0);
      for (int jj = 0; jj < cnt; ++jj) {
        const float dsj = __shfl_sync(0xffffffffu, ds, jj);
        const T* kr = kb + (int64_t)(j0 + jj) * p.sk.n;
#pragma unroll
        for (int t = 0; t < kDPL; ++t) {
          const int d = lane + 32 * t;
          if (d < D) acc[t] = fmaf(dsj, to_f<T>(kr[d]), acc[t]);
        }
      }
    }
  }
  T* dq = static_cast<T*>(p.dq) + b * p.sdq.b + h * p.sdq.h + (int64_t)i * p.sdq.n;
#pragma unroll
  for (int t = 0; t < kDPL; ++t) {
    const int d = lane + 32 * t;
    if (d < D) dq[d] = from_f<T>(acc[t] * p.scale);
  }
}

// ------------------------------------------------------------------------------------ dK, dV
template <typename T>
__global__ void __launch_bounds__(128) simt_dkdv_kernel(AttnParams p) {
  __shared__ float ks[4][kMaxD];
  __shared__ float vs[4][kMaxD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + warp, kvh = blockIdx.y, b = blockIdx.z;
  if (j >= p.Nkv) return;
  const int D = p.D, S = p.S, W = p.W, N = p.N;
  const int g = p.Hq / p.Hkv;
  const T* kr = static_cast<const T*>(p.k) + b * p.sk.b + kvh * p.sk.h + (int64_t)j * p.sk.n;
  const T* vr = static_cast<const T*>(p.v) + b * p.sv.b + kvh * p.sv.h + (int64_t)j * p.sv.n;
  for (int d = lane; d < D; d += 32) {
    ks[warp][d] = to_f<T>(kr[d]);
    vs[warp][d] = to_f<T>(vr[d]);
  }
  __syncwarp();
  float dk[kDPL], dv[kDPL];
#pragma unroll
  for (int t = 0; t < kDPL; ++t) dk[t] = dv[t] = 0.f;
  // query rows (index iq, absolute position iq + q_off) that attend key j: position >= j, inside j's sequence, and
  // (j is one of the S sinks of that sequence or position <= j + W - 1).  Without sinks the band ends at j + W - 1;
  // with sinks every later row of the sequence is a candidate and the per-row predicate decides.
  const int qo = p.q_off;
  int e_abs = p.seq_hi ? p.seq_hi[b * p.seq_bs + j] : N + qo;            // one past the last candidate position
  bool maybe_sink = S > 0;
  if (maybe_sink) {
    if (p.seq_lo == nullptr) maybe_sink = j < S;
    else if (j >= qo && j - qo < N) maybe_sink = j - p.seq_lo[b * p.seq_bs + (j - qo)] < S;   // key j shares the sequence of the query at its position
  }
  if (!maybe_sink) e_abs = (W > 0) ? min(e_abs, j + W) : j;
  const int i_beg = max(j - qo, 0), i_end = min(e_abs - qo, N);
  for (int hh = 0; hh < g; ++hh) {
    const int h = kvh * g + hh;
    const T* qb = static_cast<const T*>(p.q) + b * p.sq.b + h * p.sq.h;
    const T* dob = static_cast<const T*>(p.dout) + b * p.sdo.b + h * p.sdo.h;
    const int64_t rowb = ((int64_t)b * p.Hq + h) * N;
    for (int i0 = i_beg; i0 < i_end; i0 += 32) {
      const int i = i0 + lane;
      float pr = 0.f, ds = 0.f;
      if (i < i_end) {
        const float lse = p.lse[rowb + i];
        const int s0 = p.seq_lo ? p.seq_lo[b * p.seq_bs + i] : 0;
        const bool att = j >= s0 && (j - s0 < S || j >= i + qo - W + 1);
        if (att && lse != -INFINITY) {
          const float s = dot_row<T>(qb + (int64_t)i * p.sq.n, ks[warp], D) * p.scale;
          pr = expf(s - lse);
          const float dp = dot_row<T>(dob + (int64_t)i * p.sdo.n, vs[warp], D);
          ds = pr * (dp - p.delta[rowb + i]);
        }
      }
      const int cnt = min(32, i_end - i0);
      for (int ii = 0; ii < cnt; ++ii) {
        const float pi = __shfl_sync(0xffffffffu, pr, ii);
        const float dsi = __shfl_sync(0xffffffffu, ds, ii);
        const T* qr = qb + (int64_t)(i0 + ii) * p.sq.n;
        const T* dor = dob + (int64_t)(i0 + ii) * p.sdo.n;
#pragma unroll
        for (int t = 0; t < kDPL; ++t) {
          const int d = lane + 32 * t;
          if (d < D) {
            dv[t] = fmaf(pi, to_f<T>(dor[d]), dv[t]);
            dk[t] = fmaf(dsi, to_f<T>(qr[d]), dk[t]);
          }
        }
      }
    }
  }
  T* dkr = static_cast<T*>(p.dk) + b * p.sdk.b + kvh * p.sdk.h + (int64_t)j * p.sdk.n;
  T* dvr = static_cast<T*>(p.dv) + b * p.sdv.b + kvh * p.sdv.h + (int64_t)j * p.sdv.n;
#pragma unroll
  for (int t = 0; t < kDPL; ++t) {
    const int d = lane + 32 * t;
    if (d < D) {
      dkr[d] = from_f<T>(dk[t] * p.scale);
      dvr[d] = from_f<T>(dv[t]);
    }
  }
}

// ------------------------------------------------------------------------------------ delta + ds_aux
// delta[b,h,i] = sum_d dO*O (sink_flash_attention.py:582); per-block partial of
// ds_aux[h] = -sum exp(s_aux[h]-lse)*delta (:653-665), reduced deterministically by a second kernel.
template <typename T>
__global__ void __launch_bounds__(256) preprocess_kernel(AttnParams p, float* ds_partial) {
  __shared__ float part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp, h = blockIdx.y, b = blockIdx.z;
  float contrib = 0.f;
  if (i < p.N) {
    const T* o = static_cast<const T*>(p.o) + b * p.so.b + h * p.so.h + (int64_t)i * p.so.n;
    const T* dO = static_cast<const T*>(p.dout) + b * p.sdo.b + h * p.sdo.h + (int64_t)i * p.sdo.n;
    float s = 0.f;
    for (int d = lane; d < p.D; d += 32) s = fmaf(to_f<T>(o[d]), to_f<T>(dO[d]), s);
    s = warp_sum(s);
    const int64_t row = ((int64_t)b * p.Hq + h) * p.N + i;
    if (lane == 0) p.delta[row] = s;
    if (p.s_aux) {
      const float lse = p.lse[row];
      contrib = (lse == -INFINITY) ? 0.f : -expf(p.s_aux[h] - lse) * s;
    }
  }
  if (ds_partial) {
    if (lane == 0) part[warp] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += part[w];
      ds_partial[((int64_t)b * p.Hq + h) * gridDim.x + blockIdx.x] = t;
    }
  }
}

// 16-bit fast path of the same: LPR = D/8 lanes per row, one 16-byte load of O and of dO per lane.
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

// Every thread owns one 16-byte slice of kRowsPerThread rows and issues all of its loads before the first use:
// 2 x kRowsPerThread x 16 B in flight per thread (the one-row version reached 52 % of the DRAM bandwidth with 32 B
// in flight per thread and 16 384 short-lived blocks).
constexpr int kPreRowsPerThread = 4;
// Streamed once: keep it out of L1.  NOT the non-coherent (.nc) path: under Ulysses dO sits in a buffer the peer GPUs
// write between launches, and .nc loads returned rows of the PREVIOUS step there (tools/dev_p2p.py, fresh data per
// round: delta, hence dQ and dK, came out stale on one rank while dV -- dO through TMA -- was right).
__device__ __forceinline__ uint4 ld_nc_v4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}
template <typename T, int LPR>
__global__ void __launch_bounds__(256) preprocess_vec_kernel(AttnParams p, float* ds_partial) {
  constexpr int RPB = 256 / LPR;   // rows per block and per round
  constexpr int R = kPreRowsPerThread;
  __shared__ float part[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = tid % LPR;
  const int i0 = blockIdx.x * (RPB * R) + tid / LPR, h = blockIdx.y, b = blockIdx.z;
  const T* o = static_cast<const T*>(p.o) + b * p.so.b + h * p.so.h + sub * 8;
  const T* dO = static_cast<const T*>(p.dout) + b * p.sdo.b + h * p.sdo.h + sub * 8;
  uint4 a[R], c[R];
  float lse_r[R];           // loaded with the rows, not after the reduction: one memory round trip per block
  const float* lse_h = p.lse + ((int64_t)b * p.Hq + h) * p.N;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r * RPB;
    a[r] = c[r] = make_uint4(0u, 0u, 0u, 0u);
    lse_r[r] = -INFINITY;
    if (i < p.N) {
      a[r] = ld_nc_v4(o + (int64_t)i * p.so.n);
      c[r] = ld_nc_v4(dO + (int64_t)i * p.sdo.n);
      if (p.s_aux && sub == 0) lse_r[r] = __ldg(lse_h + i);
    }
  }
  const float sx = p.s_aux ? p.s_aux[h] : 0.f;
  float contrib = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = i0 + r * RPB;
    const uint32_t au[4] = {a[r].x, a[r].y, a[r].z, a[r].w}, cu[4] = {c[r].x, c[r].y, c[r].z, c[r].w};
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = unpack2<T>(au[t]), y = unpack2<T>(cu[t]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (i < p.N && sub == 0) {
      const int64_t row = ((int64_t)b * p.Hq + h) * p.N + i;
      p.delta[row] = s;
      if (p.s_aux) contrib += (lse_r[r] == -INFINITY) ? 0.f : -expf(sx - lse_r[r]) * s;
    }
  }
  if (ds_partial) {
    contrib = warp_sum(contrib);
    if (lane == 0) part[warp] = contrib;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += part[w];
      ds_partial[((int64_t)b * p.Hq + h) * gridDim.x + blockIdx.x] = t;
    }
  }
}

__global__ void __launch_bounds__(256) ds_aux_reduce_kernel(const float* __restrict__ ds_partial, float* ds_aux, int B,
                                                            int Hq, int nblk) {
  __shared__ float red[256];
  const int h = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b)
    for (int t = threadIdx.x; t < nblk; t += 256) s += ds_partial[((int64_t)b * Hq + h) * nblk + t];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) ds_aux[h] = red[0];
}

// ------------------------------------------------------------------------------------ decode
template <typename T>
__global__ void __launch_bounds__(128) simt_decode_kernel(DecodeParams p) {
  __shared__ float qs[kMaxD];
  __shared__ float sm_m[4], sm_l[4];
  __shared__ float sm_o[4][kMaxD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = p.D;
  const int kvh = h / (p.Hq / p.Hkv);
  const T* q = static_cast<const T*>(p.q) + b * p.sq_b + h * p.sq_h;
  for (int d = threadIdx.x; d < D; d += 128) qs[d] = to_f<T>(q[d]);
  __syncthreads();
  float m = -INFINITY, l = 0.f;
  float acc[kDPL];
#pragma unroll
  for (int t = 0; t < kDPL; ++t) acc[t] = 0.f;
  for (int seg = 0; seg < 2; ++seg) {
    const int len = p.len[seg];
    if (len <= 0) continue;
    const T* kb = static_cast<const T*>(p.k[seg]) + b * p.sk[seg].b + kvh * p.sk[seg].h;
    const T* vb = static_cast<const T*>(p.v[seg]) + b * p.sv[seg].b + kvh * p.sv[seg].h;
    const int64_t sn = p.sk[seg].n, svn = p.sv[seg].n;
    for (int j0 = warp * 32; j0 < len; j0 += 128) {
      const int j = j0 + lane;
      const bool valid = j < len;
      const float s = valid ? dot_row<T>(kb + (int64_t)j * sn, qs, D) * p.scale : -INFINITY;
      const float m_new = fmaxf(m, warp_max(s));
      const float alpha = (m == -INFINITY) ? 0.f : expf(m - m_new);
      const float pj = valid ? expf(s - m_new) : 0.f;
      l = l * alpha + warp_sum(pj);
#pragma unroll
      for (int t = 0; t < kDPL; ++t) acc[t] *= alpha;
      const int cnt = min(32, len - j0);
      for (int jj = 0; jj < cnt; ++jj) {
        const float pjj = __shfl_sync(0xffffffffu, pj, jj);
        const T* vr = vb + (int64_t)(j0 + jj) * svn;
#pragma unroll
        for (int t = 0; t < kDPL; ++t) {
          const int d = lane + 32 * t;
          if (d < D) acc[t] = fmaf(pjj, to_f<T>(vr[d]), acc[t]);
        }
      }
      m = m_new;
    }
  }
  if (lane == 0) {
    sm_m[warp] = m;
    sm_l[warp] = l;
  }
#pragma unroll
  for (int t = 0; t < kDPL; ++t) {
    const int d = lane + 32 * t;
    if (d < D) sm_o[warp][d] = acc[t];
  }
  __syncthreads();
  // merge the four warps plus the s_aux virtual split (m = s_aux, l = 1, o = 0)
  float mg = p.s_aux ? p.s_aux[h] : -INFINITY;
  for (int w = 0; w < 4; ++w) mg = fmaxf(mg, sm_m[w]);
  float lg = p.s_aux ? expf(p.s_aux[h] - mg) : 0.f;
  float a[4];
  for (int w = 0; w < 4; ++w) {
    a[w] = (sm_m[w] == -INFINITY) ? 0.f : expf(sm_m[w] - mg);
    lg += sm_l[w] * a[w];
  }
  lg = fmaxf(lg, 1e-8f);  // decode_kernel.py:222
  T* o = static_cast<T*>(p.o) + b * p.so_b + h * p.so_h;
  for (int d = threadIdx.x; d < D; d += 128) {
    float s = 0.f;
    for (int w = 0; w < 4; ++w) s += sm_o[w][d] * a[w];
    o[d] = from_f<T>(s / lg);
  }
}

template <typename F>
cudaError_t dispatch_dtype(int dtype, F&& f) {
  switch (dtype) {
    case SFA_DTYPE_BF16: return f(__nv_bfloat16{});
    case SFA_DTYPE_FP16: return f(__half{});
    case SFA_DTYPE_FP32: return f(float{});
  }
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t simt_fwd(const AttnParams& p, int dtype, cudaStream_t st) {
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    dim3 grid((p.N + 3) / 4, p.Hq, p.B);
    simt_fwd_kernel<T><<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
  });
}

template <typename T>
cudaError_t launch_preprocess_vec(const AttnParams& p, float* ds_partial, int& nblk, cudaStream_t st) {
  const int lpr = p.D / 8;
  const int rows = 256 / lpr * kPreRowsPerThread;
  nblk = (p.N + rows - 1) / rows;
  dim3 grid(nblk, p.Hq, p.B);
  switch (lpr) {
    case 4: preprocess_vec_kernel<T, 4><<<grid, 256, 0, st>>>(p, ds_partial); break;
    case 8: preprocess_vec_kernel<T, 8><<<grid, 256, 0, st>>>(p, ds_partial); break;
    case 16: preprocess_vec_kernel<T, 16><<<grid, 256, 0, st>>>(p, ds_partial); break;
    case 32: preprocess_vec_kernel<T, 32><<<grid, 256, 0, st>>>(p, ds_partial); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

static bool vec16_ok(const void* ptr, const Strides4& s) {
  return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && s.n % 8 == 0 && s.h % 8 == 0 && s.b % 8 == 0;
}

cudaError_t bwd_preprocess(const AttnParams& p, int dtype, float* ds_partial, cudaStream_t st, int* defer_reduce_nblk) {
  float* dsp = p.s_aux ? ds_partial : nullptr;
  int nblk = (p.N + 7) / 8;
  cudaError_t e;
  const bool vec = (dtype != SFA_DTYPE_FP32) && (p.D == 32 || p.D == 64 || p.D == 128 || p.D == 256) &&
                   vec16_ok(p.o, p.so) && vec16_ok(p.dout, p.sdo);
  if (vec) {
    e = (dtype == SFA_DTYPE_BF16) ? launch_preprocess_vec<__nv_bfloat16>(p, dsp, nblk, st)
                                  : launch_preprocess_vec<__half>(p, dsp, nblk, st);
  } else {
    e = dispatch_dtype(dtype, [&](auto tag) {
      using T = decltype(tag);
      dim3 grid(nblk, p.Hq, p.B);
      preprocess_kernel<T><<<grid, 256, 0, st>>>(p, dsp);
      return cudaGetLastError();
    });
  }
  if (e != cudaSuccess) return e;
  if (defer_reduce_nblk != nullptr) {      // the caller folds the reduce into a later launch
    *defer_reduce_nblk = nblk;
    return e;
  }
  if (p.s_aux && p.ds_aux) {
    ds_aux_reduce_kernel<<<p.Hq, 256, 0, st>>>(ds_partial, p.ds_aux, p.B, p.Hq, nblk);
    e = cudaGetLastError();
  }
  return e;
}

// ds_aux[h] = -sum_{b,i} exp(s_aux[h] - lse[b,h,i]) * delta[b,h,i] (sink_flash_attention.py:653-665) from the delta
// rows the fused backward kernel left in its workspace; fixed summation order -> deterministic.
__global__ void __launch_bounds__(512) ds_aux_from_delta_kernel(const float* __restrict__ delta,
                                                                const float* __restrict__ lse,
                                                                const float* __restrict__ s_aux, float* ds_aux, int B,
                                                                int Hq, int N) {
  __shared__ float red[512];
  const int h = blockIdx.x;
  const float sx = s_aux[h];
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* d = delta + ((int64_t)b * Hq + h) * N;
    const float* l = lse + ((int64_t)b * Hq + h) * N;
    for (int t = threadIdx.x; t < N; t += 512) {
      const float lv = l[t];
      s += (lv == -INFINITY) ? 0.f : -expf(sx - lv) * d[t];
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) ds_aux[h] = red[0];
}

cudaError_t ds_aux_from_delta(const float* delta, const float* lse, const float* s_aux, float* ds_aux, int B, int Hq,
                              int N, cudaStream_t st) {
  ds_aux_from_delta_kernel<<<Hq, 512, 0, st>>>(delta, lse, s_aux, ds_aux, B, Hq, N);
  return cudaGetLastError();
}

cudaError_t ds_aux_reduce(const float* partial, float* ds_aux, int B, int Hq, int nblk, cudaStream_t st) {
  ds_aux_reduce_kernel<<<Hq, 256, 0, st>>>(partial, ds_aux, B, Hq, nblk);
  return cudaGetLastError();
}

cudaError_t simt_bwd(const AttnParams& p, int dtype, int stages, cudaStream_t st) {
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    cudaError_t e = cudaSuccess;
    if (stages & 2) {
      dim3 gq((p.N + 3) / 4, p.Hq, p.B);
      simt_dq_kernel<T><<<gq, 128, 0, st>>>(p);
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
    if (stages & 4) {
      dim3 gk((p.Nkv + 3) / 4, p.Hkv, p.B);
      simt_dkdv_kernel<T><<<gk, 128, 0, st>>>(p);
      e = cudaGetLastError();
    }
    return e;
  });
}

cudaError_t simt_decode(const DecodeParams& p, int dtype, cudaStream_t st) {
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    dim3 grid(p.Hq, p.B);
    simt_decode_kernel<T><<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
  });
}

}  // namespace sfa
