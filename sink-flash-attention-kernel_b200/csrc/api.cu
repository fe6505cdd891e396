// extern "C" surface of libsinkfa (declared in include/sinkfa.h): argument validation,
// workspace carving and dispatch to the kernel families.  No device allocation, no stream sync.
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace sfa {

static thread_local char g_err[512] = "";
// Selectors are process-wide ON PURPOSE (autograd runs backward on its own thread: a thread-local selector set by
// the test thread would not reach it) and atomic, so concurrent callers never see a torn or stale-forever value.
static std::atomic<const char*> g_impl{""};
static std::atomic<int> g_force_impl{SFA_IMPL_AUTO};
static std::atomic<int> g_bwd_stages{7};
static std::atomic<int> g_debug[4] = {{0}, {0}, {0}, {0}};
static const bool g_env_fwd_v1 = getenv("SFA_FWD_V1") != nullptr;   // diagnostics: read once, not per call
#ifndef SFA_WIDE64_FWD_MIN_W
#define SFA_WIDE64_FWD_MIN_W 1024
#endif
static bool wide_fwd64(const sfa::AttnParams& p) {
  static const char* env = getenv("SFA_WIDE64_FWD");      // diagnostics: 0 / 1 force the choice
  if (env != nullptr) return env[0] == '1';
  return (p.W < p.N ? p.W : p.N) >= SFA_WIDE64_FWD_MIN_W;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void set_impl_name(const char* name) { g_impl.store(name, std::memory_order_relaxed); }
int debug_knob(int which) { return (which >= 0 && which < 4) ? g_debug[which].load(std::memory_order_relaxed) : 0; }

int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

static Strides4 mk(const int64_t* s) { return Strides4{s[0], s[1], s[2]}; }

static int check_common(int B, int Hq, int Hkv, int N, int D, int dtype, const int64_t* const* strides, int nstr) {
  if (B < 1 || Hq < 1 || Hkv < 1 || N < 1 || D < 1) {
    set_error("invalid sizes B=%d Hq=%d Hkv=%d N=%d D=%d", B, Hq, Hkv, N, D);
    return -1;
  }
  if (Hq % Hkv != 0) {
    set_error("H_q (%d) must be divisible by H_kv (%d)", Hq, Hkv);
    return -2;
  }
  if (D > 256) {
    set_error("head_dim %d > 256 not supported", D);
    return -3;
  }
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16 && dtype != SFA_DTYPE_FP32) {
    set_error("unknown dtype %d", dtype);
    return -4;
  }
  for (int t = 0; t < nstr; ++t)
    if (strides[t][3] != 1) {
      set_error("channel stride must be 1 (tensor %d has %lld)", t, (long long)strides[t][3]);
      return -5;
    }
  return 0;
}

static int cuda_ret(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace sfa

using namespace sfa;

extern "C" {

int sfa_version(void) { return 100; }
const char* sfa_last_error(void) { return g_err; }
const char* sfa_last_impl(void) { return g_impl.load(std::memory_order_relaxed); }
int sfa_set_impl(int impl) {
  if (impl != SFA_IMPL_AUTO && impl != SFA_IMPL_SIMT) {
    set_error("unknown impl %d", impl);
    return -1;
  }
  g_force_impl.store(impl);
  return 0;
}

int sfa_set_debug(int knob, int value) {
  if (knob < 0 || knob >= 4) {
    set_error("unknown debug knob %d", knob);
    return -1;
  }
  g_debug[knob].store(value);
  return 0;
}

int sfa_set_trace_buffer(void* device_buffer) {
  set_trace_buffer(static_cast<long long*>(device_buffer));
  return 0;
}

int sfa_set_bwd_stages(int mask) {
  g_bwd_stages.store(mask & 15);   // bit 3: keep the dQ + dK/dV kernel pair even where the fused kernel applies
  return 0;
}

size_t sfa_workspace_bytes(int op, int B, int Hq, int Hkv, int N, int D, int dtype) {
  (void)dtype;
  if (op == SFA_OP_FWD) return 0;
  if (op == SFA_OP_BWD) {
    // delta [B,Hq,N] + ds_aux partials [B,Hq,ceil(N/8)] + per-row ds_aux contributions [B,Hq,N]
    // + fp32 dK/dV partials of the key blocks shared by neighbouring CTAs of the fused backward kernel
    return 2 * align_up((size_t)B * Hq * N * 4, 256) + align_up((size_t)B * Hq * ((N + 7) / 8) * 4, 256) +
           align_up(tc_bwd_fused_workspace_bytes(), 256);
  }
  if (op == SFA_OP_DECODE) {
    const int splits = mma_decode_splits(B, Hq, Hkv, N);
    return align_up((size_t)B * Hq * splits * 2 * 4, 256) + align_up((size_t)B * Hq * splits * D * 4, 256);
  }
  return 0;
}

// extended geometry (sfa_attn_ext): validated and copied into the problem descriptor
static int apply_ext(AttnParams& p, const sfa_attn_ext* ext) {
  p.Nkv = p.N;
  p.q_off = 0;
  p.seq_lo = p.seq_hi = nullptr;
  p.seq_bs = 0;
  if (ext == nullptr) return 0;
  const int nkv = ext->n_kv > 0 ? ext->n_kv : p.N;
  if (ext->q_off < 0 || (int64_t)ext->q_off + p.N > nkv) {
    set_error("invalid chunk geometry: q_off=%d, %d query rows, %d key rows (need q_off >= 0 and q_off + N <= n_kv)",
              ext->q_off, p.N, nkv);
    return -11;
  }
  if ((ext->seq_lo == nullptr) != (ext->seq_hi == nullptr)) {
    set_error("seq_lo and seq_hi must be given together");
    return -11;
  }
  if (ext->seq_lo != nullptr && ext->seq_batch_stride != 0 && ext->seq_batch_stride < nkv) {
    set_error("seq_batch_stride (%lld) must be 0 (shared) or >= n_kv", (long long)ext->seq_batch_stride);
    return -11;
  }
  p.Nkv = nkv;
  p.q_off = ext->q_off;
  p.seq_lo = ext->seq_lo;
  p.seq_hi = ext->seq_hi;
  p.seq_bs = ext->seq_batch_stride;
  return 0;
}

static bool route_from_c(const sfa_sp_route* in, SpRoute& out) {
  if (in->P < 1 || in->P > 8 || in->n_local < 1 || in->heads_total < 1 || in->head_off < 0) return false;
  out.P = in->P; out.n_local = in->n_local; out.heads_total = in->heads_total; out.head_off = in->head_off;
  for (int r = 0; r < 8; ++r) out.peer[r] = r < in->P ? in->peer[r] : nullptr;
  for (int r = 0; r < in->P; ++r)
    if (out.peer[r] == nullptr) return false;
  return true;
}

static int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux, int B, int Hq,
                    int Hkv, int N, int D, int num_sink, int window, int dtype, const int64_t q_strides[4],
                    const int64_t k_strides[4], const int64_t v_strides[4], const int64_t o_strides[4], void* workspace,
                    size_t workspace_bytes, void* stream, const sfa_sp_route* o_route, const sfa_attn_ext* ext = nullptr) {
  (void)workspace;
  (void)workspace_bytes;
  const int g_force_impl = sfa::g_force_impl.load();
  const int64_t* ss[4] = {q_strides, k_strides, v_strides, o_strides};
  if (int r = check_common(B, Hq, Hkv, N, D, dtype, ss, 4)) return r;
  if (!q || !k || !v || !o || !lse) {
    set_error("null tensor pointer");
    return -6;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.k = k; p.v = v; p.o = o; p.lse = lse; p.s_aux = s_aux;
  p.sq = mk(q_strides); p.sk = mk(k_strides); p.sv = mk(v_strides); p.so = mk(o_strides);
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.N = N; p.D = D;
  p.S = num_sink < 0 ? 0 : num_sink;
  p.W = window < 0 ? 0 : window;
  p.scale = 1.0f / sqrtf((float)D);
  if (int r = apply_ext(p, ext)) return r;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SpRoute rt;
  if (o_route != nullptr) {
    if (!route_from_c(o_route, rt)) {
      set_error("invalid sequence-parallel route");
      return -9;
    }
    p.o_route = &rt;
    if (g_force_impl == SFA_IMPL_SIMT || !tc_fwd_supported(p, dtype) || !tc_fwd64_route_supported(p, dtype)) {
      set_error("routed O store needs the head_dim-64 tcgen05 forward, an HF-order local O and n_local a multiple of the tile");
      return -10;
    }
  }
  if (g_force_impl != SFA_IMPL_SIMT && tc_fwd_supported(p, dtype)) {
    set_impl_name("tcgen05");
    // head_dim 64 with a wide window (long KV loop per tile): the two-tile kernel built for that
    if (p.D == 64 && wide_fwd64(p) && tc_fwd128_supported(p, dtype) && !g_env_fwd_v1)
      return cuda_ret(tc_fwd128(p, dtype, st), "sfa_fwd(tcgen05/fwd128<64>)");
    if (tc_fwd64_supported(p, dtype) && !(g_env_fwd_v1 && p.o_route == nullptr && !p.has_ext()))
      return cuda_ret(tc_fwd64(p, dtype, st), "sfa_fwd(tcgen05/fwd64)");
    // 64 < head_dim <= 128: persistent kernel, two query tiles per CTA
    if (p.D > 64 && tc_fwd128_supported(p, dtype) && !g_env_fwd_v1)
      return cuda_ret(tc_fwd128(p, dtype, st), "sfa_fwd(tcgen05/fwd128)");
    // the one-tile-per-CTA kernel (head_dim > 64) takes packed sequences without sink tokens, not chunk offsets
    if (p.q_off == 0 && p.Nkv == p.N && !(p.seq_lo != nullptr && p.S > 0))
      return cuda_ret(tc_fwd(p, dtype, st), "sfa_fwd(tcgen05)");
  }
  set_impl_name("simt");
  return cuda_ret(simt_fwd(p, dtype, st), "sfa_fwd(simt)");
}

int sfa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux, int B, int Hq, int Hkv,
            int N, int D, int num_sink, int window, int dtype, const int64_t q_strides[4], const int64_t k_strides[4],
            const int64_t v_strides[4], const int64_t o_strides[4], void* workspace, size_t workspace_bytes,
            void* stream) {
  return fwd_impl(q, k, v, o, lse, s_aux, B, Hq, Hkv, N, D, num_sink, window, dtype, q_strides, k_strides, v_strides,
                  o_strides, workspace, workspace_bytes, stream, nullptr);
}

int sfa_fwd_sp(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux, int B, int Hq,
               int Hkv, int N, int D, int num_sink, int window, int dtype, const int64_t q_strides[4],
               const int64_t k_strides[4], const int64_t v_strides[4], const int64_t o_strides[4], void* workspace,
               size_t workspace_bytes, void* stream, const sfa_sp_route* o_route) {
  if (o_route == nullptr) {
    set_error("sfa_fwd_sp needs a route");
    return -9;
  }
  return fwd_impl(q, k, v, o, lse, s_aux, B, Hq, Hkv, N, D, num_sink, window, dtype, q_strides, k_strides, v_strides,
                  o_strides, workspace, workspace_bytes, stream, o_route);
}

static int bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
            const float* s_aux, void* dq, void* dk, void* dv, float* ds_aux, int B, int Hq, int Hkv, int N, int D,
            int num_sink, int window, int dtype, const int64_t q_strides[4], const int64_t k_strides[4],
            const int64_t v_strides[4], const int64_t o_strides[4], const int64_t do_strides[4],
            const int64_t dq_strides[4], const int64_t dk_strides[4], const int64_t dv_strides[4], void* workspace,
            size_t workspace_bytes, void* stream, const sfa_sp_route* dq_route, const sfa_attn_ext* ext = nullptr) {
  const int g_force_impl = sfa::g_force_impl.load(), g_bwd_stages = sfa::g_bwd_stages.load();
  const int64_t* ss[8] = {q_strides, k_strides, v_strides, o_strides, do_strides, dq_strides, dk_strides, dv_strides};
  if (int r = check_common(B, Hq, Hkv, N, D, dtype, ss, 8)) return r;
  if (!q || !k || !v || !o || !dout || !lse || (!dq && !dq_route) || !dk || !dv) {
    set_error("null tensor pointer");
    return -6;
  }
  const size_t need = sfa_workspace_bytes(SFA_OP_BWD, B, Hq, Hkv, N, D, dtype);
  if (!workspace || workspace_bytes < need) {
    set_error("backward workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return -7;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.k = k; p.v = v; p.o = const_cast<void*>(o); p.lse = const_cast<float*>(lse); p.s_aux = s_aux;
  p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv; p.ds_aux = ds_aux;
  p.sq = mk(q_strides); p.sk = mk(k_strides); p.sv = mk(v_strides); p.so = mk(o_strides);
  p.sdo = mk(do_strides); p.sdq = mk(dq_strides); p.sdk = mk(dk_strides); p.sdv = mk(dv_strides);
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.N = N; p.D = D;
  p.S = num_sink < 0 ? 0 : num_sink;
  p.W = window < 0 ? 0 : window;
  p.scale = 1.0f / sqrtf((float)D);
  if (int r = apply_ext(p, ext)) return r;
  p.delta = static_cast<float*>(workspace);
  const size_t rows_bytes = align_up((size_t)B * Hq * N * 4, 256);
  float* ds_partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + rows_bytes);
  p.dsrow = reinterpret_cast<float*>(static_cast<char*>(workspace) + rows_bytes + align_up((size_t)B * Hq * ((N + 7) / 8) * 4, 256));
  float* fused_part = reinterpret_cast<float*>(reinterpret_cast<char*>(p.dsrow) + rows_bytes);
  p.kv_part = fused_part;      // the kernel pair and the fused kernel never run in the same call: one region serves both
  p.kv_part_bytes = tc_bwd_fused_workspace_bytes();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool use_tc = g_force_impl != SFA_IMPL_SIMT && tc_bwd_supported(p, dtype);      // the dQ + dK/dV kernel pair
  // with the extended geometry (packed sequences, chunked prefill) the fused kernel is the only tensor-core path
  const bool use_fused = g_force_impl != SFA_IMPL_SIMT && (g_bwd_stages & 14) == 6 && (use_tc || p.has_ext()) &&
                         tc_bwd_fused_supported(p, dtype);
  SpRoute rt;
  if (dq_route != nullptr) {
    if (!route_from_c(dq_route, rt)) {
      set_error("invalid sequence-parallel route");
      return -9;
    }
    p.dq_route = &rt;
    if (!use_fused) {
      set_error("routed dQ store needs the fused head_dim-64 backward (narrow window, no sink tokens)");
      return -10;
    }
  }
  // narrow window, no sink tokens, head_dim 64: ONE kernel computes delta, the ds_aux rows, dQ, dK and dV
  if (use_fused) {
    set_impl_name("tcgen05-fused");
    if (!tc_bwd_fused_computes_delta(p)) {
      // delta + ds_aux block partials (streaming pass over O and dO), then the fused kernel; the ds_aux partials
      // are reduced by extra blocks of the fused kernel's fix-up launch
      int ds_nblk = 0;
      if (g_bwd_stages & 1)
        if (int r = cuda_ret(bwd_preprocess(p, dtype, ds_partial, st, &ds_nblk), "sfa_bwd(preprocess)")) return r;
      const bool red = (g_bwd_stages & 1) && p.s_aux && p.ds_aux;
      return cuda_ret(tc_bwd_fused(p, dtype, fused_part, red ? ds_partial : nullptr, red ? ds_nblk : 0, st),
                      "sfa_bwd(tcgen05 fused)");
    }
    // delta by the kernel's own delta warps; ds_aux from the delta rows by extra blocks of the fix-up launch
    const bool red = p.s_aux && p.ds_aux;
    return cuda_ret(tc_bwd_fused(p, dtype, fused_part, red ? p.delta : nullptr, red ? N : 0, st), "sfa_bwd(tcgen05 fused)");
  }
  // narrow windows (one KV item per tile): the dQ kernel can compute delta = rowsum(P o dP) and the ds_aux rows
  // itself -- no preprocess pass over O and dO
  const bool fused = use_tc && tc_bwd_fuses_delta(p, dtype);
  if ((g_bwd_stages & 1) && !fused)
    if (int r = cuda_ret(bwd_preprocess(p, dtype, ds_partial, st), "sfa_bwd(preprocess)")) return r;
  if (use_tc) {
    set_impl_name("tcgen05");
    if (int r = cuda_ret(tc_bwd(p, dtype, g_bwd_stages & 7, st), "sfa_bwd(tcgen05)")) return r;
    if (fused && (g_bwd_stages & 2) && p.s_aux && p.ds_aux)
      return cuda_ret(ds_aux_reduce(p.dsrow, p.ds_aux, B, Hq, N, st), "sfa_bwd(ds_aux reduce)");
    return 0;
  }
  set_impl_name("simt");
  return cuda_ret(simt_bwd(p, dtype, g_bwd_stages & 7, st), "sfa_bwd(simt)");
}

int sfa_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
            const float* s_aux, void* dq, void* dk, void* dv, float* ds_aux, int B, int Hq, int Hkv, int N, int D,
            int num_sink, int window, int dtype, const int64_t q_strides[4], const int64_t k_strides[4],
            const int64_t v_strides[4], const int64_t o_strides[4], const int64_t do_strides[4],
            const int64_t dq_strides[4], const int64_t dk_strides[4], const int64_t dv_strides[4], void* workspace,
            size_t workspace_bytes, void* stream) {
  return bwd_impl(q, k, v, o, dout, lse, s_aux, dq, dk, dv, ds_aux, B, Hq, Hkv, N, D, num_sink, window, dtype, q_strides,
                  k_strides, v_strides, o_strides, do_strides, dq_strides, dk_strides, dv_strides, workspace,
                  workspace_bytes, stream, nullptr);
}

int sfa_bwd_sp(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
               const float* s_aux, void* dk, void* dv, float* ds_aux, int B, int Hq, int Hkv, int N, int D, int num_sink,
               int window, int dtype, const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
               const int64_t o_strides[4], const int64_t do_strides[4], const int64_t dk_strides[4],
               const int64_t dv_strides[4], void* workspace, size_t workspace_bytes, void* stream,
               const sfa_sp_route* dq_route) {
  if (dq_route == nullptr) {
    set_error("sfa_bwd_sp needs a route");
    return -9;
  }
  const int64_t dq_strides[4] = {(int64_t)dq_route->n_local * dq_route->heads_total * D, D,
                                 (int64_t)dq_route->heads_total * D, 1};
  return bwd_impl(q, k, v, o, dout, lse, s_aux, nullptr, dk, dv, ds_aux, B, Hq, Hkv, N, D, num_sink, window, dtype,
                  q_strides, k_strides, v_strides, o_strides, do_strides, dq_strides, dk_strides, dv_strides, workspace,
                  workspace_bytes, stream, dq_route);
}

int sfa_fwd_ex(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux, int B, int Hq, int Hkv,
               int N, int D, int num_sink, int window, int dtype, const int64_t q_strides[4], const int64_t k_strides[4],
               const int64_t v_strides[4], const int64_t o_strides[4], void* workspace, size_t workspace_bytes,
               void* stream, const sfa_attn_ext* ext) {
  return fwd_impl(q, k, v, o, lse, s_aux, B, Hq, Hkv, N, D, num_sink, window, dtype, q_strides, k_strides, v_strides,
                  o_strides, workspace, workspace_bytes, stream, nullptr, ext);
}

int sfa_bwd_ex(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
               const float* s_aux, void* dq, void* dk, void* dv, float* ds_aux, int B, int Hq, int Hkv, int N, int D,
               int num_sink, int window, int dtype, const int64_t q_strides[4], const int64_t k_strides[4],
               const int64_t v_strides[4], const int64_t o_strides[4], const int64_t do_strides[4],
               const int64_t dq_strides[4], const int64_t dk_strides[4], const int64_t dv_strides[4], void* workspace,
               size_t workspace_bytes, void* stream, const sfa_attn_ext* ext) {
  return bwd_impl(q, k, v, o, dout, lse, s_aux, dq, dk, dv, ds_aux, B, Hq, Hkv, N, D, num_sink, window, dtype, q_strides,
                  k_strides, v_strides, o_strides, do_strides, dq_strides, dk_strides, dv_strides, workspace,
                  workspace_bytes, stream, nullptr, ext);
}

static int decode_impl(DecodeParams& p, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int g_force_impl = sfa::g_force_impl.load();
  const int L = p.len[0] + p.len[1];
  if (L < 1) {
    set_error("decode needs at least one cached key");
    return -8;
  }
  for (int s = 0; s < 2; ++s)
    if (p.len[s] <= 0) {  // keep pointers dereferenceable for zero-sized segments
      p.len[s] = 0;
      p.k[s] = p.k[1 - s];
      p.v[s] = p.v[1 - s];
      p.sk[s] = p.sk[1 - s];
      p.sv[s] = p.sv[1 - s];
    }
  if (g_force_impl != SFA_IMPL_SIMT && mma_decode_supported(p, dtype)) {
    p.splits = mma_decode_splits(p.B, p.Hq, p.Hkv, L, (p.paged && p.block_table != nullptr) ? p.page_size : 1);
    {
      const size_t ml = align_up((size_t)p.B * p.Hq * p.splits * 2 * 4, 256);
      const size_t po = align_up((size_t)p.B * p.Hq * p.splits * p.D * 4, 256);
      if (!workspace || workspace_bytes < ml + po) {
        set_error("decode workspace too small: need %zu bytes, got %zu", ml + po, workspace_bytes);
        return -7;
      }
      p.part_ml = static_cast<float*>(workspace);
      p.part_o = reinterpret_cast<float*>(static_cast<char*>(workspace) + ml);
    }
    set_impl_name("mma");
    return cuda_ret(mma_decode(p, dtype, st), "sfa_decode(mma)");
  }
  if (p.paged) {
    set_error("per-batch cache lengths / paged KV need the tensor-core decode kernel (bf16 / fp16, head_dim 64 / 128 / 256, "
              "16-byte aligned rows, page size a power of two >= 32)");
    return -12;
  }
  p.splits = 1;
  set_impl_name("simt");
  return cuda_ret(simt_decode(p, dtype, st), "sfa_decode(simt)");
}

int sfa_decode(const void* q, const void* k, const void* v, void* o, const float* s_aux, int B, int Hq, int Hkv, int Nkv,
               int D, int dtype, const int64_t q_strides[2], const int64_t k_strides[3], const int64_t v_strides[3],
               const int64_t o_strides[2], void* workspace, size_t workspace_bytes, void* stream) {
  const int64_t* none[1] = {nullptr};
  if (int r = check_common(B, Hq, Hkv, Nkv, D, dtype, none, 0)) return r;
  if (!q || !k || !v || !o) {
    set_error("null tensor pointer");
    return -6;
  }
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.o = o; p.s_aux = s_aux;
  p.k[0] = k; p.v[0] = v; p.len[0] = Nkv; p.len[1] = 0;
  p.sk[0] = mk(k_strides); p.sv[0] = mk(v_strides);
  p.sq_b = q_strides[0]; p.sq_h = q_strides[1]; p.so_b = o_strides[0]; p.so_h = o_strides[1];
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.D = D;
  p.scale = 1.0f / sqrtf((float)D);
  return decode_impl(p, dtype, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int sfa_decode_paged(const void* q, const void* k_cache, const void* v_cache, void* o, const float* s_aux,
                     const int* block_table, const int* seq_lens, int B, int Hq, int Hkv, int max_len, int D, int dtype,
                     int page_size, int64_t block_table_stride, const int64_t q_strides[2], const int64_t k_strides[3],
                     const int64_t v_strides[3], const int64_t o_strides[2], void* workspace, size_t workspace_bytes,
                     void* stream) {
  const int64_t* none[1] = {nullptr};
  if (int r = check_common(B, Hq, Hkv, max_len, D, dtype, none, 0)) return r;
  if (!q || !k_cache || !v_cache || !o) {
    set_error("null tensor pointer");
    return -6;
  }
  if (block_table != nullptr && (page_size < 32 || (page_size & (page_size - 1)) != 0)) {
    set_error("page_size must be a power of two >= 32, got %d", page_size);
    return -1;
  }
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.o = o; p.s_aux = s_aux;
  p.k[0] = k_cache; p.v[0] = v_cache;
  // the planning length: whole pages (ranges start on page boundaries); the rows' own lengths clip it in the kernel
  p.len[0] = block_table != nullptr ? (max_len + page_size - 1) / page_size * page_size : max_len;
  p.len[1] = 0;
  p.sk[0] = mk(k_strides); p.sv[0] = mk(v_strides);
  p.sq_b = q_strides[0]; p.sq_h = q_strides[1]; p.so_b = o_strides[0]; p.so_h = o_strides[1];
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.D = D;
  p.scale = 1.0f / sqrtf((float)D);
  p.paged = 1;
  p.block_table = block_table; p.bt_stride = block_table_stride; p.seq_lens = seq_lens;
  p.page_size = block_table != nullptr ? page_size : 0;
  p.lg_page = 0;
  while (block_table != nullptr && (1 << p.lg_page) < page_size) ++p.lg_page;
  return decode_impl(p, dtype, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int sfa_decode_ring(const void* q, const void* sink_k, const void* sink_v, const void* win_k, const void* win_v, void* o,
                    const float* s_aux, int B, int Hq, int Hkv, int sink_len, int window_len, int D, int dtype,
                    const int64_t q_strides[2], const int64_t sink_strides[3], const int64_t win_strides[3],
                    const int64_t o_strides[2], void* workspace, size_t workspace_bytes, void* stream) {
  const int64_t* none[1] = {nullptr};
  if (int r = check_common(B, Hq, Hkv, sink_len + window_len, D, dtype, none, 0)) return r;
  if (!q || !o || (sink_len > 0 && (!sink_k || !sink_v)) || (window_len > 0 && (!win_k || !win_v))) {
    set_error("null tensor pointer");
    return -6;
  }
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.o = o; p.s_aux = s_aux;
  p.k[0] = sink_k; p.v[0] = sink_v; p.len[0] = sink_len;
  p.k[1] = win_k; p.v[1] = win_v; p.len[1] = window_len;
  p.sk[0] = p.sv[0] = mk(sink_strides);
  p.sk[1] = p.sv[1] = mk(win_strides);
  p.sq_b = q_strides[0]; p.sq_h = q_strides[1]; p.so_b = o_strides[0]; p.so_h = o_strides[1];
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.D = D;
  p.scale = 1.0f / sqrtf((float)D);
  return decode_impl(p, dtype, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int sfa_cache_append(const void* k_new, const void* v_new, void* win_k, void* win_v, int B, int Hkv, int D, int dtype,
                     const int64_t new_strides[2], const int64_t win_strides[3], int window_size, int write_pos,
                     void* stream) {
  if (!k_new || !v_new || !win_k || !win_v) {
    set_error("null tensor pointer");
    return -6;
  }
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16 && dtype != SFA_DTYPE_FP32) {
    set_error("unknown dtype %d", dtype);
    return -4;
  }
  if (B < 1 || Hkv < 1 || D < 1 || window_size < 1 || write_pos < 0 || write_pos >= window_size) {
    set_error("invalid append geometry B=%d Hkv=%d D=%d window_size=%d write_pos=%d", B, Hkv, D, window_size, write_pos);
    return -1;
  }
  return cuda_ret(cache_append(k_new, v_new, win_k, win_v, B, Hkv, D, dtype == SFA_DTYPE_FP32 ? 4 : 2, new_strides,
                               win_strides, write_pos, static_cast<cudaStream_t>(stream)), "sfa_cache_append");
}

int sfa_ulysses_scatter(const void* src, void* const* peer_dst, int P, int rank, int mode, int B, int L, int H, int D,
                        int dtype, const int64_t src_strides[3], int dst_heads, int head_off, void* stream) {
  if (!src || !peer_dst) {
    set_error("null tensor pointer");
    return -6;
  }
  if (dtype != SFA_DTYPE_BF16 && dtype != SFA_DTYPE_FP16 && dtype != SFA_DTYPE_FP32) {
    set_error("unknown dtype %d", dtype);
    return -4;
  }
  const int es = dtype == SFA_DTYPE_FP32 ? 4 : 2;
  if (P < 1 || P > 16 || rank < 0 || rank >= P || (mode != 0 && mode != 1) || B < 1 || L < 1 || H < 1 || D < 1) {
    set_error("invalid exchange geometry P=%d rank=%d mode=%d B=%d L=%d H=%d D=%d", P, rank, mode, B, L, H, D);
    return -1;
  }
  if ((mode == 0 ? H : L) % P != 0) {
    set_error("%s (%d) must be divisible by the sequence-parallel size (%d)", mode == 0 ? "heads" : "sequence length",
              mode == 0 ? H : L, P);
    return -2;
  }
  if ((D * es) % 16 != 0 || reinterpret_cast<uintptr_t>(src) % 16 != 0 || (src_strides[0] * es) % 16 != 0 ||
      (src_strides[1] * es) % 16 != 0 || (src_strides[2] * es) % 16 != 0) {
    set_error("rows must be 16-byte aligned multiples of 16 bytes");
    return -5;
  }
  for (int r = 0; r < P; ++r)
    if (peer_dst[r] == nullptr || reinterpret_cast<uintptr_t>(peer_dst[r]) % 16 != 0) {
      set_error("peer buffer %d is null or not 16-byte aligned", r);
      return -6;
    }
  return cuda_ret(ulysses_scatter(src, peer_dst, P, rank, mode, B, L, H, D, es, src_strides, dst_heads, head_off,
                                  static_cast<cudaStream_t>(stream)), "sfa_ulysses_scatter");
}

}  // extern "C"
