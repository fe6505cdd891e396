/* libsinkfa -- C ABI of the B200 (sm_100a) sink-flash-attention kernels.
 *
 * This is the drop-in boundary for the hot path of RulinShao/sink-flash-attention-kernel.
 * Each entry point replaces a Triton launch site of the reference (paths relative to the
 * reference's sink_attention/ package):
 *
 *   sfa_fwd      <- SinkFlashAttentionFunc.forward   sink_flash_attention.py:491-566
 *                   (_sink_flash_attn_fwd_kernel, :93-194)
 *   sfa_bwd      <- SinkFlashAttentionFunc.backward  sink_flash_attention.py:568-667
 *                   (delta :582, _bwd_dkdv_kernel :256-364, _bwd_dq_kernel :371-484,
 *                    GQA group sum :648-651, ds_aux :653-665)
 *   sfa_decode   <- sink_decode_attention            decode_kernel.py:120-226
 *                   (_decode_split_kv_kernel :28-113 + the torch phase-2 reduce :201-226)
 *   sfa_decode_ring <- SinkCacheLayer.get_kv() + sink_decode_attention, fused: reads the
 *                   sink buffer and the ring window buffer in place  (cache.py:185-216)
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers; the library never
 *     allocates or frees device memory and never synchronises the stream.
 *   - strides are in ELEMENTS, ordered (batch, head, position, channel); the channel stride
 *     must be 1.  Both the reference's [B,H,N,D] layout and HF's [B,N,H,D] layout (passed
 *     as a transposed view) are accepted without copies.
 *   - softmax scale is 1/sqrt(D) (sink_flash_attention.py:505); s_aux is one fp32 logit
 *     per Q head or NULL; lse is fp32 [B,Hq,N] contiguous, natural log, sink term included.
 *   - mask: valid(i,j) = j<=i && (j<num_sink || j>=i-window+1)   (sink_flash_attention.py:30-39)
 *   - return value: 0 ok; <0 argument error; >0 a cudaError_t.  sfa_last_error() returns a
 *     thread-local description of the last non-zero return.
 *   - `stream` is a cudaStream_t passed as void*.
 */
#ifndef SINKFA_H_
#define SINKFA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFA_DTYPE_BF16 0
#define SFA_DTYPE_FP16 1
#define SFA_DTYPE_FP32 2

#define SFA_OP_FWD 0
#define SFA_OP_BWD 1
#define SFA_OP_DECODE 2

/* implementation selector for sfa_set_impl (testing / cross-checking only) */
#define SFA_IMPL_AUTO 0    /* tcgen05 kernels where the shape allows, CUDA-core kernels otherwise */
#define SFA_IMPL_SIMT 1    /* force the CUDA-core (fp32 math) kernels */

int sfa_version(void);
const char* sfa_last_error(void);
int sfa_set_impl(int impl);
/* timing aid for bench.py: run only the selected backward stages (bit0 = delta/ds_aux preprocess,
 * bit1 = dQ kernel, bit2 = dK/dV kernel); default 7 = all.  Results are complete only with 7 or 15.
 * bit3 (7 + 8 = 15): keep the dQ + dK/dV kernel pair where the one-kernel fused backward would apply
 * (narrow window, no sink tokens, head_dim 64), so that both paths can be tested on the same shape. */
int sfa_set_bwd_stages(int mask);
/* test / diagnostics knobs, all 0 by default (process-wide, atomic):
 *   knob 0: the fused backward sleeps `value` ns in its part-1 math warps before pass 2 and 4x that in epilogue
 *           group 0 before its dQ stores -- widens every cross-warp window of the kernel's pipeline (stress test);
 *   knob 1: value 1 drops the barrier that orders the P-image reads of one tile before the writes of the next
 *           (reproduces the round-1 run-to-run difference in dQ / dK on one GPU; tests only). */
int sfa_set_debug(int knob, int value);
/* performance-debug aid: a device buffer of 3*256*2 int64 into which CTA 0 of the dQ kernel appends
 * (role, event, index, clock64) records; NULL (the default) switches it off. */
int sfa_set_trace_buffer(void* device_buffer);
/* name of the kernel family the last call on this thread dispatched to ("tcgen05", "simt", "mma") */
const char* sfa_last_impl(void);

size_t sfa_workspace_bytes(int op, int B, int Hq, int Hkv, int N, int D, int dtype);

int sfa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux,
            int B, int Hq, int Hkv, int N, int D, int num_sink, int window, int dtype,
            const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
            const int64_t o_strides[4], void* workspace, size_t workspace_bytes, void* stream);

int sfa_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
            const float* s_aux, void* dq, void* dk, void* dv, float* ds_aux,
            int B, int Hq, int Hkv, int N, int D, int num_sink, int window, int dtype,
            const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
            const int64_t o_strides[4], const int64_t do_strides[4], const int64_t dq_strides[4],
            const int64_t dk_strides[4], const int64_t dv_strides[4],
            void* workspace, size_t workspace_bytes, void* stream);

/* Extended geometry: packed (varlen) sequences and chunked prefill / halo keys, inside the kernels -- the cases the
 * reference hands to stock FlashAttention (verl_patch.py:73-93, dropping s_aux and the sinks) or cannot run
 * (1 < N_q < N_kv asserts, decode_kernel.py:146).
 *   q, o, dout, dq, lse: N query rows;  k, v, dk, dv: n_kv key rows (n_kv = 0 means N).
 *   Query row iq sits at absolute key position i = iq + q_off (q_off >= 0, q_off + N <= n_kv) and attends key j iff
 *       lo(iq) <= j <= i   and   (j - lo(iq) < num_sink  or  j >= i - window + 1),
 *   lo(iq) = seq_lo ? seq_lo[b * seq_batch_stride + iq] : 0.  With ext == NULL this is sink_flash_attention.py:30-39.
 *   seq_lo [B or 1][N] int32 (device): first key position of the sequence that query row iq belongs to (from
 *   cu_seq_lens: constant over a sequence, non-decreasing);  seq_hi [B or 1][n_kv] int32 (device): for key j, one
 *   past the last absolute query position of its sequence (the key-stationary backward stops there).  Both or none.
 *   seq_batch_stride: elements between batch rows of the two arrays; 0 = one row shared by the whole batch.
 * Tensor-core paths: head_dim 64, 16-bit -- the persistent forward (packed sequences need num_sink == 0) and the fused
 * backward (narrow window, num_sink == 0, q_off a multiple of the tile's positions); everything else runs on the
 * CUDA-core kernels, which implement the full predicate.  Keys no query attends get dk = dv = 0 only if the caller
 * zero-fills dk / dv when q_off > 0 (the fused kernel writes the key blocks its tiles touch). */
typedef struct sfa_attn_ext {
  const int32_t* seq_lo;
  const int32_t* seq_hi;
  int64_t seq_batch_stride;
  int n_kv;
  int q_off;
} sfa_attn_ext;
int sfa_fwd_ex(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux,
               int B, int Hq, int Hkv, int N, int D, int num_sink, int window, int dtype,
               const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
               const int64_t o_strides[4], void* workspace, size_t workspace_bytes, void* stream,
               const sfa_attn_ext* ext);
int sfa_bwd_ex(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
               const float* s_aux, void* dq, void* dk, void* dv, float* ds_aux,
               int B, int Hq, int Hkv, int N, int D, int num_sink, int window, int dtype,
               const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
               const int64_t o_strides[4], const int64_t do_strides[4], const int64_t dq_strides[4],
               const int64_t dk_strides[4], const int64_t dv_strides[4],
               void* workspace, size_t workspace_bytes, void* stream, const sfa_attn_ext* ext);

/* q,o: [B,Hq,1,D] (strides for batch, head); k,v: [B,Hkv,Nkv,D] (strides for batch, head, position) */
int sfa_decode(const void* q, const void* k, const void* v, void* o, const float* s_aux,
               int B, int Hq, int Hkv, int Nkv, int D, int dtype,
               const int64_t q_strides[2], const int64_t k_strides[3], const int64_t v_strides[3],
               const int64_t o_strides[2], void* workspace, size_t workspace_bytes, void* stream);

/* Decode with per-batch cache lengths and (optionally) a paged KV cache -- SURVEY section 8 row f4; the reference shares
 * ONE length across the batch and keeps the cache contiguous (cache.py:11-13, decode_kernel.py:146-149).
 *   block_table != NULL: k_cache / v_cache are page pools [num_pages, page_size keys] with (page, head, position) element
 *     strides; logical page j of batch row b is physical page block_table[b * block_table_stride + j] (int32, device);
 *     page_size: a power of two >= 32.
 *   block_table == NULL: k_cache / v_cache are [B,Hkv,max_len,D] with (batch, head, position) strides; page_size ignored.
 *   seq_lens: int32 [B] on the device, keys cached for row b (clamped to max_len); NULL: max_len for every row.  A row
 *     with no key yields 0 (or attends only the s_aux sink when s_aux is given).
 * Workspace: sfa_workspace_bytes(SFA_OP_DECODE, B, Hq, Hkv, max_len rounded up to whole pages, D, dtype).
 * bf16 / fp16, head_dim 64 / 128 / 256 (the tensor-core decode kernel); other cases return -12. */
int sfa_decode_paged(const void* q, const void* k_cache, const void* v_cache, void* o, const float* s_aux,
                     const int* block_table, const int* seq_lens, int B, int Hq, int Hkv, int max_len, int D, int dtype,
                     int page_size, int64_t block_table_stride, const int64_t q_strides[2], const int64_t k_strides[3],
                     const int64_t v_strides[3], const int64_t o_strides[2], void* workspace, size_t workspace_bytes,
                     void* stream);

/* Ring-aware decode: attends sink_k/v[:, :, :sink_len] and window_k/v[:, :, :window_len] in place
 * (softmax is order-invariant, so the ring needs no linearisation).  Buffers are
 * [B,Hkv,num_sink,D] and [B,Hkv,window_size,D] with the given (batch, head, position) strides. */
int sfa_decode_ring(const void* q, const void* sink_k, const void* sink_v, const void* win_k, const void* win_v,
                    void* o, const float* s_aux, int B, int Hq, int Hkv, int sink_len, int window_len, int D,
                    int dtype, const int64_t q_strides[2], const int64_t sink_strides[3],
                    const int64_t win_strides[3], const int64_t o_strides[2],
                    void* workspace, size_t workspace_bytes, void* stream);

/* Device-side append of one decoded token to the ring window buffers (replaces the two strided copies of
 * SinkCacheLayer._decode, cache.py:129-147): k_new / v_new [B,Hkv,1,D] with (batch, head) element strides are written
 * to position write_pos (0 <= write_pos < window_size) of win_k / win_v [B,Hkv,window_size,D] with (batch, head,
 * position) strides.  Channel stride 1.  The ring state (write_pos, window_len) stays with the caller, as in the
 * reference.  Pair it with sfa_decode_ring: no per-step linearisation of the cache (cache.py:185-216). */
int sfa_cache_append(const void* k_new, const void* v_new, void* win_k, void* win_v, int B, int Hkv, int D, int dtype,
                     const int64_t new_strides[2], const int64_t win_strides[3], int window_size, int write_pos,
                     void* stream);

/* Ulysses sequence-parallel exchange (BASELINE configs[4]; the reference leaves this all-to-all to verl,
 * verl_patch.py:15-20): one kernel reads the local tensor and stores every row into its final place in the
 * destination rank's receive buffer through CUDA peer mappings (peer_dst: HOST array of P device pointers, entry
 * `rank` being the local buffer).  A cross-rank barrier on the same stream must follow.
 *   mode 0 (sequence -> heads): src [B, L=n, H, D] with element strides src_strides (batch, position, head);
 *           head h -> rank h / (H/P), row (b, rank*n + i, head_off + h % (H/P)) of dst [B, P*n, dst_heads, D].
 *   mode 1 (heads -> sequence): src [B, L=P*n, H=hl, D]; position i -> rank i / n,
 *           row (b, i % n, head_off + rank*hl + h) of dst [B, n, dst_heads, D]. */
int sfa_ulysses_scatter(const void* src, void* const* peer_dst, int P, int rank, int mode, int B, int L, int H, int D,
                        int dtype, const int64_t src_strides[3], int dst_heads, int head_off, void* stream);

/* Fused exchange on the output side of the Ulysses layout: the forward stores every O tile ALSO into the receive
 * buffer of the rank that owns its positions (TMA store to a peer mapping, same staged tile as the local store),
 * the backward stores dQ ONLY there -- no separate scatter pass over O / dQ.  peer[s] is rank s's buffer
 * [B, n_local, heads_total, D] contiguous; this rank's head h lands at head_off + h; position i goes to rank
 * i / n_local.  Supported by the head_dim-64 tcgen05 forward (local O in [B, N, H, D] stride order, n_local a
 * multiple of the tile's positions) and the fused backward; otherwise the call returns -10 and nothing is launched
 * (fall back to sfa_fwd / sfa_bwd + sfa_ulysses_scatter). */
typedef struct sfa_sp_route {
  int P, n_local, heads_total, head_off;
  void* peer[8];
} sfa_sp_route;
int sfa_fwd_sp(const void* q, const void* k, const void* v, void* o, float* lse, const float* s_aux,
               int B, int Hq, int Hkv, int N, int D, int num_sink, int window, int dtype,
               const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
               const int64_t o_strides[4], void* workspace, size_t workspace_bytes, void* stream,
               const sfa_sp_route* o_route);
int sfa_bwd_sp(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
               const float* s_aux, void* dk, void* dv, float* ds_aux,
               int B, int Hq, int Hkv, int N, int D, int num_sink, int window, int dtype,
               const int64_t q_strides[4], const int64_t k_strides[4], const int64_t v_strides[4],
               const int64_t o_strides[4], const int64_t do_strides[4],
               const int64_t dk_strides[4], const int64_t dv_strides[4],
               void* workspace, size_t workspace_bytes, void* stream, const sfa_sp_route* dq_route);

#ifdef __cplusplus
}
#endif
#endif /* SINKFA_H_ */
