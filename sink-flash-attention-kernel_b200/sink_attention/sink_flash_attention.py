"""sink_flash_attention(q, k, v, num_sink, window_size, s_aux) -- the autograd boundary.

Mirrors the reference operator (sink_attention/sink_flash_attention.py:491-689 of
RulinShao/sink-flash-attention-kernel): same signature, defaults, shape asserts, saved tensors and
return tuple; the Triton launches are replaced by the sm_100a kernels behind libsinkfa.

Attention pattern for query i:  keys j <= i with (j < num_sink  or  j >= i - window_size + 1).
``s_aux`` ([H_q], any float dtype) is the gpt-oss learnable sink logit: it adds exp(s_aux) to
the softmax denominator and nothing to the numerator.
"""
from __future__ import annotations

import torch

from . import _lib


class SinkFlashAttentionFunc(torch.autograd.Function):
    """forward -> O; backward -> (dq, dk, dv, None, None, ds_aux)  (reference :568-667)."""

    @staticmethod
    def forward(ctx, q, k, v, num_sink, window_size, s_aux=None):
        B, H_q, N, D = q.shape
        H_kv = k.shape[1]
        assert k.shape == (B, H_kv, N, D)          # reference :496
        assert v.shape == (B, H_kv, N, D)          # reference :497
        assert H_q % H_kv == 0                     # reference :498
        if k.dtype != q.dtype or v.dtype != q.dtype:
            raise TypeError("q, k, v must share one dtype")
        if q.dtype not in _lib.DTYPE_CODE:
            raise TypeError(f"unsupported dtype {q.dtype} (bf16 / fp16 / fp32)")
        use_s_aux = s_aux is not None
        s_aux_f32 = _lib._s_aux_f32(s_aux, H_q)    # reference :500-503
        o, lse = _lib.fwd(q, k, v, num_sink, window_size, s_aux_f32)
        ctx.save_for_backward(q, k, v, o, lse, s_aux_f32 if use_s_aux else torch.empty(0, device=q.device))
        ctx.num_sink = num_sink
        ctx.window_size = window_size
        ctx.use_s_aux = use_s_aux
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse, s_aux_saved = ctx.saved_tensors
        dq, dk, dv, ds_aux = _lib.bwd(
            q, k, v, o, do, lse, ctx.num_sink, ctx.window_size, s_aux_saved if ctx.use_s_aux else None)
        return dq, dk, dv, None, None, ds_aux


def sink_flash_attention(q, k, v, num_sink=4, window_size=512, s_aux=None):
    """Flash attention with attention sinks (drop-in for the reference's function of the same name).

    Args:
        q: [B, H_q, N, D]; k, v: [B, H_kv, N, D] with H_q % H_kv == 0 (MHA / GQA / MQA).
           bf16 / fp16 run on the tcgen05 kernels (D in {64, 128}); fp32 and other head dims run on
           the CUDA-core kernels.  Transposed views (HF [B, N, H, D] layout) are consumed in place.
        num_sink: number of always-visible leading tokens (default 4).
        window_size: causal sliding window, counting the query itself (default 512).
        s_aux: optional per-Q-head sink logit [H_q].
    Returns:
        O [B, H_q, N, D] in q's dtype.
    """
    return SinkFlashAttentionFunc.apply(q, k, v, num_sink, window_size, s_aux)


def sink_flash_attention_with_lse(q, k, v, num_sink=4, window_size=512, s_aux=None):
    """Forward only; returns (O, LSE) with LSE the natural-log normaliser incl. the s_aux term
    (what the reference saves for backward, :192,556).  Used by the parity tests."""
    s_aux_f32 = _lib._s_aux_f32(s_aux, q.shape[1])
    return _lib.fwd(q, k, v, num_sink, window_size, s_aux_f32)


# =================================================================================================
# Extended geometry: packed (varlen) sequences and chunked prefill, inside the kernels.
# The reference hands these cases to stock FlashAttention (verl_patch.py:73-93, which drops s_aux and the sink
# tokens) or asserts (1 < N_q < N_kv, decode_kernel.py:146; sink_flash_attention.py:496).
# =================================================================================================
def sequence_bounds_from_cu_seqlens(cu_seqlens: torch.Tensor, total: int):
    """cu_seqlens [n_seq + 1] (int, CUDA, as HF / flash-attn pass it) -> (seq_lo, seq_hi) int32 [total]: for every
    packed position the first position of its sequence and one past its last.  Positions beyond cu_seqlens[-1]
    (padding of the packed buffer) form single-token sequences.  No host synchronisation."""
    cu = cu_seqlens.to(torch.int64)
    pos = torch.arange(total, device=cu.device)
    idx = torch.bucketize(pos, cu[1:], right=True)                     # sequence index of every position
    pad = idx >= cu.numel() - 1
    idx = idx.clamp(max=cu.numel() - 2)
    lo = torch.where(pad, pos, cu[idx])
    hi = torch.where(pad, pos + 1, cu[idx + 1])
    return lo.to(torch.int32).contiguous(), hi.to(torch.int32).contiguous()


def sequence_bounds_from_position_ids(position_ids: torch.Tensor):
    """position_ids [B, N] that restart at 0 at every packed sequence (what HF passes for padding-free batches,
    verl_patch.py:182-193) -> (seq_lo, seq_hi) int32 [B, N].  A sequence starts wherever position_ids does not
    continue the previous position by +1.  No host synchronisation."""
    B, N = position_ids.shape
    pos = torch.arange(N, device=position_ids.device).expand(B, N)
    p = position_ids.to(torch.int64)
    start = torch.ones(B, N, dtype=torch.bool, device=p.device)
    start[:, 1:] = p[:, 1:] != p[:, :-1] + 1
    lo = torch.cummax(torch.where(start, pos, torch.zeros_like(pos)), dim=1).values
    nxt = torch.where(start, pos, torch.full_like(pos, N))             # a start at p ends every sequence before it
    nxt = torch.cat([nxt[:, 1:], torch.full((B, 1), N, device=p.device, dtype=nxt.dtype)], dim=1)
    hi = torch.flip(torch.cummin(torch.flip(nxt, dims=[1]), dim=1).values, dims=[1])
    return lo.to(torch.int32).contiguous(), hi.to(torch.int32).contiguous()


class SinkFlashAttentionExFunc(torch.autograd.Function):
    """Same contract as SinkFlashAttentionFunc plus (seq_lo, seq_hi, q_off); N_kv may exceed N_q."""

    @staticmethod
    def forward(ctx, q, k, v, num_sink, window_size, s_aux, seq_lo, seq_hi, q_off):
        B, H_q, N, D = q.shape
        H_kv, N_kv = k.shape[1], k.shape[2]
        assert k.shape == (B, H_kv, N_kv, D) and v.shape == k.shape
        assert H_q % H_kv == 0
        assert 0 <= q_off and q_off + N <= N_kv, f"q_off={q_off}, N_q={N}, N_kv={N_kv}"
        if k.dtype != q.dtype or v.dtype != q.dtype:
            raise TypeError("q, k, v must share one dtype")
        use_s_aux = s_aux is not None
        s_aux_f32 = _lib._s_aux_f32(s_aux, H_q)
        ext = _lib.make_ext(N, N_kv, q_off, seq_lo, seq_hi)
        o, lse = _lib.fwd(q, k, v, num_sink, window_size, s_aux_f32, ext=ext)
        ctx.save_for_backward(q, k, v, o, lse, s_aux_f32 if use_s_aux else torch.empty(0, device=q.device),
                              seq_lo if seq_lo is not None else torch.empty(0, device=q.device),
                              seq_hi if seq_hi is not None else torch.empty(0, device=q.device))
        ctx.cfg = (num_sink, window_size, use_s_aux, seq_lo is not None, q_off)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse, s_aux_saved, seq_lo, seq_hi = ctx.saved_tensors
        num_sink, window_size, use_s_aux, packed, q_off = ctx.cfg
        ext = _lib.make_ext(q.shape[2], k.shape[2], q_off, seq_lo if packed else None, seq_hi if packed else None)
        dq, dk, dv, ds_aux = _lib.bwd(q, k, v, o, do, lse, num_sink, window_size, s_aux_saved if use_s_aux else None,
                                      ext=ext)
        return dq, dk, dv, None, None, ds_aux, None, None, None


def sink_flash_attention_varlen(q, k, v, cu_seqlens=None, num_sink=4, window_size=512, s_aux=None, *,
                                position_ids=None, seq_bounds=None):
    """Sink attention over PACKED sequences: ``q [B, H_q, N, D]``, ``k, v [B, H_kv, N, D]`` hold several sequences
    back to back along N; a row attends only keys of its own sequence, sinks are the first ``num_sink`` tokens of
    each sequence.  Boundaries come from ``cu_seqlens`` ([n_seq + 1], B must be 1), from ``position_ids`` ([B, N],
    restarting at 0) or from precomputed ``seq_bounds = (seq_lo, seq_hi)``.  Differentiable; s_aux as in
    ``sink_flash_attention``."""
    B, _, N, _ = q.shape
    if seq_bounds is None:
        if cu_seqlens is not None:
            assert B == 1, "cu_seqlens describes one packed row (B == 1)"
            seq_bounds = sequence_bounds_from_cu_seqlens(cu_seqlens, N)
        elif position_ids is not None:
            seq_bounds = sequence_bounds_from_position_ids(position_ids)
        else:
            raise ValueError("need cu_seqlens, position_ids or seq_bounds")
    seq_lo, seq_hi = seq_bounds
    return SinkFlashAttentionExFunc.apply(q, k, v, num_sink, window_size, s_aux, seq_lo, seq_hi, 0)


def sink_flash_attention_chunk(q, k, v, num_sink=4, window_size=512, s_aux=None, q_offset=None):
    """Chunked prefill: ``q [B, H_q, N_q, D]`` are the LAST ``N_q`` positions (or positions ``q_offset ..``) of a
    context whose keys ``k, v [B, H_kv, N_kv, D]`` are all present (``1 <= N_q <= N_kv``).  Differentiable."""
    q_off = k.shape[2] - q.shape[2] if q_offset is None else int(q_offset)
    return SinkFlashAttentionExFunc.apply(q, k, v, num_sink, window_size, s_aux, None, None, q_off)
