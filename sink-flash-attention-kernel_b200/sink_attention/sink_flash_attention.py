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
