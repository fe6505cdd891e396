"""``patch_for_generation``: HF ``model.generate()`` with sink attention and the sink+ring KV cache.

Behaviour follows the reference's ``sink_attention/generate_patch.py`` (:52-187): a process-global
swap of transformers' ``_flash_attention_forward``; prefill (N_q > 1) runs
``sink_flash_attention(q, k, v, num_sink, window_size)`` and a single-token step runs
``sink_decode_attention(q, k, v)`` over the K/V the cache returned; only varlen / packed /
non-causal calls fall back; ``s_aux``, ``sliding_window``, ``attention_mask`` and ``softcap`` are not
consulted on this path (as in the reference).  Returns a ``SinkAttentionCache`` to pass as
``past_key_values``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import cache as _cache
from .cache import SinkAttentionCache
from .decode_kernel import sink_decode_attention
from .sink_flash_attention import sink_flash_attention
from .verl_patch import _is_packed

_original_flash_attention_forward = None

_GENERATION_CONFIG = {"num_sink": 4, "window_size": 4096, "enabled": False}


def _generation_flash_attention_forward(
    query_states: torch.Tensor,
    key_states: torch.Tensor,
    value_states: torch.Tensor,
    attention_mask: Optional[torch.Tensor],
    query_length: int,
    is_causal: bool = True,
    dropout: float = 0.0,
    position_ids: Optional[torch.Tensor] = None,
    softmax_scale: Optional[float] = None,
    sliding_window: Optional[int] = None,
    use_top_left_mask: bool = False,
    softcap: Optional[float] = None,
    deterministic: Optional[bool] = None,
    cu_seq_lens_q: Optional[torch.LongTensor] = None,
    cu_seq_lens_k: Optional[torch.LongTensor] = None,
    max_length_q: Optional[int] = None,
    max_length_k: Optional[int] = None,
    target_dtype: Optional[torch.dtype] = None,
    implementation: Optional[str] = None,
    **kwargs,
):
    varlen = all(x is not None for x in (cu_seq_lens_q, cu_seq_lens_k, max_length_q, max_length_k))
    packed = position_ids is not None and query_states.size(0) > 0 and _is_packed(position_ids)
    if varlen or packed or not is_causal:
        return _original_flash_attention_forward(
            query_states, key_states, value_states, attention_mask, query_length,
            is_causal=is_causal, dropout=dropout, position_ids=position_ids, softmax_scale=softmax_scale,
            sliding_window=sliding_window, use_top_left_mask=use_top_left_mask, softcap=softcap,
            deterministic=deterministic, cu_seq_lens_q=cu_seq_lens_q, cu_seq_lens_k=cu_seq_lens_k,
            max_length_q=max_length_q, max_length_k=max_length_k, target_dtype=target_dtype,
            implementation=implementation, **kwargs)

    q = query_states.transpose(1, 2)         # [B,H,N,D] views; the kernels take strides
    k = key_states.transpose(1, 2)
    v = value_states.transpose(1, 2)
    if q.shape[2] > 1:
        out = sink_flash_attention(q, k, v, num_sink=_GENERATION_CONFIG["num_sink"],
                                   window_size=_GENERATION_CONFIG["window_size"])
    else:
        layer = _cache.ring_layer_of(key_states)
        if layer is not None:
            # the cache handed out views of its ring: attend sink + ring buffers in place (sfa_decode_ring) --
            # softmax is order-invariant and RoPE is already applied, so no linearisation copy is needed
            out = layer.decode_attention(q)
        else:
            out = sink_decode_attention(q, k, v)  # every cached key is attended
    out = out.transpose(1, 2)
    return out if out.is_contiguous() else out.contiguous()


def patch_for_generation(model=None, num_sink: int = 4, window_size: int = 4096) -> SinkAttentionCache:
    """Patch transformers for cached generation and return a fresh ``SinkAttentionCache``.

    ``model`` is accepted for API symmetry and not used: the patch is process-global.
    """
    global _original_flash_attention_forward
    _GENERATION_CONFIG.update(num_sink=num_sink, window_size=window_size, enabled=True)
    _cache._FAST_DECODE["enabled"] = True
    import transformers.modeling_flash_attention_utils as fa_utils
    if fa_utils._flash_attention_forward is not _generation_flash_attention_forward:
        _original_flash_attention_forward = fa_utils._flash_attention_forward   # never save our own hook
    fa_utils._flash_attention_forward = _generation_flash_attention_forward
    try:
        from transformers.integrations import flash_attention
        flash_attention._flash_attention_forward = _generation_flash_attention_forward
    except (ImportError, AttributeError):
        pass
    return SinkAttentionCache(num_sink=num_sink, window_size=window_size)


def unpatch_generation():
    """Restore the original ``_flash_attention_forward``."""
    global _original_flash_attention_forward
    if _original_flash_attention_forward is None:
        return
    import transformers.modeling_flash_attention_utils as fa_utils
    fa_utils._flash_attention_forward = _original_flash_attention_forward
    try:
        from transformers.integrations import flash_attention
        flash_attention._flash_attention_forward = _original_flash_attention_forward
    except (ImportError, AttributeError):
        pass
    _GENERATION_CONFIG["enabled"] = False
    _cache._FAST_DECODE["enabled"] = False
    _original_flash_attention_forward = None
