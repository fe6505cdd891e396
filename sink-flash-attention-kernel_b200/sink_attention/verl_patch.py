"""verl / HuggingFace hook: route ``_flash_attention_forward`` to the sink-attention kernels.

Behaviour follows the reference's ``sink_attention/verl_patch.py`` (:33-263): the replacement takes
HF's ``[B, N, H, D]`` tensors, pops gpt-oss's ``s_aux`` from ``**kwargs``, falls back to the saved
original for varlen / packed / non-causal / padded / soft-capped calls, routes ``N_q != N_kv`` to
the decode kernel, uses ``num_sink=0`` with ``window = sliding_window or N`` for prefill, slices
``s_aux`` to the local heads under Ulysses sequence parallelism, and ignores ``softmax_scale``
(the kernels use 1/sqrt(D), as the reference does).

Difference (layout-fused boundary): the kernels take strides, so the ``transpose(1, 2)`` views are
consumed in place and the output is produced directly in ``[B, N, H, D]`` memory -- the
reference's four ``.contiguous()`` copies per call (:164-166,177) disappear.
"""
from __future__ import annotations

from typing import Optional

import torch

from .decode_kernel import sink_decode_attention
from .sink_flash_attention import sink_flash_attention

_original_flash_attention_forward = None


def _is_packed(position_ids: Optional[torch.Tensor]) -> bool:
    """Packed sequences restart their position ids mid-row (reference :182-193)."""
    if position_ids is None or position_ids.dim() < 2 or position_ids.size(1) <= 1:
        return False
    return bool((position_ids[:, 1:] < position_ids[:, :-1]).any().item())


def _ulysses_rank(sp_size: int) -> int:
    try:
        from verl.utils.ulysses import get_ulysses_sequence_parallel_rank
        return int(get_ulysses_sequence_parallel_rank())
    except (ImportError, RuntimeError):
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_rank() % sp_size
        return 0


def _local_s_aux(s_aux: Optional[torch.Tensor], h_q: int, allow_mismatch_none: bool) -> Optional[torch.Tensor]:
    """s_aux covers all heads; after the Ulysses all-to-all a rank holds H_total/sp of them
    (reference :132-154 for prefill, :101-116 for decode)."""
    if s_aux is None:
        return None
    h_total = s_aux.shape[0]
    if h_total == h_q:
        return s_aux
    if h_total > h_q and h_total % h_q == 0:
        r = _ulysses_rank(h_total // h_q)
        return s_aux[r * h_q:(r + 1) * h_q]
    return None


def _sink_flash_attention_forward(
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
    s_aux = kwargs.pop("s_aux", None)
    varlen = all(x is not None for x in (cu_seq_lens_q, cu_seq_lens_k, max_length_q, max_length_k))
    packed = position_ids is not None and query_states.size(0) > 0 and _is_packed(position_ids)
    if varlen or packed or not is_causal or attention_mask is not None or softcap is not None:
        if s_aux is not None:
            kwargs["s_aux"] = s_aux          # the stock FA path ignores it
        return _original_flash_attention_forward(
            query_states, key_states, value_states, attention_mask, query_length,
            is_causal=is_causal, dropout=dropout, position_ids=position_ids, softmax_scale=softmax_scale,
            sliding_window=sliding_window, use_top_left_mask=use_top_left_mask, softcap=softcap,
            deterministic=deterministic, cu_seq_lens_q=cu_seq_lens_q, cu_seq_lens_k=cu_seq_lens_k,
            max_length_q=max_length_q, max_length_k=max_length_k, target_dtype=target_dtype,
            implementation=implementation, **kwargs)

    n_q, n_kv, h_q = query_states.shape[1], key_states.shape[1], query_states.shape[2]
    q = query_states.transpose(1, 2)         # [B,H,N,D] views; no copies
    k = key_states.transpose(1, 2)
    v = value_states.transpose(1, 2)
    s_local = _local_s_aux(s_aux, h_q, True)
    if n_q != n_kv:                          # cached decode step (reference :98-126)
        out = sink_decode_attention(q, k, v, s_aux=s_local)
        return out.transpose(1, 2).contiguous()
    window = sliding_window if sliding_window is not None else n_q
    out = sink_flash_attention(q, k, v, num_sink=0, window_size=window, s_aux=s_local)
    out = out.transpose(1, 2)
    return out if out.is_contiguous() else out.contiguous()


def patch_verl_with_sink_attention():
    """Swap ``_flash_attention_forward`` in transformers (and in verl's Ulysses wrapper if it is
    imported).  Call before building the trainer/model.  Idempotent (reference :196-239)."""
    global _original_flash_attention_forward
    if _original_flash_attention_forward is not None:
        return
    import transformers.modeling_flash_attention_utils as fa_utils
    _original_flash_attention_forward = fa_utils._flash_attention_forward
    fa_utils._flash_attention_forward = _sink_flash_attention_forward
    try:
        from transformers.integrations import flash_attention
        flash_attention._flash_attention_forward = _sink_flash_attention_forward
    except (ImportError, AttributeError):
        pass
    try:
        import verl.models.transformers.monkey_patch as verl_mp
        verl_mp._flash_attention_forward = _sink_flash_attention_forward
    except (ImportError, AttributeError):
        pass
    print("[SinkAttention] Patched _flash_attention_forward with the sm_100a sink-attention kernels "
          "(s_aux from kwargs, per-layer sliding_window, Ulysses s_aux slicing)")


def unpatch_verl():
    """Restore the original ``_flash_attention_forward`` everywhere it was swapped."""
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
    try:
        import verl.models.transformers.monkey_patch as verl_mp
        verl_mp._flash_attention_forward = _original_flash_attention_forward
    except (ImportError, AttributeError):
        pass
    _original_flash_attention_forward = None
    print("[SinkAttention] Restored original flash attention")
