"""verl / HuggingFace hook: route ``_flash_attention_forward`` to the sink-attention kernels.

Behaviour follows the reference's ``sink_attention/verl_patch.py`` (:33-263): the replacement takes
HF's ``[B, N, H, D]`` tensors, pops gpt-oss's ``s_aux`` from ``**kwargs``, falls back to the saved
original for non-causal / padded / soft-capped calls, routes ``N_q != N_kv`` to the decode kernel,
uses ``num_sink=0`` with ``window = sliding_window or N`` for prefill, slices ``s_aux`` to the local
heads under Ulysses sequence parallelism, and ignores ``softmax_scale`` (the kernels use 1/sqrt(D),
as the reference does).

Difference (packed sequences stay on the sink kernels): the reference sends varlen (``cu_seq_lens``)
and packed (``position_ids`` restarting mid-row) calls to stock FlashAttention (:73-93), which
silently drops ``s_aux`` -- for exactly the padding-free batches verl trains on.  Here the sequence
boundaries go INTO the kernels (``sink_flash_attention_varlen``: a row never attends across a
boundary), and a chunked-prefill call (``1 < N_q < N_kv``), which the reference's decode kernel
rejects (decode_kernel.py:146), runs on ``sink_flash_attention_chunk``.

Difference (layout-fused boundary): the kernels take strides, so the ``transpose(1, 2)`` views are
consumed in place and the output is produced directly in ``[B, N, H, D]`` memory -- the
reference's four ``.contiguous()`` copies per call (:164-166,177) disappear.
"""
from __future__ import annotations

from typing import Optional

import torch

from .decode_kernel import sink_decode_attention
from .sink_flash_attention import (sequence_bounds_from_cu_seqlens, sequence_bounds_from_position_ids,
                                   sink_flash_attention, sink_flash_attention_chunk, sink_flash_attention_varlen)

_original_flash_attention_forward = None


def _is_packed(position_ids: Optional[torch.Tensor]) -> bool:
    """Packed sequences restart their position ids mid-row (reference :182-193).  Host synchronisation (``.item()``):
    the verl hook below no longer calls it -- it derives the sequence bounds on the device instead -- the generation
    hook still does, as the reference."""
    if position_ids is None or position_ids.dim() < 2 or position_ids.size(1) <= 1:
        return False
    return bool((position_ids[:, 1:] < position_ids[:, :-1]).any().item())


# HF hands the SAME position_ids tensor to every layer of a forward pass: derive the bounds once per tensor
_BOUNDS_CACHE = {"key": None, "val": None}


def _bounds_from_position_ids(position_ids: torch.Tensor):
    key = (position_ids.data_ptr(), tuple(position_ids.shape), position_ids._version, position_ids.device)
    if _BOUNDS_CACHE["key"] != key:
        _BOUNDS_CACHE["val"] = sequence_bounds_from_position_ids(position_ids)
        _BOUNDS_CACHE["key"] = key
    return _BOUNDS_CACHE["val"]


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
    n_q, n_kv, h_q = query_states.shape[1], key_states.shape[1], query_states.shape[2]
    # position_ids that restart mid-row mark packed sequences (reference :182-193, a `.item()` host sync per layer
    # call).  Here the bounds are derived ON THE DEVICE and go into the kernels: no sync, and a batch that is not packed
    # simply gets one sequence per row.  (Self-attention calls only; a cached call, N_q != N_kv, is never packed.)
    packed = (not varlen) and position_ids is not None and position_ids.dim() == 2 and n_q == n_kv and n_q > 1 and \
        tuple(position_ids.shape) == (query_states.shape[0], n_q)
    # varlen self-attention only: the same boundaries on both sides (what padding-free training passes)
    varlen_ok = varlen and n_q == n_kv and cu_seq_lens_q.shape == cu_seq_lens_k.shape and \
        (cu_seq_lens_q is cu_seq_lens_k or cu_seq_lens_q.data_ptr() == cu_seq_lens_k.data_ptr() or
         bool(torch.equal(cu_seq_lens_q, cu_seq_lens_k)))
    if (varlen and not varlen_ok) or not is_causal or attention_mask is not None \
            or softcap is not None:
        if s_aux is not None:
            kwargs["s_aux"] = s_aux          # the stock FA path ignores it
        return _original_flash_attention_forward(
            query_states, key_states, value_states, attention_mask, query_length,
            is_causal=is_causal, dropout=dropout, position_ids=position_ids, softmax_scale=softmax_scale,
            sliding_window=sliding_window, use_top_left_mask=use_top_left_mask, softcap=softcap,
            deterministic=deterministic, cu_seq_lens_q=cu_seq_lens_q, cu_seq_lens_k=cu_seq_lens_k,
            max_length_q=max_length_q, max_length_k=max_length_k, target_dtype=target_dtype,
            implementation=implementation, **kwargs)

    q = query_states.transpose(1, 2)         # [B,H,N,D] views; no copies
    k = key_states.transpose(1, 2)
    v = value_states.transpose(1, 2)
    s_local = _local_s_aux(s_aux, h_q, True)
    if varlen or packed:
        # packed sequences: boundaries go into the kernels, s_aux and the per-layer window are kept
        window = sliding_window if sliding_window is not None else n_q
        if varlen:
            B = query_states.shape[0]
            if B > 1:                        # cu_seq_lens index the flattened batch
                q, k, v = (t.transpose(1, 2).reshape(1, B * n_q, t.shape[1], t.shape[3]).transpose(1, 2) for t in (q, k, v))
            bounds = sequence_bounds_from_cu_seqlens(cu_seq_lens_q, q.shape[2])
            window = sliding_window if sliding_window is not None else (max_length_q if isinstance(max_length_q, int) else q.shape[2])
        else:
            bounds = _bounds_from_position_ids(position_ids)
        out = sink_flash_attention_varlen(q, k, v, num_sink=0, window_size=window, s_aux=s_local, seq_bounds=bounds)
        out = out.transpose(1, 2)
        out = out if out.is_contiguous() else out.contiguous()
        return out.reshape(query_states.shape)
    if n_q != n_kv:
        if n_q == 1:                         # cached decode step (reference :98-126)
            out = sink_decode_attention(q, k, v, s_aux=s_local)
            return out.transpose(1, 2).contiguous()
        # chunked prefill: the queries are the last n_q positions of the cached context
        window = sliding_window if sliding_window is not None else n_kv
        out = sink_flash_attention_chunk(q, k, v, num_sink=0, window_size=window, s_aux=s_local)
        out = out.transpose(1, 2)
        return out if out.is_contiguous() else out.contiguous()
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
