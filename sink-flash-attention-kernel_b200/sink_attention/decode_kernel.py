"""sink_decode_attention(q, k, v, s_aux) -- single-query attention over the sink+window cache.

Mirrors the reference entry point (sink_attention/decode_kernel.py:120-226).  The reference runs a
Triton split-KV pass per Q head and reduces the partials with ~12 torch ops; here one kernel pass
serves every Q head of a GQA group per KV read and a small combine kernel folds in ``s_aux`` as
the virtual split (m = s_aux, l = 1, o = 0).
"""
from __future__ import annotations

import torch

from . import _lib


def sink_decode_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, s_aux: torch.Tensor = None) -> torch.Tensor:
    """q [B,H_q,1,D]; k,v [B,H_kv,N_kv,D]; s_aux [H_q] or None -> [B,H_q,1,D] in q's dtype."""
    B, H_q, N_q, D = q.shape
    H_kv = k.shape[1]
    assert N_q == 1, f"sink_decode_attention requires N_q=1, got {N_q}"                 # reference :146
    assert H_q % H_kv == 0, f"H_q ({H_q}) must be divisible by H_kv ({H_kv})"          # reference :147
    # the reference additionally needs a power-of-two D (tl.arange, :148-149); any D <= 256 works here
    assert 1 <= D <= 256, f"D={D} must be in [1, 256]"
    assert k.shape == v.shape and k.shape[0] == B and k.shape[3] == D
    s_aux_f32 = _lib._s_aux_f32(s_aux, H_q)
    return _lib.decode(q, k, v, s_aux_f32)


def sink_decode_attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, seq_lens: torch.Tensor,
                                 s_aux: torch.Tensor = None) -> torch.Tensor:
    """Per-batch cache lengths (the reference shares one length across the batch,
    cache.py:11-13): q [B,H_q,1,D]; k,v [B,H_kv,N_max,D]; seq_lens int32 [B] on the device -- row b attends
    k[b, :, :seq_lens[b]].  A row with no key yields 0 (only the s_aux sink, if given, is attended)."""
    B, H_q, N_q, D = q.shape
    assert N_q == 1, f"sink_decode_attention_varlen requires N_q=1, got {N_q}"
    assert H_q % k.shape[1] == 0 and k.shape == v.shape and k.shape[0] == B and k.shape[3] == D
    return _lib.decode_paged(q, k, v, _lib._s_aux_f32(s_aux, H_q), seq_lens=seq_lens)


def sink_decode_attention_paged(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, block_table: torch.Tensor,
                                seq_lens: torch.Tensor, s_aux: torch.Tensor = None, max_len: int = None) -> torch.Tensor:
    """Paged KV cache (row f4): k_cache, v_cache [num_pages, page_size, H_kv, D] (page_size a power of two >= 32);
    block_table int32 [B, max_pages]: physical page of each logical page; seq_lens int32 [B].  ``max_len`` (an upper
    bound of seq_lens, host-known) shortens the planned key range; default: the block table's capacity."""
    B, H_q, N_q, D = q.shape
    assert N_q == 1, f"sink_decode_attention_paged requires N_q=1, got {N_q}"
    assert k_cache.shape == v_cache.shape and k_cache.dim() == 4 and k_cache.shape[3] == D
    assert H_q % k_cache.shape[2] == 0 and block_table.shape[0] == B
    return _lib.decode_paged(q, k_cache, v_cache, _lib._s_aux_f32(s_aux, H_q), seq_lens=seq_lens,
                             block_table=block_table, max_len=max_len)
