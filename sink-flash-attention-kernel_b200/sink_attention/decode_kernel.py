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
