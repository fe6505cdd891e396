"""KV cache for sink-attention inference: a fixed sink buffer plus a ring (circular) window buffer.

Behaviour-compatible with the reference's ``SinkCacheLayer`` / ``SinkAttentionCache``
(sink_attention/cache.py:29-330): same constructor arguments, same ``update()`` contract (prefill
stores the first ``num_sink`` and the last ``window_size`` tokens and returns the FULL K/V; decode
appends at ``write_pos`` and returns the chronologically linearised ``[sink, window]`` K/V), same
HF ``Cache`` glue.  Additions for the B200 path:

  * ``decode_attention(q, s_aux)`` attends the two buffers IN PLACE through ``sfa_decode_ring``
    (softmax is order-invariant, RoPE is already applied), skipping the per-step linearisation copy
    that costs more bytes than the attention itself (reference cache.py:185-216).
  * ``is_initialized`` exists without transformers too (the reference only gets it from the HF mixin).

  * ``append()`` is ONE kernel launch for K and V (``sfa_cache_append``) on CUDA tensors.
  * ring fast path: under ``patch_for_generation`` a decode-step ``update()`` returns zero-copy VIEWS of the ring
    instead of the linearised copy, and registers the layer under the view's data pointer; the generation hook
    finds the layer from the key tensor it receives and calls ``decode_attention`` on the buffers in place.
    Direct users of ``update()`` / ``get_kv()`` (no patch active) keep the chronological contract.

All batch rows share one sequence length; ring state is host-side integers (as in the reference).
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Tuple

import torch

# ring buffers that a decode-step update() handed out as views while the generation patch was active:
# data_ptr of window_k -> SinkCacheLayer (weak).  The generation hook looks its key tensor up here.
_RING_VIEWS: "weakref.WeakValueDictionary[int, SinkCacheLayer]" = weakref.WeakValueDictionary()
_FAST_DECODE = {"enabled": False}      # switched by patch_for_generation / unpatch_generation


def ring_layer_of(key_states: torch.Tensor) -> Optional["SinkCacheLayer"]:
    """The cache layer whose ring buffer `key_states` is a view of (None if it is an ordinary tensor)."""
    if not _FAST_DECODE["enabled"] or not _RING_VIEWS:
        return None
    layer = _RING_VIEWS.get(key_states.data_ptr())
    if layer is None or layer.window_k is None or layer.window_k.data_ptr() != key_states.data_ptr():
        return None
    return layer

try:  # HF is optional: without it the classes are plain Python objects
    from transformers.cache_utils import Cache as _HFCache, CacheLayerMixin as _HFLayer
    _HAS_HF = True
except Exception:  # pragma: no cover - transformers missing or too old
    _HFCache, _HFLayer, _HAS_HF = object, object, False


_APPEND_DTYPES = (torch.bfloat16, torch.float16, torch.float32)


class SinkCacheLayer(_HFLayer):
    """One layer: ``sink_k/v [B,H_kv,num_sink,D]`` + ring ``window_k/v [B,H_kv,window_size,D]``."""

    is_sliding = True

    def __init__(self, num_sink: int, window_size: int):
        if _HAS_HF:
            super().__init__()
        self.is_initialized = False
        self.num_sink = int(num_sink)
        self.window_size = int(window_size)
        self.sink_k: Optional[torch.Tensor] = None
        self.sink_v: Optional[torch.Tensor] = None
        self.window_k: Optional[torch.Tensor] = None
        self.window_v: Optional[torch.Tensor] = None
        self.sink_len = 0        # populated sink slots
        self.window_len = 0      # populated ring slots
        self.write_pos = 0       # ring slot the next decoded token goes to
        self.prefilled = False
        self.seen_tokens = 0

    # -- allocation -------------------------------------------------------------------------
    def lazy_initialization(self, key_states: torch.Tensor, value_states: Optional[torch.Tensor] = None):
        B, H, _, D = key_states.shape
        opts = dict(dtype=key_states.dtype, device=key_states.device)
        self.sink_k = torch.zeros(B, H, self.num_sink, D, **opts)
        self.sink_v = torch.zeros(B, H, self.num_sink, D, **opts)
        self.window_k = torch.zeros(B, H, self.window_size, D, **opts)
        self.window_v = torch.zeros(B, H, self.window_size, D, **opts)
        self.is_initialized = True

    # -- prefill ----------------------------------------------------------------------------
    def _prefill(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        n = k.shape[2]
        S, W = self.num_sink, self.window_size
        self.seen_tokens = n
        ns = min(n, S)
        if ns:
            self.sink_k[:, :, :ns].copy_(k[:, :, :ns])
            self.sink_v[:, :, :ns].copy_(v[:, :, :ns])
        self.sink_len = ns
        rest = n - ns                      # tokens after the sinks
        keep = min(rest, W)                # the most recent `keep` of them live in the ring
        if keep:
            self.window_k[:, :, :keep].copy_(k[:, :, n - keep:])
            self.window_v[:, :, :keep].copy_(v[:, :, n - keep:])
        self.window_len = keep
        # ring full -> next write wraps to slot 0; otherwise continue after the last filled slot
        self.write_pos = 0 if (W == 0 or keep == W) else keep
        self.prefilled = True
        return k, v                        # the prefill kernel masks over the full sequence

    # -- decode -----------------------------------------------------------------------------
    def append(self, k: torch.Tensor, v: torch.Tensor) -> None:
        """Write one token ([B,H_kv,1,D]) into the ring, evicting the oldest when full."""
        self.seen_tokens += 1
        if self.window_k.is_cuda and self.window_k.dtype in _APPEND_DTYPES:
            from . import _lib
            _lib.cache_append(k, v, self.window_k, self.window_v, self.write_pos)      # K and V, one launch
        else:
            self.window_k[:, :, self.write_pos].copy_(k[:, :, 0])
            self.window_v[:, :, self.write_pos].copy_(v[:, :, 0])
        self.write_pos = (self.write_pos + 1) % self.window_size
        self.window_len = min(self.window_len + 1, self.window_size)

    def _decode(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        self.append(k, v)
        if _FAST_DECODE["enabled"] and self.window_k.is_cuda and k.shape[2] == 1:
            # generation fast path: no linearisation copy (reference cache.py:185-216 copies the whole cache every
            # step).  The views carry the right LENGTH (sink_len + window_len keys) for HF's bookkeeping only when
            # sink_len == 0, so the hook never reads them: it resolves the layer and attends the buffers in place.
            _RING_VIEWS[self.window_k.data_ptr()] = self
            return self.window_k[:, :, :self.window_len], self.window_v[:, :, :self.window_len]
        return self.get_kv()

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor,
               cache_kwargs: Optional[dict] = None, *args, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        if not self.is_initialized:
            self.lazy_initialization(key_states, value_states)
        if not self.prefilled:
            return self._prefill(key_states, value_states)
        out = None
        for t in range(key_states.shape[2]):       # multi-token decode goes token by token
            out = self._decode(key_states[:, :, t:t + 1], value_states[:, :, t:t + 1])
        return out

    def get_kv(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Chronological ``[sink, window]`` K/V, oldest first: ``[B,H_kv,sink_len+window_len,D]``."""
        ks, vs = [self.sink_k[:, :, :self.sink_len]], [self.sink_v[:, :, :self.sink_len]]
        wl, wp = self.window_len, self.write_pos
        if wl:
            if wl < self.window_size or wp == 0:
                ks.append(self.window_k[:, :, :wl])
                vs.append(self.window_v[:, :, :wl])
            else:                                   # full ring: oldest entry sits at write_pos
                ks += [self.window_k[:, :, wp:], self.window_k[:, :, :wp]]
                vs += [self.window_v[:, :, wp:], self.window_v[:, :, :wp]]
        return torch.cat(ks, dim=2), torch.cat(vs, dim=2)

    def decode_attention(self, q: torch.Tensor, s_aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Attention of one query token [B,H_q,1,D] over the resident cache, read in place."""
        from . import _lib
        s32 = _lib._s_aux_f32(s_aux, q.shape[1])
        return _lib.decode_ring(q, self.sink_k, self.sink_v, self.window_k, self.window_v,
                                self.sink_len, self.window_len, s32)

    # -- HF CacheLayer protocol -------------------------------------------------------------
    def get_seq_length(self, *args, **kwargs) -> int:
        return self.sink_len + self.window_len

    def get_mask_sizes(self, cache_position=None, *args, **kwargs) -> Tuple[int, int]:
        return self.get_seq_length(), 0

    def get_max_cache_shape(self) -> int:
        return self.num_sink + self.window_size

    def reorder_cache(self, beam_idx: torch.LongTensor):
        if self.sink_k is None:
            return
        idx = beam_idx.to(self.sink_k.device)
        self.sink_k = self.sink_k.index_select(0, idx)
        self.sink_v = self.sink_v.index_select(0, idx)
        self.window_k = self.window_k.index_select(0, idx)
        self.window_v = self.window_v.index_select(0, idx)

    def reset(self) -> None:
        for t in (self.sink_k, self.sink_v, self.window_k, self.window_v):
            if t is not None:
                t.zero_()
        self.sink_len = self.window_len = self.write_pos = self.seen_tokens = 0
        self.prefilled = False

    def __repr__(self):
        return (f"SinkCacheLayer(num_sink={self.num_sink}, window_size={self.window_size}, "
                f"sink_len={self.sink_len}, window_len={self.window_len}, write_pos={self.write_pos})")


class SinkAttentionCache(_HFCache):
    """Per-model cache: a lazily grown list of ``SinkCacheLayer`` usable as HF ``past_key_values``."""

    def __init__(self, num_sink: int = 4, window_size: int = 4096):
        self.num_sink = num_sink
        self.window_size = window_size
        self._seen_tokens = 0
        if _HAS_HF:
            super().__init__(layers=[])
        else:
            self.layers: List[SinkCacheLayer] = []

    def __len__(self) -> int:
        return len(self.layers)

    def __getitem__(self, idx: int) -> SinkCacheLayer:
        return self.layers[idx]

    def __repr__(self) -> str:
        return (f"SinkAttentionCache(num_sink={self.num_sink}, window_size={self.window_size}, "
                f"layers={len(self.layers)}, seen_tokens={self._seen_tokens})")

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, layer_idx: int,
               cache_kwargs: Optional[dict] = None, *args, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        while len(self.layers) <= layer_idx:
            self.layers.append(SinkCacheLayer(self.num_sink, self.window_size))
        out = self.layers[layer_idx].update(key_states, value_states, cache_kwargs)
        if layer_idx == 0:
            self._seen_tokens = self.layers[0].seen_tokens
        return out

    def get_seq_length(self, layer_idx: int = 0, *args, **kwargs) -> int:
        return self.layers[layer_idx].get_seq_length() if layer_idx < len(self.layers) else 0

    def get_max_cache_length(self) -> int:
        return self.num_sink + self.window_size

    def get_max_cache_shape(self, layer_idx: int = 0) -> int:
        return self.num_sink + self.window_size

    def get_mask_sizes(self, cache_position=None, layer_idx: int = 0, *args, **kwargs) -> Tuple[int, int]:
        if layer_idx < len(self.layers):
            return self.layers[layer_idx].get_mask_sizes(cache_position)
        n = 0 if cache_position is None else int(getattr(cache_position, "shape", [cache_position])[0])
        return n, 0

    def reorder_cache(self, beam_idx: torch.LongTensor):
        for layer in self.layers:
            layer.reorder_cache(beam_idx)

    @property
    def seen_tokens(self) -> int:
        return self._seen_tokens
