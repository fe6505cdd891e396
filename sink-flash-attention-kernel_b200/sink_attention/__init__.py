"""sink_attention -- B200 (sm_100a) native drop-in for RulinShao/sink-flash-attention-kernel.

Same 12 public names as the reference package (sink_attention/__init__.py:1-28); the Triton kernels
are replaced by hand-written CUDA behind the C ABI in ``include/sinkfa.h`` (``libsinkfa.so``).
"""
from .sink_flash_attention import (
    sink_flash_attention,
    sink_flash_attention_with_lse,
    sink_flash_attention_varlen,
    sink_flash_attention_chunk,
    SinkFlashAttentionFunc,
)
from .verl_patch import patch_verl_with_sink_attention, unpatch_verl
from .sp_utils import (
    prepare_sink_kv_for_sp,
    reduce_sink_kv_grads,
    SinkAttentionSPWrapper,
    UlyssesSinkAttention,
    HaloSinkAttention,
    ulysses_seq_to_head,
    ulysses_head_to_seq,
)
from .cache import SinkCacheLayer, SinkAttentionCache
from .decode_kernel import sink_decode_attention, sink_decode_attention_varlen, sink_decode_attention_paged
from .generate_patch import patch_for_generation, unpatch_generation
from .subprocess_eval import subprocess_generate

__version__ = "0.1.0"

__all__ = [
    "sink_flash_attention",
    "patch_verl_with_sink_attention",
    "unpatch_verl",
    "prepare_sink_kv_for_sp",
    "reduce_sink_kv_grads",
    "SinkAttentionSPWrapper",
    "SinkCacheLayer",
    "SinkAttentionCache",
    "sink_decode_attention",
    "patch_for_generation",
    "unpatch_generation",
    "subprocess_generate",
    # B200 additions
    "sink_flash_attention_with_lse",
    "sink_flash_attention_varlen",
    "sink_flash_attention_chunk",
    "UlyssesSinkAttention",
    "HaloSinkAttention",
    "ulysses_seq_to_head",
    "ulysses_head_to_seq",
    "sink_decode_attention_varlen",
    "sink_decode_attention_paged",
]
