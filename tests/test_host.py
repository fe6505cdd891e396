"""CPU: host-side logic -- C-ABI exports, cache ring semantics (known answers recorded from the
reference's SinkCacheLayer), patch/unpatch behaviour and routing, s_aux slicing, error behaviour."""
import ctypes
import json
import os
import re

import pytest
import torch

import sink_attention as sa
from sink_attention import _lib, generate_patch, verl_patch
from _util import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "sinkfa.h")) as f:
        hdr = f.read()
    declared = sorted(set(re.findall(r"\b(sfa_[a-z_]+)\s*\(", hdr)))
    assert declared, "no declarations found in include/sinkfa.h"
    for name in declared:
        assert hasattr(lib, name), f"libsinkfa.so does not export {name}"
    assert set(declared) == set(_lib.EXPORTS)
    assert lib.sfa_version() >= 100
    assert lib.sfa_workspace_bytes(_lib.OP_FWD, 1, 8, 8, 256, 64, 0) == 0
    assert lib.sfa_workspace_bytes(_lib.OP_BWD, 1, 8, 8, 256, 64, 0) >= 8 * 256 * 4
    # the product library exports the operator surface only; the micro-probes live in their own library
    assert not any(hasattr(lib, n) for n in ("sfa_probe_umma", "sfa_probe_tma_bw"))
    from sink_attention import _probe
    plib = _probe.load()
    with open(os.path.join(ROOT, "include", "sinkfa_probe.h")) as f:
        phdr = f.read()
    pdecl = sorted(set(re.findall(r"\b(sfa_probe_[a-z_]+)\s*\(", phdr)))
    assert set(pdecl) == set(_probe.EXPORTS)
    for name in pdecl:
        assert hasattr(plib, name), f"libsinkfa_probe.so does not export {name}"


def test_public_api_surface_matches_reference():
    ref_names = ["sink_flash_attention", "patch_verl_with_sink_attention", "unpatch_verl",
                 "prepare_sink_kv_for_sp", "reduce_sink_kv_grads", "SinkAttentionSPWrapper", "SinkCacheLayer",
                 "SinkAttentionCache", "sink_decode_attention", "patch_for_generation", "unpatch_generation",
                 "subprocess_generate"]                        # reference sink_attention/__init__.py:15-28
    for n in ref_names:
        assert hasattr(sa, n) and n in sa.__all__
    import inspect
    sig = inspect.signature(sa.sink_flash_attention)
    assert list(sig.parameters) == ["q", "k", "v", "num_sink", "window_size", "s_aux"]
    assert sig.parameters["num_sink"].default == 4 and sig.parameters["window_size"].default == 512
    sig = inspect.signature(sa.sink_decode_attention)
    assert list(sig.parameters) == ["q", "k", "v", "s_aux"]
    sig = inspect.signature(sa.patch_for_generation)
    assert [p.default for p in sig.parameters.values()] == [None, 4, 4096]


def test_no_cpu_fallback():
    q = torch.randn(1, 2, 8, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sa.sink_flash_attention(q, q, q, 0, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sa.sink_decode_attention(q[:, :, :1], q, q)


def test_argument_errors_come_back_through_the_abi():
    lib = _lib.load()
    arr = (ctypes.c_int64 * 4)(1, 1, 1, 1)
    rc = lib.sfa_fwd(None, None, None, None, None, None, 1, 6, 4, 8, 64, 0, 4, 0, arr, arr, arr, arr, None, 0, None)
    assert rc < 0 and b"divisible" in lib.sfa_last_error()
    rc = lib.sfa_fwd(None, None, None, None, None, None, 1, 4, 4, 8, 64, 0, 4, 0, arr, arr, arr, arr, None, 0, None)
    assert rc < 0 and b"null" in lib.sfa_last_error()
    rc = lib.sfa_set_impl(99)
    assert rc < 0
    # paged decode: argument checks happen before any CUDA call
    a2, a3 = (ctypes.c_int64 * 2)(64, 64), (ctypes.c_int64 * 3)(1, 1, 1)
    dummy = ctypes.c_void_p(16)
    rc = lib.sfa_decode_paged(dummy, dummy, dummy, dummy, None, dummy, None, 1, 8, 2, 100, 64, 0, 48, 4, a2, a3, a3, a2,
                              None, 0, None)
    assert rc < 0 and b"power of two" in lib.sfa_last_error()
    rc = lib.sfa_decode_paged(dummy, None, dummy, dummy, None, dummy, None, 1, 8, 2, 100, 64, 0, 64, 4, a2, a3, a3, a2,
                              None, 0, None)
    assert rc < 0 and b"null" in lib.sfa_last_error()


def test_paged_decode_wrappers_refuse_cpu_tensors():
    q = torch.randn(2, 8, 1, 64)
    pool = torch.randn(4, 32, 2, 64)
    bt = torch.zeros(2, 2, dtype=torch.int32)
    sl = torch.tensor([10, 40], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sa.sink_decode_attention_paged(q, pool, pool, bt, sl)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sa.sink_decode_attention_varlen(q, torch.randn(2, 2, 40, 64), torch.randn(2, 2, 40, 64), sl)


# ------------------------------------------------------------------------------------------------ cache
def _ids(lo, hi):
    return torch.arange(lo, hi, dtype=torch.float32).view(1, 1, -1, 1)


def test_cache_traces_recorded_from_reference():
    with open(os.path.join(GOLDEN, "cache_traces.json")) as f:
        data = json.load(f)
    for case in data["cases"]:
        layer = sa.SinkCacheLayer(case["S"], case["W"])
        n0 = case["n_prefill"]
        k_out, v_out = layer.update(_ids(0, n0), _ids(0, n0) + 0.5)
        assert k_out.shape[2] == n0                               # prefill returns full K/V
        for t, expect in enumerate(case["trace"]):
            k_out, v_out = layer.update(_ids(n0 + t, n0 + t + 1), _ids(n0 + t, n0 + t + 1) + 0.5)
            assert [int(x) for x in k_out.flatten().tolist()] == expect
            assert torch.equal(v_out, k_out + 0.5)
            assert layer.get_seq_length() == len(expect)
        assert layer.get_max_cache_shape() == case["S"] + case["W"]


def test_cache_container_and_hf_protocol():
    cache = sa.SinkAttentionCache(num_sink=2, window_size=4)
    assert len(cache) == 0 and cache.get_seq_length() == 0
    for layer in range(3):
        k, v = cache.update(_ids(0, 5).expand(2, 3, 5, 8).clone(), _ids(0, 5).expand(2, 3, 5, 8).clone(), layer)
        assert k.shape == (2, 3, 5, 8)
    assert len(cache) == 3 and cache.seen_tokens == 5
    assert cache.get_seq_length() == 2 + 3 and cache.get_max_cache_length() == 6
    k, v = cache.update(_ids(5, 6).expand(2, 3, 1, 8).clone(), _ids(5, 6).expand(2, 3, 1, 8).clone(), 0)
    assert k.shape[2] == 6 and cache.seen_tokens == 6
    assert k[0, 0, :, 0].tolist() == [0, 1, 2, 3, 4, 5]
    k, v = cache.update(_ids(6, 7).expand(2, 3, 1, 8).clone(), _ids(6, 7).expand(2, 3, 1, 8).clone(), 0)
    assert k[0, 0, :, 0].tolist() == [0, 1, 3, 4, 5, 6]            # token 2 evicted, sinks kept
    # multi-token decode goes through the ring token by token
    k, v = cache.update(_ids(7, 10).expand(2, 3, 3, 8).clone(), _ids(7, 10).expand(2, 3, 3, 8).clone(), 0)
    assert k[0, 0, :, 0].tolist() == [0, 1, 6, 7, 8, 9]
    cache.reorder_cache(torch.tensor([1, 0]))
    assert cache[0].get_mask_sizes(None) == (6, 0)
    assert "SinkAttentionCache(num_sink=2" in repr(cache)


def test_cache_works_without_is_initialized_from_hf():
    layer = sa.SinkCacheLayer(1, 2)
    assert layer.is_initialized is False
    layer.update(_ids(0, 1), _ids(0, 1))
    assert layer.is_initialized and layer.sink_len == 1 and layer.window_len == 0


# ------------------------------------------------------------------------------------------------ patches
def test_verl_patch_swaps_and_restores():
    import transformers.modeling_flash_attention_utils as fa_utils
    from transformers.integrations import flash_attention as fa_int
    orig = fa_utils._flash_attention_forward
    sa.patch_verl_with_sink_attention()
    try:
        assert fa_utils._flash_attention_forward is verl_patch._sink_flash_attention_forward
        assert fa_int._flash_attention_forward is verl_patch._sink_flash_attention_forward
        sa.patch_verl_with_sink_attention()                      # idempotent: original not overwritten
        assert verl_patch._original_flash_attention_forward is orig
    finally:
        sa.unpatch_verl()
    assert fa_utils._flash_attention_forward is orig and fa_int._flash_attention_forward is orig


def test_generation_patch_swaps_and_restores():
    import transformers.modeling_flash_attention_utils as fa_utils
    orig = fa_utils._flash_attention_forward
    cache = sa.patch_for_generation(None, num_sink=3, window_size=77)
    try:
        assert isinstance(cache, sa.SinkAttentionCache) and (cache.num_sink, cache.window_size) == (3, 77)
        assert fa_utils._flash_attention_forward is generate_patch._generation_flash_attention_forward
        assert generate_patch._GENERATION_CONFIG == {"num_sink": 3, "window_size": 77, "enabled": True}
        sa.patch_for_generation(None, 4, 4096)                   # re-patch keeps the true original
        assert generate_patch._original_flash_attention_forward is orig
    finally:
        sa.unpatch_generation()
    assert fa_utils._flash_attention_forward is orig
    assert generate_patch._GENERATION_CONFIG["enabled"] is False


def test_verl_forward_routing(monkeypatch):
    """Fallback conditions (reference verl_patch.py:73-93), decode routing (:98-126), num_sink=0 and
    window = sliding_window or N (:158-174), s_aux popped from kwargs and sliced for local heads."""
    calls = []
    monkeypatch.setattr(verl_patch, "_original_flash_attention_forward",
                        lambda *a, **kw: calls.append(("orig", kw)) or "orig")
    monkeypatch.setattr(verl_patch, "sink_flash_attention",
                        lambda q, k, v, num_sink, window_size, s_aux: calls.append(
                            ("prefill", tuple(q.shape), num_sink, window_size, s_aux)) or q)
    monkeypatch.setattr(verl_patch, "sink_decode_attention",
                        lambda q, k, v, s_aux=None: calls.append(("decode", tuple(q.shape), tuple(k.shape), s_aux)) or q)
    monkeypatch.setattr(verl_patch, "sink_flash_attention_chunk",
                        lambda q, k, v, num_sink, window_size, s_aux: calls.append(
                            ("chunk", tuple(q.shape), tuple(k.shape), num_sink, window_size, s_aux)) or q)
    monkeypatch.setattr(verl_patch, "sink_flash_attention_varlen",
                        lambda q, k, v, num_sink, window_size, s_aux, seq_bounds: calls.append(
                            ("varlen", tuple(q.shape), num_sink, window_size, s_aux, seq_bounds)) or q)
    f = verl_patch._sink_flash_attention_forward
    q = torch.zeros(2, 10, 8, 16)          # HF layout [B, N, H, D]
    kv = torch.zeros(2, 10, 2, 16)
    s_aux = torch.arange(8.0)
    out = f(q, kv, kv, None, 10, sliding_window=4, s_aux=s_aux, layer_idx=0, attn_implementation="x")
    assert out.shape == (2, 10, 8, 16)
    assert calls[-1][:4] == ("prefill", (2, 8, 10, 16), 0, 4) and calls[-1][4] is s_aux
    f(q, kv, kv, None, 10, sliding_window=None)
    assert calls[-1][:4] == ("prefill", (2, 8, 10, 16), 0, 10) and calls[-1][4] is None
    f(q, kv, kv, None, 10, softmax_scale=0.5)                     # scale is ignored, not a fallback
    assert calls[-1][0] == "prefill"
    # Ulysses: s_aux holds 16 heads, this rank has 8 -> rank 0 slice (no dist, no verl)
    f(q, kv, kv, None, 10, s_aux=torch.arange(16.0))
    assert calls[-1][4].tolist() == list(range(8))
    f(q, kv, kv, None, 10, s_aux=torch.arange(5.0))              # mismatch -> dropped
    assert calls[-1][4] is None
    # decode: N_q == 1 over a longer cache
    f(q[:, :1], kv, kv, None, 1, s_aux=s_aux)
    assert calls[-1][:3] == ("decode", (2, 8, 1, 16), (2, 2, 10, 16))
    # chunked prefill (1 < N_q < N_kv): the reference's decode kernel asserts N_q == 1 (decode_kernel.py:146); here the
    # chunk kernel takes it, still with num_sink = 0, window = sliding_window or N_kv, s_aux
    f(q[:, :4], kv, kv, None, 4, s_aux=s_aux, sliding_window=6)
    assert calls[-1][:5] == ("chunk", (2, 8, 4, 16), (2, 2, 10, 16), 0, 6) and calls[-1][5] is s_aux
    # packed sequences stay on the sink kernels WITH s_aux (the reference falls back to stock FA and drops it, :73-93)
    cu = torch.tensor([0, 4, 10], dtype=torch.int32)
    f(q[:1], kv[:1], kv[:1], None, 10, s_aux=s_aux, sliding_window=5, cu_seq_lens_q=cu, cu_seq_lens_k=cu,
      max_length_q=6, max_length_k=6)
    assert calls[-1][:4] == ("varlen", (1, 8, 10, 16), 0, 5) and calls[-1][4] is s_aux
    lo, hi = calls[-1][5]
    assert lo.tolist() == [0, 0, 0, 0, 4, 4, 4, 4, 4, 4] and hi.tolist() == [4, 4, 4, 4, 10, 10, 10, 10, 10, 10]
    f(q, kv, kv, None, 10, s_aux=s_aux, position_ids=torch.tensor([[0, 1, 2, 0, 1, 2, 3, 4, 5, 6], [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]]))
    assert calls[-1][:4] == ("varlen", (2, 8, 10, 16), 0, 10)
    lo, hi = calls[-1][5]
    assert lo.tolist() == [[0, 0, 0, 3, 3, 3, 3, 3, 3, 3], [0] * 8 + [8, 8]]
    assert hi.tolist() == [[3, 3, 3, 10, 10, 10, 10, 10, 10, 10], [8] * 8 + [10, 10]]
    # fallbacks to the saved original: non-causal, soft-capped, padded, cross-length varlen
    cu_k = torch.tensor([0, 5, 10], dtype=torch.int32)
    for kw in (dict(is_causal=False), dict(softcap=30.0),
               dict(cu_seq_lens_q=cu, cu_seq_lens_k=cu_k, max_length_q=6, max_length_k=6)):
        assert f(q[:1], kv[:1], kv[:1], None, 10, s_aux=s_aux, **kw) == "orig"
        assert calls[-1][0] == "orig" and calls[-1][1]["s_aux"] is s_aux
    assert f(q, kv, kv, torch.ones(2, 10), 10) == "orig"          # padding mask
    # monotone position ids: one sequence per row -- same kernels, bounds [0, N); no host sync decides this any more
    f(q, kv, kv, None, 10, position_ids=torch.arange(10)[None].expand(2, -1))
    assert calls[-1][0] == "varlen" and calls[-1][5][0].tolist() == [[0] * 10] * 2 and calls[-1][5][1].tolist() == [[10] * 10] * 2
    # a cached call (N_q != N_kv) with position ids goes to the chunk / decode kernels
    f(q[:, :4], kv, kv, None, 4, position_ids=torch.arange(6, 10)[None].expand(2, -1))
    assert calls[-1][0] == "chunk"


def test_sequence_bounds_helpers():
    from sink_attention.sink_flash_attention import sequence_bounds_from_cu_seqlens, sequence_bounds_from_position_ids
    lo, hi = sequence_bounds_from_cu_seqlens(torch.tensor([0, 3, 3, 7]), 9)       # an empty sequence, 2 padding slots
    assert lo.tolist() == [0, 0, 0, 3, 3, 3, 3, 7, 8] and hi.tolist() == [3, 3, 3, 7, 7, 7, 7, 8, 9]
    lo, hi = sequence_bounds_from_position_ids(torch.tensor([[0, 1, 0, 0, 1, 2]]))
    assert lo.tolist() == [[0, 0, 2, 3, 3, 3]] and hi.tolist() == [[2, 2, 3, 6, 6, 6]]
    assert lo.dtype == torch.int32 and hi.dtype == torch.int32


def test_generation_forward_routing(monkeypatch):
    calls = []
    monkeypatch.setattr(generate_patch, "_original_flash_attention_forward", lambda *a, **kw: "orig")
    monkeypatch.setattr(generate_patch, "sink_flash_attention",
                        lambda q, k, v, num_sink, window_size: calls.append(("prefill", num_sink, window_size)) or q)
    monkeypatch.setattr(generate_patch, "sink_decode_attention", lambda q, k, v: calls.append(("decode",)) or q)
    generate_patch._GENERATION_CONFIG.update(num_sink=2, window_size=9)
    f = generate_patch._generation_flash_attention_forward
    q, kv = torch.zeros(1, 6, 4, 8), torch.zeros(1, 6, 2, 8)
    assert f(q, kv, kv, torch.ones(1, 6), 6, sliding_window=3, softcap=1.0, s_aux=torch.zeros(4)).shape == (1, 6, 4, 8)
    assert calls[-1] == ("prefill", 2, 9)          # attention_mask / sliding_window / softcap / s_aux ignored
    f(q[:, :1], kv, kv, None, 1)
    assert calls[-1] == ("decode",)
    assert f(q, kv, kv, None, 6, is_causal=False) == "orig"
    generate_patch._GENERATION_CONFIG.update(num_sink=4, window_size=4096)
