"""Model-level integration tests on the GPU (reference intent: tests/test_gpt_oss_model.py,
tests/test_inference.py:241-271, generate_patch.py:131-168, verl_patch.py:196-239).  The reference's own model tests
need downloaded checkpoints; these use tiny random-weight models of the same architectures:

  * GptOssForCausalLM (sliding + full attention layers, learnable `sinks` -> s_aux) eager vs
    patch_verl_with_sink_attention() + flash_attention_2: logits must agree;
  * LlamaForCausalLM generate() through patch_for_generation (sink + ring KV cache, decode kernel) vs eager generate():
    token-identical.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import sink_attention as sa
from sink_attention import _lib


def _tiny_gpt_oss(dtype, head_dim):
    from transformers import GptOssConfig, GptOssForCausalLM
    cfg = GptOssConfig(vocab_size=128, hidden_size=4 * head_dim, intermediate_size=64, num_hidden_layers=2,
                       layer_types=["sliding_attention", "full_attention"], num_attention_heads=4, num_key_value_heads=2,
                       head_dim=head_dim, sliding_window=8, num_local_experts=2, num_experts_per_tok=1,
                       max_position_embeddings=256, attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    model = GptOssForCausalLM(cfg).to("cuda", dtype).eval()
    with torch.no_grad():
        for layer in model.model.layers:
            layer.self_attn.sinks.normal_(0.0, 1.0)
    return model


@pytest.mark.parametrize("dtype,head_dim,tol", [(torch.float32, 16, 2e-4), (torch.bfloat16, 64, 6e-2)])
def test_tiny_gpt_oss_eager_vs_verl_patch(dtype, head_dim, tol):
    model = _tiny_gpt_oss(dtype, head_dim)
    ids = torch.randint(0, 128, (2, 40), generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        ref = model(ids).logits.float()
    calls = []
    orig_fn, orig_var = sa.verl_patch.sink_flash_attention, sa.verl_patch.sink_flash_attention_varlen

    def spy(q, k, v, num_sink, window_size, s_aux):
        calls.append((tuple(q.shape), num_sink, window_size, None if s_aux is None else tuple(s_aux.shape)))
        return orig_fn(q, k, v, num_sink=num_sink, window_size=window_size, s_aux=s_aux)

    def spy_var(q, k, v, num_sink, window_size, s_aux, seq_bounds):
        # HF hands position_ids to the hook: the sequence bounds are derived on the device (one sequence per row here)
        assert seq_bounds[0].unique().tolist() == [0] and seq_bounds[1].unique().tolist() == [q.shape[2]]
        calls.append((tuple(q.shape), num_sink, window_size, None if s_aux is None else tuple(s_aux.shape)))
        return orig_var(q, k, v, num_sink=num_sink, window_size=window_size, s_aux=s_aux, seq_bounds=seq_bounds)
    sa.patch_verl_with_sink_attention()
    sa.verl_patch.sink_flash_attention = spy
    sa.verl_patch.sink_flash_attention_varlen = spy_var
    try:
        model.config._attn_implementation = "flash_attention_2"
        with torch.no_grad():
            got = model(ids).logits.float()
    finally:
        sa.verl_patch.sink_flash_attention = orig_fn
        sa.verl_patch.sink_flash_attention_varlen = orig_var
        sa.unpatch_verl()
        model.config._attn_implementation = "eager"
    # one call per layer: sliding layer with window 8, full layer with window = N; num_sink = 0; s_aux = sinks [4]
    assert [(c[1], c[2], c[3]) for c in calls] == [(0, 8, (4,)), (0, 40, (4,))], calls
    assert calls[0][0] == (2, 4, 40, head_dim)
    err = (got - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), (err, ref.abs().max().item())
    if dtype == torch.bfloat16:
        assert (got.argmax(-1) == ref.argmax(-1)).float().mean().item() > 0.9


def test_tiny_gpt_oss_verl_patch_backward():
    """Training step through the patch: gradients of the sinks (ds_aux) and of a projection weight vs eager."""
    model = _tiny_gpt_oss(torch.float32, 16).train()
    ids = torch.randint(0, 128, (1, 32), generator=torch.Generator().manual_seed(2)).cuda()

    def grads():
        model.zero_grad(set_to_none=True)
        model(ids).logits.float().square().mean().backward()
        l0 = model.model.layers[0].self_attn
        return l0.sinks.grad.clone(), l0.q_proj.weight.grad.clone(), model.model.layers[1].self_attn.sinks.grad.clone()
    ref = grads()
    sa.patch_verl_with_sink_attention()
    try:
        model.config._attn_implementation = "flash_attention_2"
        got = grads()
    finally:
        sa.unpatch_verl()
        model.config._attn_implementation = "eager"
    for a, b in zip(got, ref):
        assert (a - b).abs().max().item() < 1e-4 * max(1.0, b.abs().max().item()) + 1e-6


@pytest.mark.parametrize("window,num_sink", [(64, 0), (8, 2)])
def test_tiny_llama_generate_through_generation_patch(window, num_sink):
    """window >= total length: sink attention == full causal, so greedy generation must be TOKEN-IDENTICAL to eager.
    window 8 + 2 sinks: the cache evicts; the patched prefill + ring decode must equal the same model run token by
    token through the un-cached sink_flash_attention (last row of a full masked prefill, reference
    tests/test_inference.py:54-199)."""
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=128, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=16, max_position_embeddings=256)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).to("cuda", torch.float32).eval()
    ids = torch.randint(0, 128, (1, 12), generator=torch.Generator().manual_seed(3)).cuda()
    new = 10
    if window >= 12 + new:
        with torch.no_grad():
            ref = model.generate(ids, max_new_tokens=new, do_sample=False)
    else:
        # oracle: no cache -- re-run the whole prefix through sink_flash_attention (prefill kernel) for every new token
        sa.patch_for_generation(model, num_sink=num_sink, window_size=window)
        try:
            model.config._attn_implementation = "flash_attention_2"
            cur = ids
            with torch.no_grad():
                for _ in range(new):
                    logits = model(cur, use_cache=False).logits[:, -1]
                    cur = torch.cat([cur, logits.argmax(-1, keepdim=True)], dim=1)
            ref = cur
        finally:
            sa.unpatch_generation()
            model.config._attn_implementation = "eager"
    cache = sa.patch_for_generation(model, num_sink=num_sink, window_size=window)
    try:
        model.config._attn_implementation = "flash_attention_2"
        with torch.no_grad():
            got = model.generate(ids, max_new_tokens=new, do_sample=False, past_key_values=cache)
    finally:
        sa.unpatch_generation()
        model.config._attn_implementation = "eager"
    assert torch.equal(got, ref), (got.tolist(), ref.tolist())
    assert len(cache) == 2 and cache.seen_tokens == 12 + new - 1
    if window < 12 + new:
        assert cache[0].window_len == window and cache[0].sink_len == num_sink
