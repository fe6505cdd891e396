"""CPU: the oracle replays the golden fixtures that oracle/make_golden.py generated from the
REFERENCE (eager paths + Triton kernels under the interpreter).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import golden_cases as gc
import sink_oracle as orc
from _util import GOLDEN, load_decode, load_prefill, maxdiff


@pytest.mark.parametrize("case", gc.PREFILL_CASES, ids=[c[0] for c in gc.PREFILL_CASES])
def test_oracle_prefill_matches_reference(case):
    (q, k, v, do, s_aux), z = load_prefill(case)
    S, W = case[6], case[7]
    o, lse = orc.sink_attention_fwd(q, k, v, S, W, s_aux)
    dq, dk, dv, ds = orc.sink_attention_bwd(q, k, v, do, S, W, s_aux)
    assert maxdiff(o, torch.from_numpy(z["o"])) < 2e-5
    assert maxdiff(lse, torch.from_numpy(z["lse"])) < 2e-5
    if z["lse_triton"].size:                                      # LSE the reference kernel saved
        assert maxdiff(lse, torch.from_numpy(z["lse_triton"])) < 2e-5
    if z["dq"].size:
        assert maxdiff(dq, torch.from_numpy(z["dq"])) < 2e-5
        assert maxdiff(dk, torch.from_numpy(z["dk"])) < 2e-5
        assert maxdiff(dv, torch.from_numpy(z["dv"])) < 2e-5
    else:
        got = gc.checksum(dq, dk, dv)
        assert abs(got - float(z["grad_checksum"])) < 1e-4 * float(z["grad_checksum"])
    if s_aux is not None:
        assert maxdiff(ds, torch.from_numpy(z["ds_aux"])) < 2e-5


@pytest.mark.parametrize("case", gc.PREFILL_CASES[:6], ids=[c[0] for c in gc.PREFILL_CASES[:6]])
def test_eager_restatement_autograd(case):
    """The eager restatement (what bench.py times as the CPU baseline) reproduces the reference's
    outputs and, through autograd, its gradients."""
    (q, k, v, do, s_aux), z = load_prefill(case)
    S, W = case[6], case[7]
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    sr = s_aux.clone().requires_grad_(True) if s_aux is not None else None
    o = orc.eager_sink_attention(qr, kr, vr, S, W, sr)
    o.backward(do)
    assert maxdiff(o, torch.from_numpy(z["o"])) < 1e-6
    if z["dq"].size:
        assert maxdiff(qr.grad, torch.from_numpy(z["dq"])) < 1e-6
        assert maxdiff(kr.grad, torch.from_numpy(z["dk"])) < 1e-6
        assert maxdiff(vr.grad, torch.from_numpy(z["dv"])) < 1e-6


@pytest.mark.parametrize("case", gc.DECODE_CASES, ids=[c[0] for c in gc.DECODE_CASES])
def test_oracle_decode_matches_reference(case):
    (q, k, v, s_aux), z = load_decode(case)
    o = orc.decode_attention(q, k, v, s_aux)
    assert maxdiff(o, torch.from_numpy(z["o"])) < 2e-5


def test_mask_formula_and_pair_count():
    for (n, s, w) in [(16, 2, 3), (64, 4, 16), (33, 0, 1), (40, 50, 8), (20, 3, 0), (128, 0, 128)]:
        m = orc.attended_mask(n, s, w)
        for i in range(n):
            for j in range(n):
                assert bool(m[i, j]) == ((j <= i) and (j < s or j >= i - w + 1))
        assert int(m.sum()) == orc.attended_pairs(n, s, w)
    # SURVEY.md 8(d) figures
    assert orc.attended_pairs(256, 4, 128) == 25146
    assert orc.attended_pairs(8192, 0, 128) == 1040448


def test_degenerate_rows():
    """window 0 and no sinks: O = 0 and LSE = s_aux; without s_aux LSE = -inf."""
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(1, 2, 9, 16, generator=g) for _ in range(3))
    s_aux = torch.tensor([0.3, -1.2])
    o, lse = orc.sink_attention_fwd(q, k, v, 0, 0, s_aux)
    assert o.abs().max().item() == 0.0
    assert torch.allclose(lse, s_aux.double()[None, :, None].expand_as(lse))
    o, lse = orc.sink_attention_fwd(q, k, v, 0, 0, None)
    assert o.abs().max().item() == 0.0 and bool(torch.isinf(lse).all())


def test_ring_cache_model_matches_reference_traces():
    with open(os.path.join(GOLDEN, "cache_traces.json")) as f:
        data = json.load(f)
    for case in data["cases"]:
        m = orc.RingCacheModel(case["S"], case["W"])
        m.prefill(case["n_prefill"])
        for expect in case["trace"]:
            m.decode()
            assert m.linear() == expect


@pytest.mark.parametrize("shape", [(2, 4, 2, 70, 16, 3, 9), (1, 4, 4, 50, 8, 0, 7), (1, 2, 1, 40, 8, 50, 5),
                                   (1, 2, 1, 40, 8, 4, 0), (1, 2, 2, 33, 8, 2, 1), (1, 2, 1, 30, 8, 0, 64)])
def test_sampled_oracle_matches_full_oracle(shape):
    """The row-sampled oracle (no N x N matrix; used by the GPU suite at the full BASELINE sizes) is the same
    arithmetic as the full oracle, which the golden vectors pin to the reference."""
    B, Hq, Hkv, N, D, S, W = shape
    g = torch.Generator().manual_seed(N + W)
    mk = lambda H: torch.randn(B, H, N, D, generator=g, dtype=torch.float64)
    q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
    s_aux = torch.randn(Hq, generator=g, dtype=torch.float64)
    o, lse = orc.sink_attention_fwd(q, k, v, S, W, s_aux)
    dq, dk, dv, _ = orc.sink_attention_bwd(q, k, v, do, S, W, s_aux)
    rows = torch.stack(torch.meshgrid(torch.arange(B), torch.arange(Hq), torch.arange(N), indexing="ij"), -1).reshape(-1, 3)
    keys = torch.stack(torch.meshgrid(torch.arange(B), torch.arange(Hkv), torch.arange(N), indexing="ij"), -1).reshape(-1, 3)
    o_s, l_s = orc.sampled_fwd(q, k, v, S, W, s_aux, rows)
    dq_s = orc.sampled_dq(q, k, v, do, o, lse, S, W, rows)
    dk_s, dv_s = orc.sampled_dkdv(q, k, v, do, o, lse, S, W, keys)
    for got, ref in ((o_s, o), (dq_s, dq), (dk_s, dk), (dv_s, dv)):
        assert (got - ref.reshape(-1, D)).abs().max().item() < 1e-12
    assert (l_s - lse.reshape(-1)).abs().max().item() < 1e-12


def test_paged_decode_oracle_equals_contiguous_rows():
    """The paged / per-batch-length checker is decode_attention per row: an identity block table over a pool cut from a
    contiguous cache gives the contiguous result, a permuted pool with the matching table gives the same rows, and a
    shorter seq_len equals decoding over the truncated cache."""
    g = torch.Generator().manual_seed(5)
    B, Hq, Hkv, D, page, npg = 3, 8, 2, 16, 4, 5
    N = page * npg
    q = torch.randn(B, Hq, 1, D, generator=g)
    k = torch.randn(B, Hkv, N, D, generator=g)
    v = torch.randn(B, Hkv, N, D, generator=g)
    s_aux = torch.randn(Hq, generator=g)
    full = orc.decode_attention(q, k, v, s_aux)
    pool_k = k.transpose(1, 2).reshape(B * npg, page, Hkv, D)        # [B, N, Hkv, D] -> pages
    pool_v = v.transpose(1, 2).reshape(B * npg, page, Hkv, D)
    table = torch.arange(B * npg).view(B, npg)
    lens = torch.full((B,), N)
    assert torch.allclose(orc.decode_attention_paged(q, pool_k, pool_v, table, lens, s_aux), full, atol=1e-12)
    perm = torch.randperm(B * npg, generator=g)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(B * npg)
    assert torch.allclose(orc.decode_attention_paged(q, pool_k[perm], pool_v[perm], inv[table], lens, s_aux), full, atol=1e-12)
    lens = torch.tensor([7, 0, N])
    out = orc.decode_attention_paged(q, pool_k, pool_v, table, lens, s_aux)
    assert torch.allclose(out[0:1], orc.decode_attention(q[0:1], k[0:1, :, :7], v[0:1, :, :7], s_aux), atol=1e-12)
    assert float(out[1].abs().max()) == 0.0
    assert torch.allclose(out[2:3], full[2:3], atol=1e-12)
    assert torch.allclose(orc.decode_attention_paged(q, k, v, None, lens, s_aux), out, atol=1e-12)
