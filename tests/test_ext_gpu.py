"""Packed (varlen) sequences and chunked prefill inside the kernels: parity with the oracle applied
sequence by sequence / to the full context.  The reference hands packed calls to stock FlashAttention (dropping s_aux,
verl_patch.py:73-93) and asserts on 1 < N_q < N_kv (decode_kernel.py:146); the semantics checked here are the
reference's own mask (sink_flash_attention.py:30-39) restricted to each sequence."""
import pytest
import torch

import sink_oracle as orc
from _util import excess, maxdiff

pytestmark = pytest.mark.gpu

import sink_attention as sa
from sink_attention import _lib


def _oracle_packed(q, k, v, do, lens, S, W, s_aux):
    """per-sequence oracle, concatenated along N (inputs on the CPU, [B=1,H,N,D])."""
    outs = {n: [] for n in ("o", "lse", "dq", "dk", "dv")}
    ds = None
    p0 = 0
    for L in lens:
        sl = slice(p0, p0 + L)
        qs, ks, vs, dos = q[:, :, sl], k[:, :, sl], v[:, :, sl], do[:, :, sl]
        o, lse = orc.sink_attention_fwd(qs, ks, vs, S, W, s_aux)
        dq, dk, dv, d = orc.sink_attention_bwd(qs, ks, vs, dos, S, W, s_aux)
        for n, t in zip(("o", "lse", "dq", "dk", "dv"), (o, lse, dq, dk, dv)):
            outs[n].append(t)
        if d is not None:
            ds = d if ds is None else ds + d
        p0 += L
    return {n: torch.cat(t, dim=2) for n, t in outs.items()}, ds


@pytest.mark.parametrize("dtype,Hq,Hkv,D,S,W,lens,impls", [
    # gpt-oss training shape class: head_dim 64, no sink tokens, narrow window -> tcgen05 forward + fused backward
    (torch.bfloat16, 16, 2, 64, 0, 128, [300, 17, 1, 512, 130], ("tcgen05", "tcgen05-fused")),
    (torch.float16, 8, 1, 64, 0, 32, [64, 64, 70], ("tcgen05", "tcgen05-fused")),
    # full-attention layer over a packed batch (window = longest sequence): tcgen05 forward, dQ + dK/dV tensor-core pair
    (torch.bfloat16, 8, 2, 64, 0, 600, [600, 40, 333], ("tcgen05", "tcgen05")),
    (torch.float16, 4, 4, 64, 0, 2000, [1, 700, 299, 1000], ("tcgen05", "tcgen05")),
    # per-sequence sink tokens, other head dims, fp32: CUDA-core kernels with the full predicate
    (torch.bfloat16, 8, 2, 64, 3, 16, [50, 2, 90], ("simt", "simt")),
    (torch.float32, 4, 4, 32, 2, 9, [33, 70, 5], ("simt", "simt")),
    # head_dim 128 (and 80 on the same kernels): one-tile-per-CTA forward + kernel pair, all with the sequence bounds
    (torch.bfloat16, 8, 2, 128, 0, 64, [100, 156], ("tcgen05", "tcgen05")),
    (torch.bfloat16, 8, 2, 128, 0, 512, [300, 212, 77], ("tcgen05", "tcgen05")),
    (torch.bfloat16, 4, 1, 80, 0, 100, [130, 90], ("tcgen05", "tcgen05")),
    # long packed sequences at head_dim 64 with a wide window: the two-tile forward (fwd128_kernel<T, 64>) and the
    # D-generic dK/dV kernel, both with the sequence bounds
    (torch.bfloat16, 8, 2, 64, 0, 2048, [1300, 1700], ("tcgen05", "tcgen05")),
    (torch.bfloat16, 8, 1, 128, 0, 4096, [900, 1, 1400], ("tcgen05", "tcgen05")),
])
def test_packed_sequences_match_per_sequence_oracle(dtype, Hq, Hkv, D, S, W, lens, impls):
    N = sum(lens)
    g = torch.Generator().manual_seed(N + W)
    mk = lambda H: torch.randn(1, N, H, D, generator=g).to(dtype)
    q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)                       # HF layout, as the patch receives it
    s_aux = torch.randn(Hq, generator=g) * 0.5
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    qd, kd, vd = (t.cuda().transpose(1, 2).requires_grad_(True) for t in (q, k, v))
    sd = s_aux.cuda().requires_grad_(True)
    o = sa.sink_flash_attention_varlen(qd, kd, vd, cu, S, W, sd)
    fwd_impl = _lib.last_impl()
    o.backward(do.cuda().transpose(1, 2))
    bwd_impl = _lib.last_impl()
    assert (fwd_impl, bwd_impl) == impls
    ref, ds_ref = _oracle_packed(*(t.transpose(1, 2).float() for t in (q, k, v, do)), lens, S, W, s_aux)
    lowp = dtype != torch.float32
    assert maxdiff(o, ref["o"]) < (2e-2 if lowp else 1e-4)
    for got, name in ((qd.grad, "dq"), (kd.grad, "dk"), (vd.grad, "dv")):
        assert excess(got, ref[name], 5e-2, 5e-2) <= 1.0 if lowp else maxdiff(got, ref[name]) < 2e-4, name
    assert maxdiff(sd.grad, ds_ref) < (1e-2 if lowp else 2e-3) * max(1.0, float(ds_ref.abs().max()))
    # leak test: changing one sequence must not change any other sequence's rows (bit-exact)
    k2 = kd.detach().clone()
    k2[:, :, lens[0]:lens[0] + lens[1]] += 1.0
    o2 = sa.sink_flash_attention_varlen(qd.detach(), k2, vd.detach(), cu, S, W, sd.detach())
    keep = torch.ones(N, dtype=torch.bool)
    keep[lens[0]:lens[0] + lens[1]] = False
    assert torch.equal(o2[:, :, keep], o.detach()[:, :, keep])


def test_packed_batch_rows_from_position_ids():
    """B = 2 rows packed differently (position_ids restarting mid-row), head_dim 64: tcgen05 forward + fused backward."""
    B, N, Hq, Hkv, D, W = 2, 256, 8, 1, 64, 64
    lens = [[100, 156], [256]]
    g = torch.Generator().manual_seed(4)
    mk = lambda H: torch.randn(B, H, N, D, generator=g).to(torch.bfloat16)
    q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
    pos = torch.stack([torch.cat([torch.arange(L) for L in ls]) for ls in lens]).cuda()
    qd, kd, vd = (t.cuda().requires_grad_(True) for t in (q, k, v))
    o = sa.sink_flash_attention_varlen(qd, kd, vd, None, 0, W, None, position_ids=pos)
    o.backward(do.cuda())
    assert _lib.last_impl() == "tcgen05-fused"
    for b in range(B):
        ref, _ = _oracle_packed(*(t[b:b + 1].float() for t in (q, k, v, do)), lens[b], 0, W, None)
        assert maxdiff(o[b:b + 1], ref["o"]) < 2e-2
        for got, name in ((qd.grad, "dq"), (kd.grad, "dk"), (vd.grad, "dv")):
            assert excess(got[b:b + 1], ref[name], 5e-2, 5e-2) <= 1.0


@pytest.mark.parametrize("dtype,Hq,Hkv,D,S,W,Nq,Nkv,impls", [
    (torch.bfloat16, 16, 2, 64, 0, 128, 256, 384, ("tcgen05", "tcgen05-fused")),    # halo-style: 128 extra keys in front
    (torch.bfloat16, 8, 1, 64, 0, 100, 200, 1000, ("tcgen05", "tcgen05-fused")),     # q_off = 800 (multiple of 16)
    (torch.bfloat16, 8, 1, 64, 0, 128, 77, 200, ("tcgen05", "simt")),               # q_off = 123: not a tile multiple
    (torch.bfloat16, 8, 2, 64, 4, 64, 64, 300, ("tcgen05", "simt")),                # sink tokens at the context start
    (torch.float32, 4, 2, 32, 2, 16, 10, 50, ("simt", "simt")),
    (torch.bfloat16, 8, 8, 128, 0, 4096, 128, 512, ("simt", "simt")),
])
def test_chunked_prefill_matches_full_context(dtype, Hq, Hkv, D, S, W, Nq, Nkv, impls):
    B = 2
    g = torch.Generator().manual_seed(Nq + Nkv)
    q = torch.randn(B, Hq, Nkv, D, generator=g).to(dtype)               # full-context queries; the chunk is its tail
    k = torch.randn(B, Hkv, Nkv, D, generator=g).to(dtype)
    v = torch.randn(B, Hkv, Nkv, D, generator=g).to(dtype)
    do = torch.randn(B, Hq, Nkv, D, generator=g).to(dtype)
    off = Nkv - Nq
    do[:, :, :off] = 0                                                  # only the chunk's rows send gradient
    s_aux = torch.randn(Hq, generator=g) * 0.5
    qd = q[:, :, off:].cuda().requires_grad_(True)
    kd, vd = k.cuda().requires_grad_(True), v.cuda().requires_grad_(True)
    sd = s_aux.cuda().requires_grad_(True)
    o = sa.sink_flash_attention_chunk(qd, kd, vd, S, W, sd)
    fwd_impl = _lib.last_impl()
    o.backward(do[:, :, off:].cuda())
    assert (fwd_impl, _lib.last_impl()) == impls
    o_ref, _ = orc.sink_attention_fwd(q.float(), k.float(), v.float(), S, W, s_aux)
    dq_r, dk_r, dv_r, ds_r = orc.sink_attention_bwd(q.float(), k.float(), v.float(), do.float(), S, W, s_aux)
    lowp = dtype != torch.float32
    assert maxdiff(o, o_ref[:, :, off:]) < (2e-2 if lowp else 1e-4)
    for got, ref in ((qd.grad, dq_r[:, :, off:]), (kd.grad, dk_r), (vd.grad, dv_r)):
        assert (excess(got, ref, 5e-2, 5e-2) <= 1.0) if lowp else (maxdiff(got, ref) < 2e-4)
    assert maxdiff(sd.grad, ds_r) < (1e-2 if lowp else 2e-3) * max(1.0, float(ds_r.abs().max()))
    # keys that no chunk row attends have exactly zero gradient
    if S == 0 and off - W + 1 > 0:
        assert float(kd.grad[:, :, : off - W + 1].abs().max()) == 0.0 and float(vd.grad[:, :, : off - W + 1].abs().max()) == 0.0


def test_verl_patch_keeps_s_aux_for_packed_batches():
    """The hook end to end on a padding-free batch: cu_seq_lens in, sink kernels with s_aux out (the reference would
    have called stock FA2 without s_aux here)."""
    from sink_attention import verl_patch
    N, Hq, Hkv, D, W = 96, 8, 2, 64, 16
    lens = [40, 56]
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, N, Hq, D, generator=g).to("cuda", torch.bfloat16)
    k = torch.randn(1, N, Hkv, D, generator=g).to("cuda", torch.bfloat16)
    v = torch.randn(1, N, Hkv, D, generator=g).to("cuda", torch.bfloat16)
    s_aux = (torch.randn(Hq, generator=g) * 0.5).cuda()
    cu = torch.tensor([0, 40, 96], dtype=torch.int32, device="cuda")
    out = verl_patch._sink_flash_attention_forward(q, k, v, None, N, sliding_window=W, s_aux=s_aux, cu_seq_lens_q=cu,
                                                   cu_seq_lens_k=cu, max_length_q=56, max_length_k=56)
    assert out.shape == (1, N, Hq, D) and out.is_contiguous()
    zero = torch.zeros(1, Hq, N, D)
    ref, _ = _oracle_packed(q.cpu().transpose(1, 2).float(), k.cpu().transpose(1, 2).float(), v.cpu().transpose(1, 2).float(),
                            zero, lens, 0, W, s_aux.cpu())
    assert maxdiff(out.transpose(1, 2), ref["o"]) < 2e-2


# ------------------------------------------------------------------------------------------------
# decode with per-batch cache lengths and a paged KV cache (the reference shares ONE length
# across the batch and keeps the cache contiguous, cache.py:11-13): every batch row must equal the oracle
# (decode_kernel.py:205-226 restated, tests/test_decode_kernel.py:19-55) over its own keys
# ------------------------------------------------------------------------------------------------
def _decode_rows_oracle(q, k_rows, v_rows, s_aux):
    """q [B,Hq,1,D] cpu; k_rows / v_rows: per batch row [Hkv, len_b, D] cpu."""
    outs = []
    for b in range(q.shape[0]):
        if k_rows[b].shape[1] == 0:
            outs.append(torch.zeros_like(q[b:b + 1], dtype=torch.float64))     # nothing (but the s_aux sink) to attend
        else:
            outs.append(orc.decode_attention(q[b:b + 1], k_rows[b][None], v_rows[b][None], s_aux))
    return torch.cat(outs, dim=0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Hq,Hkv,D,Nmax,lens,use_aux", [
    (4, 16, 2, 64, 700, [700, 1, 333, 64], True),
    (3, 8, 8, 128, 300, [17, 300, 0], True),            # a row with no key: only the s_aux sink
    (2, 8, 2, 64, 96, [96, 33], False),
    (64, 64, 8, 64, 1500, None, True),                   # many units: one CTA per unit
])
def test_decode_per_batch_lengths(dtype, B, Hq, Hkv, D, Nmax, lens, use_aux):
    g = torch.Generator().manual_seed(B + Nmax)
    if lens is None:
        lens = torch.randint(1, Nmax + 1, (B,), generator=g).tolist()
    q = torch.randn(B, Hq, 1, D, generator=g).to(dtype)
    k = torch.randn(B, Hkv, Nmax, D, generator=g).to(dtype)
    v = torch.randn(B, Hkv, Nmax, D, generator=g).to(dtype)
    s_aux = torch.randn(Hq, generator=g) * 0.5 if use_aux else None
    sl = torch.tensor(lens, dtype=torch.int32, device="cuda")
    o = sa.sink_decode_attention_varlen(q.cuda(), k.cuda(), v.cuda(), sl, s_aux.cuda() if use_aux else None)
    assert _lib.last_impl() == "mma"
    ref = _decode_rows_oracle(q, [k[b, :, :lens[b]] for b in range(B)], [v[b, :, :lens[b]] for b in range(B)], s_aux)
    assert maxdiff(o, ref) < 1e-2
    assert maxdiff(o, orc.decode_attention_paged(q, k, v, None, lens, s_aux)) < 1e-2
    # equal lengths == the reference entry point, bit for bit (same kernel body, same key ranges)
    full = torch.full((B,), Nmax, dtype=torch.int32, device="cuda")
    o_full = sa.sink_decode_attention_varlen(q.cuda(), k.cuda(), v.cuda(), full, s_aux.cuda() if use_aux else None)
    o_plain = sa.sink_decode_attention(q.cuda(), k.cuda(), v.cuda(), s_aux.cuda() if use_aux else None)
    assert torch.equal(o_full, o_plain)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Hq,Hkv,D,page,max_pages,lens,use_aux", [
    (4, 16, 2, 64, 32, 12, [384, 1, 200, 33], True),
    (3, 8, 4, 128, 64, 5, [320, 65, 0], True),
    (2, 32, 8, 64, 256, 3, [700, 256], False),
    (48, 64, 8, 64, 128, 9, None, True),
])
def test_decode_paged_kv(dtype, B, Hq, Hkv, D, page, max_pages, lens, use_aux):
    g = torch.Generator().manual_seed(B * page + max_pages)
    cap = page * max_pages
    if lens is None:
        lens = torch.randint(1, cap + 1, (B,), generator=g).tolist()
    num_pages = B * max_pages + 3
    q = torch.randn(B, Hq, 1, D, generator=g).to(dtype)
    k_cache = torch.randn(num_pages, page, Hkv, D, generator=g).to(dtype)       # the vLLM layout
    v_cache = torch.randn(num_pages, page, Hkv, D, generator=g).to(dtype)
    # every logical page gets its own physical page, in a shuffled order; unused table entries point at page 0
    perm = torch.randperm(num_pages, generator=g)[:B * max_pages].view(B, max_pages).to(torch.int32)
    for b in range(B):
        perm[b, (lens[b] + page - 1) // page:] = 0
    s_aux = torch.randn(Hq, generator=g) * 0.5 if use_aux else None
    sl = torch.tensor(lens, dtype=torch.int32, device="cuda")
    o = sa.sink_decode_attention_paged(q.cuda(), k_cache.cuda(), v_cache.cuda(), perm.cuda(), sl,
                                       s_aux.cuda() if use_aux else None, max_len=max(lens))
    assert _lib.last_impl() == "mma"

    def rows(cache):
        out = []
        for b in range(B):
            pages = [cache[int(perm[b, j])] for j in range((lens[b] + page - 1) // page)]       # [page, Hkv, D] each
            flat = torch.cat(pages, dim=0)[:lens[b]] if pages else cache.new_zeros(0, Hkv, D)
            out.append(flat.transpose(0, 1))                                                      # [Hkv, len, D]
        return out
    ref = _decode_rows_oracle(q, rows(k_cache), rows(v_cache), s_aux)
    assert maxdiff(o, ref) < 1e-2
    assert maxdiff(o, orc.decode_attention_paged(q, k_cache, v_cache, perm, lens, s_aux)) < 1e-2     # the oracle's own paging
    # the default planning length (the table's capacity) gives the same rows
    o2 = sa.sink_decode_attention_paged(q.cuda(), k_cache.cuda(), v_cache.cuda(), perm.cuda(), sl,
                                        s_aux.cuda() if use_aux else None)
    assert maxdiff(o2, ref) < 1e-2


def test_decode_paged_rejects_bad_page_size():
    q = torch.randn(1, 8, 1, 64, device="cuda", dtype=torch.bfloat16)
    cache = torch.randn(4, 48, 2, 64, device="cuda", dtype=torch.bfloat16)      # 48: not a power of two
    bt = torch.zeros(1, 2, dtype=torch.int32, device="cuda")
    sl = torch.tensor([50], dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        sa.sink_decode_attention_paged(q, cache, cache, bt, sl)
