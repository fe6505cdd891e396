#!/usr/bin/env python
"""C3 decode timing (graph replay, L2 flushed)."""
import os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa
dev = "cuda"; dt = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (B, Hq, Hkv, Nkv, D) in ((64, 64, 8, 4100, 64), (64, 32, 8, 4100, 128), (8, 64, 8, 4100, 64), (1, 64, 8, 16384, 64)):
    q = torch.randn(B, Hq, 1, D, device=dev, generator=g).to(dt)
    k = torch.randn(B, Hkv, Nkv, D, device=dev, generator=g).to(dt)
    v = torch.randn(B, Hkv, Nkv, D, device=dev, generator=g).to(dt)
    s_aux = torch.randn(Hq, device=dev, generator=g)
    fn = lambda: sa.sink_decode_attention(q, k, v, s_aux)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): keep = fn()
    ts = []
    for it in range(15):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    us = statistics.median(ts)
    byt = 2 * B * Hkv * Nkv * D * 2 + 2 * B * Hq * D * 2
    print(f"B={B} Hq={Hq} Hkv={Hkv} Nkv={Nkv} D={D}: {us:.1f} us, {byt / us / 1e3:.0f} GB/s ({byt / us / 1e3 / 6539.9 * 100:.1f} % of measured HBM peak)")


def _time(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): keep = fn()
    ts = []
    for it in range(15):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


# paged KV / per-batch lengths at the C3 shape (sfa_decode_paged): pages of 128 keys in a shuffled pool
B, Hq, Hkv, D, page = 64, 64, 8, 64, 128
for lens_kind in ("all 4100", "uniform 1..4100"):
    max_pages = (4100 + page - 1) // page
    num_pages = B * max_pages
    q = torch.randn(B, Hq, 1, D, device=dev, generator=g).to(dt)
    kc = torch.randn(num_pages, page, Hkv, D, device=dev, generator=g).to(dt)
    vc = torch.randn(num_pages, page, Hkv, D, device=dev, generator=g).to(dt)
    bt = torch.randperm(num_pages, device=dev, generator=g).view(B, max_pages).to(torch.int32)
    s_aux = torch.randn(Hq, device=dev, generator=g)
    lens = torch.full((B,), 4100, dtype=torch.int32, device=dev) if lens_kind == "all 4100" else \
        torch.randint(1, 4101, (B,), device=dev, generator=g).to(torch.int32)
    us = _time(lambda: sa.sink_decode_attention_paged(q, kc, vc, bt, lens, s_aux, max_len=4100))
    byt = 2 * int(lens.sum()) * Hkv * D * 2 + 2 * B * Hq * D * 2
    print(f"paged (page {page}) B={B} lens {lens_kind}: {us:.1f} us, {byt / us / 1e3:.0f} GB/s of attended K/V "
          f"({byt / us / 1e3 / 6539.9 * 100:.1f} % of measured HBM peak)")
k = torch.randn(B, Hkv, 4100, D, device=dev, generator=g).to(dt)
v = torch.randn(B, Hkv, 4100, D, device=dev, generator=g).to(dt)
lens = torch.randint(1, 4101, (B,), device=dev, generator=g).to(torch.int32)
us = _time(lambda: sa.sink_decode_attention_varlen(q, k, v, lens, s_aux))
byt = 2 * int(lens.sum()) * Hkv * D * 2 + 2 * B * Hq * D * 2
print(f"contiguous cache, per-batch lengths uniform 1..4100: {us:.1f} us, {byt / us / 1e3:.0f} GB/s of attended K/V")
