#!/usr/bin/env python
"""GPU-side times of the C1 kernels with the CPU launch path taken out: each stage is captured in a CUDA graph
and replayed (L2 flushed before every replay, CUDA events around the replay)."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402
from sink_attention import _lib  # noqa: E402

B, N, Hq, Hkv, D, S, W = 1, 8192, 64, 8, 64, 0, 128
if len(sys.argv) > 1 and sys.argv[1] == "c2s":
    B, N, Hq, Hkv, D, S, W = 1, 8192, 32, 8, 128, 4, 4096
if len(sys.argv) > 1 and sys.argv[1] == "c1full":      # gpt-oss FULL-attention layer (every second layer): window = N
    B, N, Hq, Hkv, D, S, W = 1, 8192, 64, 8, 64, 0, 8192
if os.environ.get("SFA_TG_W"): W = int(os.environ["SFA_TG_W"])
if os.environ.get("SFA_TG_S"): S = int(os.environ["SFA_TG_S"])
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
dt = torch.bfloat16
q = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)
k = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt)
v = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt)
do = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)
s_aux = torch.randn(Hq, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
o, lse = sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)


INNER = 10


def _replay_us(gr, reps):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts), min(ts)


def _capture(body):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for it in range(INNER):
            flush.fill_(it)
            body()
    return gr


_flush_only = None


def graph_time(fn, reps=7):
    """(median, min) us of one `fn`: a graph of INNER x (L2 flush, fn) minus a graph of INNER x (L2 flush), / INNER.
    (CUDA event timestamps on this box tick every 2.048 us: a single ~50 us replay cannot resolve a 1 us change.)"""
    global _flush_only
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if _flush_only is None:
        g0 = _capture(lambda: None)
        _replay_us(g0, 2)
        _flush_only = _replay_us(g0, reps)
    gr = _capture(fn)
    _replay_us(gr, 2)
    med, mn = _replay_us(gr, reps)
    return round((med - _flush_only[0]) / INNER, 2), round((mn - _flush_only[1]) / INNER, 2)


def stage(mask):
    def fn():
        lib.sfa_set_bwd_stages(mask)
        _lib.bwd(q, k, v, o, do, lse, S, W, s_aux)
        lib.sfa_set_bwd_stages(7)
    return fn


print("fwd            us (median, min):", graph_time(lambda: sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)))
print("bwd preprocess us:", graph_time(stage(1)))
print("bwd dq         us:", graph_time(stage(2)))
print("bwd dkdv       us:", graph_time(stage(4)))
print("bwd dq+dkdv    us:", graph_time(stage(6)), "(fused kernel + fix-up when the shape allows)")
print("bwd all        us:", graph_time(stage(7)))


def full():
    o2, lse2 = sa.sink_flash_attention_with_lse(q, k, v, S, W, s_aux)
    _lib.bwd(q, k, v, o2, do, lse2, S, W, s_aux)


t_all = graph_time(full)
print("fwd+bwd        us:", t_all)
w_, s_ = max(W, 0), max(S, 0)
pairs = sum(min(i + 1, w_) + min(s_, max(0, i - w_ + 1)) for i in range(N)) * B * Hq
print(f"masked TFLOP/s fwd+bwd: {14 * D * pairs / (t_all[0] * 1e-6) / 1e12:.1f}  ({14 * D * pairs / (t_all[0] * 1e-6) / 1e12 / 1628.7 * 100:.1f} % of the bf16 burst peak)")
