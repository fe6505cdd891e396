#!/usr/bin/env python
"""Development check of the peer-memory Ulysses path (torchrun, one rank per GPU): p2p=True against the NCCL
all-to-all path on the same inputs (outputs and gradients must be bit-identical: same kernels, same data), then
timings of both (eager) and of a CUDA-graph replay of the p2p step."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
import sink_attention as sa  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, n, Hq, Hkv, D, S, W = 1, int(os.environ.get("N_LOCAL", 8192)), 64, 8, 64, 0, 128
g = torch.Generator(device=dev).manual_seed(1 + rank)
dt = torch.bfloat16
mk = lambda H: torch.randn(B, n, H, D, device=dev, generator=g).to(dt)
q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
s_aux = (torch.randn(Hq, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 0.5)


def run(mod):
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    ss = s_aux.clone().requires_grad_(True)
    o = mod(qq, kk, vv, ss)
    o.backward(do)
    return o.detach(), qq.grad, kk.grad, vv.grad, ss.grad


nccl = sa.UlyssesSinkAttention(S, W, None)
p2p = sa.UlyssesSinkAttention(S, W, None, p2p=True)
# fresh data every round: a peer store that is not yet visible when the barrier releases the reader shows up as
# stale rows of the previous round
worst = {}
ok = True
for rnd in range(int(os.environ.get("ROUNDS", 6))):
    q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
    ref = run(nccl)
    got = run(p2p)
    ref2 = run(nccl)
    got2 = run(p2p)
    torch.cuda.synchronize()
    for name, a, b, a2, b2 in zip(("o", "dq", "dk", "dv", "ds_aux"), got, ref, got2, ref2):
        d = (a.float() - b.float()).abs().max().item()
        worst[name] = max(worst.get(name, 0.0), d)
        worst[name + "(p2p rerun)"] = max(worst.get(name + "(p2p rerun)", 0.0), (a.float() - a2.float()).abs().max().item())
        worst[name + "(nccl rerun)"] = max(worst.get(name + "(nccl rerun)", 0.0), (b.float() - b2.float()).abs().max().item())
        ok &= (d == 0.0) if name != "ds_aux" else (d < 1e-5)
print(f"[rank {rank}] " + ("MATCH" if ok else "MISMATCH") + " over the rounds; max |p2p - nccl| = " +
      ", ".join(f"{n} {d:.3e}" for n, d in worst.items()) + "\n", end="", flush=True)


def timeit(fn, steps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
ss = s_aux.clone().requires_grad_(True)


def step(mod):
    for t in (qq, kk, vv, ss):
        t.grad = None
    mod(qq, kk, vv, ss).backward(do)


t_nccl = timeit(lambda: step(nccl))
t_p2p = timeit(lambda: step(p2p))
print(f"[rank {rank}] eager step: nccl {t_nccl:.3f} ms   p2p {t_p2p:.3f} ms", flush=True)
try:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step(p2p)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step(p2p)
    torch.cuda.synchronize()
    dist.barrier()
    t_graph = timeit(gr.replay)
    print(f"[rank {rank}] graph replay of the p2p step: {t_graph:.3f} ms", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"[rank {rank}] graph capture failed: {type(e).__name__}: {e}", flush=True)
# ---- pieces: one scatter of q (67 MB, half of it to the peer), one barrier, one clone
from sink_attention import _lib  # noqa: E402
bufs = p2p._bufs[0]


def ev_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps * 1e3


hq_l = Hq // world
t_sc = ev_time(lambda: _lib.ulysses_scatter(q, bufs.peer[0], rank, 0, bufs.tot, 0))
t_sc1 = ev_time(lambda: _lib.ulysses_scatter(bufs.do_full, bufs.peer[1], rank, 1, Hq, 0))
t_bar = ev_time(lambda: bufs.barrier())
t_cl = ev_time(lambda: bufs.o_seq.clone())
nbytes = q.numel() * 2
print(f"[rank {rank}] scatter q (mode 0, {nbytes / 1e6:.0f} MB, {(world - 1) / world:.2f} remote): {t_sc:.1f} us = "
      f"{nbytes * (world - 1) / world / t_sc / 1e3:.0f} GB/s over NVLink; mode 1: {t_sc1:.1f} us; barrier {t_bar:.1f} us; "
      f"clone {t_cl:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
