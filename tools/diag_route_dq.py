#!/usr/bin/env python
"""Diagnostic for the open routed-dQ issue (torchrun, 2 GPUs): is the fused backward itself nondeterministic when its
dQ stores go to a peer (local dK / dV differ between two calls on the same inputs), or only what arrives at the peer?"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
from sink_attention import _lib  # noqa: E402
from sink_attention.sp_utils import _P2PBuffers  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, n, Hq, Hkv, D, S, W = 1, 8192, 64, 8, 64, 0, 128
hq_l, hkv_l = Hq // world, Hkv // world
bufs = _P2PBuffers(None, B, n, Hq, Hkv, D, torch.bfloat16, dev)
g = torch.Generator(device=dev).manual_seed(3 + rank)
s32 = (torch.randn(hq_l, device=dev, generator=g) * 0.5)
res = {"local dk": 0.0, "local dv": 0.0, "received dq": 0.0, "routed vs unrouted dk": 0.0, "routed vs scattered dq": 0.0}
for rnd in range(4):
    mk = lambda H: torch.randn(B, n, H, D, device=dev, generator=g).to(torch.bfloat16)
    q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
    _lib.ulysses_scatter(q, bufs.peer[0], rank, 0, bufs.tot, 0)
    _lib.ulysses_scatter(k, bufs.peer[0], rank, 0, bufs.tot, hq_l)
    _lib.ulysses_scatter(v, bufs.peer[0], rank, 0, bufs.tot, hq_l + hkv_l)
    _lib.ulysses_scatter(do, bufs.peer[2], rank, 0, hq_l, 0)
    bufs.barrier()
    full = bufs.qkv_full
    qh, kh, vh = (full[:, :, :hq_l].transpose(1, 2), full[:, :, hq_l:hq_l + hkv_l].transpose(1, 2),
                  full[:, :, hq_l + hkv_l:].transpose(1, 2))
    do_h = bufs.do_full.transpose(1, 2)
    o, lse = _lib.fwd(qh, kh, vh, S, W, s32)
    dq_u, dk_u, dv_u, _ = _lib.bwd(qh, kh, vh, o, do_h, lse, S, W, s32)           # unrouted, local
    _lib.ulysses_scatter(dq_u.transpose(1, 2), bufs.peer[3], rank, 1, Hq + 2 * Hkv, 0)
    bufs.barrier()
    dq_recv_u = bufs.g_seq[:, :, :Hq].clone()
    torch.cuda.synchronize()
    bufs.barrier()
    snaps = []
    route = _lib.make_route(bufs.peer[3], n, Hq + 2 * Hkv, rank * hq_l)
    for rep in range(2):
        bufs.g_seq.zero_()
        torch.cuda.synchronize()
        bufs.barrier()
        _, dk_r, dv_r, _ = _lib.bwd(qh, kh, vh, o, do_h, lse, S, W, s32, dq_route=route)
        bufs.barrier()
        snaps.append((dk_r.clone(), dv_r.clone(), bufs.g_seq[:, :, :Hq].clone()))
        torch.cuda.synchronize()
        bufs.barrier()
    md = lambda a, b: (a.float() - b.float()).abs().max().item()
    res["local dk"] = max(res["local dk"], md(snaps[0][0], snaps[1][0]))
    res["local dv"] = max(res["local dv"], md(snaps[0][1], snaps[1][1]))
    res["received dq"] = max(res["received dq"], md(snaps[0][2], snaps[1][2]))
    res["routed vs unrouted dk"] = max(res["routed vs unrouted dk"], md(snaps[0][0], dk_u), md(snaps[1][0], dk_u))
    res["routed vs scattered dq"] = max(res["routed vs scattered dq"], md(snaps[0][2], dq_recv_u), md(snaps[1][2], dq_recv_u))
print(f"[rank {rank}] max differences between two routed calls on the same inputs: " +
      ", ".join(f"{k_} {v_:.3e}" for k_, v_ in res.items()) + "\n", end="", flush=True)
dist.barrier()
dist.destroy_process_group()
