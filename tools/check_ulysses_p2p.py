#!/usr/bin/env python
"""Parity of the sharded (Ulysses, peer-memory) path on real GPUs -- run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/check_ulysses_p2p.py [--n-local 2048] [--rounds 8]

Checks (reference anchors: verl_patch.py:132-154 for the s_aux slice rule, the un-sharded operator
sink_flash_attention.py:491-689 for everything else):
  1. peer-memory path (routed O and routed dQ, the default) vs the NCCL all-to-all path on the same fresh inputs,
     `rounds` rounds, every call run twice, one rank delayed by a sleep kernel on alternating rounds: O, dQ, dK, dV
     bit-identical, ds_aux <= 1e-5, reruns bit-identical;
  2. against the UN-SHARDED operator on one GQA group (its q heads + kv head gathered over the ranks): O bit-identical,
     gradients within the kernel tolerance (the fused backward's CTA partition differs with the head count);
  3. forward, forward, backward, backward through ONE layer instance (two buffer sets, _P2PBuffers.pending).
`verify()` is imported by bench.py (--gpus N) so that the driver-run scaling bench checks the path it times.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))


def verify(dev, n_local=2048, rounds=4, Hq=64, Hkv=8, D=64, S=0, W=128, skew=True):
    import sink_attention as sa
    rank, world = dist.get_rank(), dist.get_world_size()
    dt = torch.bfloat16
    B = 1
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    mk = lambda H: torch.randn(B, n_local, H, D, device=dev, generator=g).to(dt)
    s_aux = torch.randn(Hq, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 0.5
    nccl = sa.UlyssesSinkAttention(S, W, None)
    p2p = sa.UlyssesSinkAttention(S, W, None, p2p=True)

    def run(mod, q, k, v, do):
        qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
        ss = s_aux.clone().requires_grad_(True)
        o = mod(qq, kk, vv, ss)
        o.backward(do)
        return o.detach(), qq.grad, kk.grad, vv.grad, ss.grad

    names = ("o", "dq", "dk", "dv", "ds_aux")
    worst = {n: 0.0 for n in names}
    rerun = {n: 0.0 for n in names}
    for rnd in range(rounds):
        q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
        if skew and (rnd % world) == rank:
            torch.cuda._sleep(int(2e6) * (1 + rnd % 3))       # ~1-3 ms: this rank enters the step late
        ref = run(nccl, q, k, v, do)
        got = run(p2p, q, k, v, do)
        if skew and ((rnd + 1) % world) == rank:
            torch.cuda._sleep(int(3e6))
        got2 = run(p2p, q, k, v, do)
        torch.cuda.synchronize()
        for n, a, b, a2 in zip(names, got, ref, got2):
            worst[n] = max(worst[n], (a.float() - b.float()).abs().max().item())
            rerun[n] = max(rerun[n], (a.float() - a2.float()).abs().max().item())
    routed = {"o": bool(p2p._bufs[0].route_o), "dq": bool(p2p._bufs[0].route_dq)}
    ok_nccl = all(worst[n] == 0.0 for n in names[:4]) and worst["ds_aux"] < 1e-5
    ok_rerun = all(rerun[n] == 0.0 for n in names)

    # ---- 2. un-sharded operator on the GQA group of kv head 0 (owned by rank 0 after the exchange)
    grp = Hq // Hkv
    def gather(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return torch.cat(out, dim=1)
    qf, kf, vf, dof = gather(q[:, :, :grp]), gather(k[:, :, :1]), gather(v[:, :, :1]), gather(do[:, :, :grp])
    qq, kk, vv = (t.transpose(1, 2).detach().requires_grad_(True) for t in (qf, kf, vf))
    ss = s_aux[:grp].clone().requires_grad_(True)
    o_u = sa.sink_flash_attention(qq, kk, vv, S, W, ss)
    o_u.backward(dof.transpose(1, 2))
    sl = slice(rank * n_local, (rank + 1) * n_local)
    un = {"o": (got[0][:, :, :grp].float() - o_u.transpose(1, 2)[:, sl].float()).abs().max().item()}
    for n, a, b in (("dq", got[1][:, :, :grp], qq.grad.transpose(1, 2)[:, sl]), ("dk", got[2][:, :, :1], kk.grad.transpose(1, 2)[:, sl]),
                    ("dv", got[3][:, :, :1], vv.grad.transpose(1, 2)[:, sl])):
        un[n] = ((a.float() - b.float()).abs() / (2e-2 + 1e-2 * b.float().abs())).max().item()     # <= 1: within tolerance
    ok_un = un["o"] == 0.0 and max(un["dq"], un["dk"], un["dv"]) <= 1.0

    # ---- 3. fwd, fwd, bwd, bwd through one layer instance
    ins = []
    for _ in range(2):
        q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
        ins.append((q, k, v, do, run(nccl, q, k, v, do)))
    leaves, outs = [], []
    for q, k, v, do, _ in ins:
        l = [t.clone().requires_grad_(True) for t in (q, k, v)] + [s_aux.clone().requires_grad_(True)]
        leaves.append(l)
        outs.append(p2p(*l))
    for o, (_, _, _, do, _) in zip(outs, ins):
        o.backward(do)
    torch.cuda.synchronize()
    ffbb = 0.0
    for l, o, (_, _, _, _, ref) in zip(leaves, outs, ins):
        for a, b in zip((o.detach(), l[0].grad, l[1].grad, l[2].grad), ref[:4]):
            ffbb = max(ffbb, (a.float() - b.float()).abs().max().item())
    ok_ffbb = ffbb == 0.0
    # ---- 4. halo-exchange mode (narrow window, no sink tokens): peer-memory path vs the portable point-to-point path
    # (same kernels on the same data: bit-identical) and vs the un-sharded operator on the gathered sequence
    halo_res = None
    if S == 0 and W - 1 <= n_local:
        hp = sa.HaloSinkAttention(W, None, p2p=True)
        hn = sa.HaloSinkAttention(W, None, p2p=False)
        hw = {n: 0.0 for n in names}
        for rnd in range(max(2, rounds // 2)):
            q, k, v, do = mk(Hq), mk(Hkv), mk(Hkv), mk(Hq)
            if skew and (rnd % world) == rank:
                torch.cuda._sleep(int(2e6))
            a = run(hp, q, k, v, do)
            b = run(hn, q, k, v, do)
            torch.cuda.synchronize()
            for n, x, y in zip(names, a, b):
                hw[n] = max(hw[n], (x.float() - y.float()).abs().max().item())
        qf, kf, vf, dof = gather(q), gather(k), gather(v), gather(do)
        qq, kk, vv = (t.transpose(1, 2).detach().requires_grad_(True) for t in (qf, kf, vf))
        ss = s_aux.clone().requires_grad_(True)
        o_u = sa.sink_flash_attention(qq, kk, vv, 0, W, ss)
        o_u.backward(dof.transpose(1, 2))
        hu = {"o": (a[0].float() - o_u.transpose(1, 2)[:, sl].float()).abs().max().item()}
        for n, x, y in (("dq", a[1], qq.grad.transpose(1, 2)[:, sl]), ("dk", a[2], kk.grad.transpose(1, 2)[:, sl]),
                        ("dv", a[3], vv.grad.transpose(1, 2)[:, sl])):
            hu[n] = ((x.float() - y.float()).abs() / (2e-2 + 1e-2 * y.float().abs())).max().item()
        ds_sum = a[4].clone()
        dist.all_reduce(ds_sum)
        hu["ds_aux_rel"] = ((ds_sum - ss.grad).abs().max() / ss.grad.abs().max().clamp_min(1e-6)).item()
        ok_halo = all(hw[n] == 0.0 for n in names) and hu["o"] < 2e-2 and max(hu["dq"], hu["dk"], hu["dv"]) <= 1.0 and \
            hu["ds_aux_rel"] < 1e-3
        halo_res = {"ok": bool(ok_halo), "max_abs_p2p_vs_portable": hw, "vs_unsharded": hu}
    ok_halo_all = halo_res is None or halo_res["ok"]
    res = {"ok": bool(ok_nccl and ok_rerun and ok_un and ok_ffbb and ok_halo_all), "halo": halo_res, "rounds": rounds, "n_local": n_local, "world": world,
           "routed": routed, "max_abs_p2p_vs_nccl": worst, "max_abs_rerun": rerun, "vs_unsharded_group0": un,
           "fwd_fwd_bwd_bwd_max_abs": ffbb, "buffer_sets": len(p2p._bufs)}
    flag = torch.tensor([1.0 if res["ok"] else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok_all_ranks"] = bool(flag.item() == 1.0)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-local", type=int, default=2048)
    ap.add_argument("--rounds", type=int, default=8)
    args = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = verify(dev, args.n_local, args.rounds)
    print(f"[rank {rank}] " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok_all_ranks"] else 1)


if __name__ == "__main__":
    main()
