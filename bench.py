#!/usr/bin/env python
"""bench.py -- headline benchmark of the sink-attention hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one fwd+bwd pass of `sink_flash_attention` over one synthetic batch.
  N = 1 : BASELINE.json configs[1] -- gpt-oss-20b attention layer, B=1 N=8192 Hq=64/Hkv=8 D=64
          window=128 s_aux, bf16.  metric = masked-FLOP TFLOPS (14*D*pairs*B*Hq per step).
  N > 1 : the same layer under Ulysses sequence parallelism over NVLink (torchrun, one rank per
          GPU): every rank holds an 8192-token chunk of a (8192*N)-token sequence in HF layout,
          all-to-all -> full sequence for Hq/N heads -> attention -> all-to-all back (and the
          mirror image in backward).  Per-GPU work is fixed -> "scaling": "weak".
One JSON line is printed by rank 0.  Extra objects: roofline (dominant kernel, measured live with
CUDA events), cpu_baseline (the reference's eager masked-softmax path restated in oracle/, timed
on the host cores on a bounded sample), e2e (same step through the public API from pinned host
buffers, H2D/D2H inside the timed region), decode (configs[3] HBM GB/s), clocks.
`--impl reference` times the reference's own CPU implementation of the path (the oracle port of
its eager path: the reference is Python/Triton, nothing to compile into oracle/_ref).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))

C1 = dict(B=1, N=8192, Hq=64, Hkv=8, D=64, S=0, W=128)            # BASELINE.json configs[1]
C2 = dict(B=4, N=16384, Hq=32, Hkv=8, D=128, S=4, W=4096)          # BASELINE.json configs[2]
C3 = dict(B=64, Nkv=4100, Hq=64, Hkv=8, D=64)                      # BASELINE.json configs[3]
C4_N_TOTAL = 131072                                                # BASELINE.json configs[4]: Ulysses SP, N=131072 in total
METRIC = "fwd+bwd masked-FLOP TFLOPS (gpt-oss-20b attention layer, bf16)"
UNIT = "TFLOP/s"


def attended_pairs(n, s, w):
    w, s = max(w, 0), max(s, 0)
    # sum_i [min(i+1, W) + min(S, max(0, i-W+1))] in closed form
    full = max(0, n - w)
    win = (min(n, w) * (min(n, w) + 1)) // 2 + full * w
    sink = 0
    if s > 0 and n > w:
        m = n - w                                   # rows i = w .. n-1 see min(s, i-w+1) sinks
        t = min(s, m)
        sink = t * (t + 1) // 2 + (m - t) * s
    return win + sink


def masked_flops(cfg, n=None):
    n = cfg["N"] if n is None else n
    return 14 * cfg["D"] * attended_pairs(n, cfg["S"], cfg["W"]) * cfg["B"] * cfg["Hq"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"          # B200_PROFILING.md stated fallback


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's eager masked-softmax path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_eager_sample(n_sample, heads, steps, warmup, threads):
    """fwd+bwd of the eager path (oracle port of tests/test_s_aux.py:16-72) on a bounded sample of
    the C1 workload: the first `n_sample` positions of `heads` q heads (eager materialises N x N)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sink_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(42)
    hkv = max(1, heads // (C1["Hq"] // C1["Hkv"]))
    q = torch.randn(1, heads, n_sample, C1["D"], generator=g, requires_grad=True)
    k = torch.randn(1, hkv, n_sample, C1["D"], generator=g, requires_grad=True)
    v = torch.randn(1, hkv, n_sample, C1["D"], generator=g, requires_grad=True)
    s_aux = (torch.randn(heads, generator=g) * 0.5).requires_grad_(True)
    do = torch.randn(1, heads, n_sample, C1["D"], generator=g)
    times = []
    for it in range(warmup + steps):
        for t in (q, k, v, s_aux):
            t.grad = None
        t0 = time.perf_counter()
        o = orc.eager_sink_attention(q, k, v, C1["S"], C1["W"], s_aux)
        o.backward(do)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    flops = 14 * C1["D"] * attended_pairs(n_sample, C1["S"], C1["W"]) * heads
    return flops, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample, heads = 2048, 16
    flops, times = cpu_eager_sample(n_sample, heads, args.steps, max(args.warmup, 1), threads)
    ms = 1e3 * sum(times) / len(times)
    val = flops / (ms * 1e-3) / 1e12
    sample = (f"first {n_sample} positions x {heads} q heads of the C1 layer (eager path materialises N x N scores; "
              f"fp32, torch {threads} threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gpt-oss-20b attention layer fwd+bwd (BASELINE configs[1]): B=1 N=8192 Hq=64 Hkv=8 D=64 "
                               "window=128 s_aux", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import sink_attention as sa
    from sink_attention import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    real_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout: keep fd 1 clean for the one JSON line
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, tf_burst, tf_sust, peak_src = load_peaks()
    cfg = C1
    B, N, Hq, Hkv, D, S, W = (cfg[k] for k in ("B", "N", "Hq", "Hkv", "D", "S", "W"))
    dt = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(rank + 1)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            flush_l2()
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return ts

    flush_only = {}

    def _replay_ms(gr, reps):
        ts = []
        for _ in range(reps):
            a, b = ev(), ev()
            a.record()
            gr.replay()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    def _capture(body, inner):
        gr = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(gr):
            for it in range(inner):
                flush_buf.fill_(it)
                keep.append(body())
        return gr, keep

    def graph_timed(fn, reps=7, inner=10):
        """ms of one `fn`: a graph of `inner` x (L2 flush, fn) minus a graph of `inner` x (L2 flush), / inner --
        the CUDA event clock of this box ticks every ~2 us, too coarse for one 40-100 us kernel, and an eager
        Python call between two events measures the host launch path, not the kernel."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if inner not in flush_only:
            g0, _ = _capture(lambda: None, inner)
            _replay_ms(g0, 2)
            flush_only[inner] = _replay_ms(g0, reps)
        gr, keep = _capture(fn, inner)
        _replay_ms(gr, 2)
        ms_ = (_replay_ms(gr, reps) - flush_only[inner]) / inner
        del keep
        return ms_

    def e2e_pipelined(host_in, host_out, run_step, steps, warm=2):
        """End-to-end steps through the public API from pinned HOST buffers, software-pipelined like a training loop
        with input prefetch: step i's H2D copy (copy stream), step i-1's kernels (main stream) and step i-2's D2H
        copy of O, dQ, dK, dV, ds_aux (second copy stream) overlap; EVERY step still moves all of its inputs in and
        all of its results out.  Wall clock over `steps` steps between full device synchronisations -> ms per step."""
        s_in, s_out, main = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
        slots = [[torch.empty_like(h, device=dev).requires_grad_(h.is_floating_point()) for h in host_in] for _ in range(2)]
        done = [None, None]

        def one(i):
            sl = i % 2
            with torch.cuda.stream(s_in), torch.no_grad():
                if done[sl] is not None:
                    s_in.wait_event(done[sl])                # the kernels of step i-2 have read this slot
                for d_, h_ in zip(slots[sl], host_in):
                    d_.copy_(h_, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(s_in)
            main.wait_event(ready)
            for d_ in slots[sl]:
                d_.grad = None
            results = run_step(slots[sl])
            done[sl] = torch.cuda.Event()
            done[sl].record(main)
            with torch.cuda.stream(s_out), torch.no_grad():
                s_out.wait_event(done[sl])
                for r_, h_ in zip(results, host_out):
                    h_.copy_(r_, non_blocking=True)
                    r_.record_stream(s_out)

        for i in range(warm):
            one(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(warm, warm + steps):
            one(i)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps * 1e3

    s_aux = (torch.randn(Hq, device=dev, generator=g) * 0.5).requires_grad_(True)
    step_graph = None
    if world == 1:
        q = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt).requires_grad_(True)
        k = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt).requires_grad_(True)
        v = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt).requires_grad_(True)
        do = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)

        def eager_step():
            for t in (q, k, v, s_aux):
                t.grad = None
            o = sa.sink_flash_attention(q, k, v, S, W, s_aux)
            o.backward(do)

        # The timed step is a CUDA-graph replay of exactly the launches the autograd Function makes (sfa_fwd,
        # then sfa_bwd: the fused delta/dQ/dK/dV kernel and its fix-up, which also reduces ds_aux): the three kernels take
        # ~0.17 ms, less than the Python/ctypes launch path around them, so the eager number measures the host.
        qd_, kd_, vd_, sd_ = q.detach(), k.detach(), v.detach(), s_aux.detach()

        def c_abi_step():
            o_, lse_ = _lib.fwd(qd_, kd_, vd_, S, W, sd_)
            return _lib.bwd(qd_, kd_, vd_, o_, do, lse_, S, W, sd_)

        for _ in range(3):
            c_abi_step()
        torch.cuda.synchronize()
        step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph):
            graph_out = c_abi_step()

        def step():
            step_graph.replay()
        n_total = N
        # fwd; bwd = fused delta/dQ/dK/dV kernel + its fix-up (which also reduces ds_aux); SFA_FUSED_DELTA=0 puts the
        # separate delta/ds_aux preprocess pass back (wide windows / sinks: preprocess + ds_aux reduce + dQ kernel + dK/dV kernel)
        launches_per_step = 4 if os.environ.get("SFA_FUSED_DELTA", "1") == "0" else 3
        workload = ("gpt-oss-20b attention layer fwd+bwd (BASELINE configs[1]): B=1 N=8192 Hq=64 Hkv=8 D=64 window=128 "
                    "s_aux bf16")
        parallelism = "single GPU"
    else:
        assert Hkv % world == 0, "Ulysses needs the GPU count to divide H_kv=8"
        # BASELINE configs[4]: ONE 131072-token sequence sharded over the ranks (SFA_BENCH_NTOTAL overrides for development)
        n_total = int(os.environ.get("SFA_BENCH_NTOTAL", C4_N_TOTAL))
        N = n_total // world
        # ---- parity of the sharded path BEFORE timing it (tools/check_ulysses_p2p.py): peer-memory path vs NCCL path
        # (bit-exact, multi-round, skewed ranks), vs the un-sharded operator on one GQA group, fwd-fwd-bwd-bwd
        parity_check = None
        if os.environ.get("SFA_BENCH_SKIP_PARITY") is None and os.environ.get("SFA_BENCH_ULY", "p2p") != "nccl":
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            try:
                import check_ulysses_p2p
                parity_check = check_ulysses_p2p.verify(dev, n_local=2048, rounds=4)
            except Exception as e:      # noqa: BLE001
                parity_check = {"ok": False, "ok_all_ranks": False, "error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
            dist.barrier()
        q = torch.randn(B, N, Hq, D, device=dev, generator=g).to(dt).requires_grad_(True)       # HF layout chunk
        k = torch.randn(B, N, Hkv, D, device=dev, generator=g).to(dt).requires_grad_(True)
        v = torch.randn(B, N, Hkv, D, device=dev, generator=g).to(dt).requires_grad_(True)
        do = torch.randn(B, N, Hq, D, device=dev, generator=g).to(dt)
        # The exchange runs as peer-memory scatter kernels of libsinkfa over NVLink (sp_utils._UlyssesP2PAttention:
        # no NCCL call, no pack / unpack copies); SFA_BENCH_ULY=nccl selects the NCCL all-to-all path instead.
        want_p2p = os.environ.get("SFA_BENCH_ULY", "p2p") != "nccl"
        uly = sa.UlyssesSinkAttention(num_sink=S, window_size=W, sp_group=None, head_chunks=1, p2p=want_p2p)

        def eager_uly_step():
            for t in (q, k, v, s_aux):
                t.grad = None
            o = uly(q, k, v, s_aux)
            o.backward(do)

        if want_p2p:
            ok = torch.ones(1, device=dev)
            try:
                eager_uly_step()
                torch.cuda.synchronize()
            except Exception as e:      # noqa: BLE001  -- symmetric memory unavailable on this box
                print(f"[bench] rank {rank}: peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL",
                      file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                want_p2p = False
                uly = sa.UlyssesSinkAttention(num_sink=S, window_size=W, sp_group=None, head_chunks=1)
        step = eager_uly_step
        exch = "peer-memory scatter kernels over NVLink" if want_p2p else "NCCL all-to-all"
        step_mode = f"eager autograd step, {exch} each side"
        # The eager step is bound by the host (~25 launches around 0.2 ms of attention kernels).  The peer-memory
        # step is plain kernels on one stream, so the whole fwd+bwd is captured in ONE CUDA graph that every rank
        # replays (the NCCL path is captured only on request: its capture hung once on a 2-GPU box).
        if want_p2p or os.environ.get("SFA_BENCH_ULY_GRAPH") == "1":
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        eager_uly_step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                dist.barrier()
                uly_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(uly_graph):
                    eager_uly_step()
                torch.cuda.synchronize()
                step = uly_graph.replay
                step_graph = uly_graph
                step_mode = f"CUDA-graph replay of the autograd Ulysses step (attention kernels + {exch} each side)"
            except Exception as e:      # noqa: BLE001
                print(f"[bench] rank {rank}: graph capture of the Ulysses step failed ({type(e).__name__}: {e}); timing the eager step",
                      file=sys.stderr)
                step = eager_uly_step
        # ---- halo-exchange sequence parallelism beside it (narrow window, no sink tokens: only the W - 1 keys in front
        # of each chunk cross NVLink, sp_utils.HaloSinkAttention) -- same 131072-token sequence, same chunks
        halo_line = None
        if os.environ.get("SFA_BENCH_SKIP_HALO") is None and S == 0 and W - 1 <= N:
            try:
                hmod = sa.HaloSinkAttention(W, None, p2p=True)

                def halo_step():
                    for t in (q, k, v, s_aux):
                        t.grad = None
                    hmod(q, k, v, s_aux).backward(do)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        halo_step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                dist.barrier()
                hgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(hgraph):
                    halo_step()
                torch.cuda.synchronize()
                for _ in range(3):
                    hgraph.replay()
                torch.cuda.synchronize()
                ht = []
                for _ in range(args.steps):
                    flush_l2()
                    dist.barrier()
                    a_, b_ = ev(), ev()
                    a_.record()
                    hgraph.replay()
                    b_.record()
                    b_.synchronize()
                    ht.append(a_.elapsed_time(b_))
                hms = torch.tensor([sum(ht) / len(ht)], device=dev)
                dist.all_reduce(hms, op=dist.ReduceOp.MAX)
                halo_rows = hmod._bufs[0].halo
                halo_line = {"ms_per_step": hms.item(),
                             "value": 14 * D * attended_pairs(n_total, S, W) * B * Hq / (hms.item() * 1e-3) / 1e12, "unit": UNIT,
                             "halo_rows": halo_rows,
                             "nvlink_bytes_per_rank_per_step": 2 * 2 * B * halo_rows * Hkv * D * 2,
                             "how": "CUDA-graph replay of HaloSinkAttention(p2p=True) fwd+bwd: K/V halo rows stored into the next "
                                    "rank's symmetric buffer, attention kernels with q_off = halo, halo dK/dV stored back; max over ranks"}
                del hgraph
            except Exception as e:      # noqa: BLE001
                halo_line = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
            dist.barrier()
        # fwd: 3 scatter + barrier + attention + scatter + barrier (+ clone); bwd: scatter + barrier +
        # fused + fix-up + 3 scatter + barrier (+ 3 copies)  ->  12 kernels of libsinkfa per step (p2p path; no preprocess
        # launch since the fused backward computes delta itself)
        launches_per_step = 12 if want_p2p else 1 + 2
        workload = (f"gpt-oss-20b attention layer fwd+bwd under Ulysses SP (BASELINE configs[4]): ONE {n_total}-token sequence, "
                    f"{N}-token chunk per rank, Hq=64 Hkv=8 D=64 window=128 s_aux bf16, {exch} each side")
        parallelism = f"ulysses_sp{world}"

    # ---- the timed region: warm-up, barrier + sync, K steps (per-step CUDA events, L2 flushed between), sync
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_step = []
    for _ in range(args.steps):
        flush_l2()
        if world > 1:
            dist.barrier()
        a, b = ev(), ev()
        a.record()
        step()
        b.record()
        b.synchronize()
        per_step.append(a.elapsed_time(b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(per_step) / len(per_step)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    # masked FLOPs of the whole job: every rank's Hq/world heads over the full n_total sequence
    job_flops = 14 * D * attended_pairs(n_total, S, W) * B * Hq
    value = job_flops / (ms * 1e-3) / 1e12

    line = None
    if rank == 0:
        fwd_impl = bwd_impl = None
        roof = None
        decode = None
        e2e = None
        cpu_base = None
        if world == 1:
            # ---- per-kernel times: each stage captured in its own CUDA graph, replayed with the L2 flushed,
            # CUDA events around the replay (on the launching stream) -> dominant kernel roofline
            qd, kd, vd = q.detach(), k.detach(), v.detach()
            sd = s_aux.detach()

            t_eager = timed(eager_step, 5, 3)
            stage_ms = {}
            stage_ms["fwd"] = graph_timed(lambda: sa.sink_flash_attention_with_lse(qd, kd, vd, S, W, sd))
            fwd_impl = _lib.last_impl()
            o_s, lse_s = sa.sink_flash_attention_with_lse(qd, kd, vd, S, W, sd)
            # which backward does the step run?  narrow window + no sink tokens + head_dim 64 -> ONE kernel for
            # dQ, dK and dV (bwdf_sm100.cu) + a small fix-up; otherwise the dQ / dK/dV kernel pair
            _lib.bwd(qd, kd, vd, o_s, do, lse_s, S, W, sd)
            bwd_impl = _lib.last_impl()
            if bwd_impl == "tcgen05-fused" and os.environ.get("SFA_FUSED_DELTA", "1") != "0":
                stages = (("bwd_fused(delta,ds_aux,dq,dk,dv)", 6),)
            elif bwd_impl == "tcgen05-fused":
                stages = (("bwd_preprocess(delta,ds_aux)", 1), ("bwd_fused(dq,dk,dv)", 6))
            else:
                stages = (("bwd_preprocess(delta,ds_aux)", 1), ("bwd_dq", 2), ("bwd_dkdv", 4))
            for name, mask in stages:
                def one_stage(mask=mask):
                    _lib.load().sfa_set_bwd_stages(mask)
                    try:
                        return _lib.bwd(qd, kd, vd, o_s, do, lse_s, S, W, sd)
                    finally:
                        _lib.load().sfa_set_bwd_stages(7)
                try:
                    stage_ms[name] = graph_timed(one_stage)
                except Exception:      # noqa: BLE001  -- nothing to capture: the stage is fused into another kernel
                    torch.cuda.synchronize()
                    stage_ms[name] = 0.0
            e = 2
            bytes_alg = {
                "fwd": 2 * B * Hq * N * D * e + 2 * B * Hkv * N * D * e + 4 * B * Hq * N,
                "bwd_preprocess(delta,ds_aux)": 2 * B * Hq * N * D * e + 2 * 4 * B * Hq * N,
                "bwd_dq": 3 * B * Hq * N * D * e + 2 * B * Hkv * N * D * e + 2 * 4 * B * Hq * N,
                "bwd_dkdv": 2 * B * Hq * N * D * e + 4 * B * Hkv * N * D * e + 2 * 4 * B * Hq * N,
                # Q, dO in, dQ out; K, V in, dK, dV out; lse, delta in
                "bwd_fused(dq,dk,dv)": 3 * B * Hq * N * D * e + 4 * B * Hkv * N * D * e + 2 * 4 * B * Hq * N,
                # Q, O, dO in, dQ out; K, V in, dK, dV out; lse in (delta stays on chip / in L2)
                "bwd_fused(delta,ds_aux,dq,dk,dv)": 4 * B * Hq * N * D * e + 4 * B * Hkv * N * D * e + 4 * B * Hq * N,
            }
            pairs = attended_pairs(N, S, W) * B * Hq
            flops_alg = {"fwd": 4 * D * pairs, "bwd_preprocess(delta,ds_aux)": 2 * B * Hq * N * D,
                         "bwd_dq": 6 * D * pairs, "bwd_dkdv": 8 * D * pairs,
                         "bwd_fused(dq,dk,dv)": 10 * D * pairs, "bwd_fused(delta,ds_aux,dq,dk,dv)": 10 * D * pairs}
            dom = max(stage_ms, key=stage_ms.get)
            dur = stage_ms[dom] * 1e-3
            ach = bytes_alg[dom] / dur / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes": bytes_alg[dom], "avg_launch_ms": stage_ms[dom],
                    "tensor_tflops": flops_alg[dom] / dur / 1e12, "tensor_frac_of_burst_peak": flops_alg[dom] / dur / 1e12 / tf_burst,
                    "kernel_ms": stage_ms, "impl": {"fwd": fwd_impl, "bwd": bwd_impl}}
            traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(traffic_path):       # dram__bytes_read+write per launch from the committed ncu --set full capture
                with open(traffic_path) as f:
                    tr = json.load(f)
                roof["traffic"] = tr.get(dom)
                roof["traffic_source"] = tr.get("source")
            roof["eager_api_ms_per_step"] = sum(t_eager) / len(t_eager)
            # the same step back to back WITHOUT the flush (a step touches ~306 MB, 2.4x the L2; the flush leaves 126 MB
            # of dirty lines whose write-back runs under the next kernel's reads): reported beside the flushed value
            gr_nf = torch.cuda.CUDAGraph()
            keep_nf = []
            with torch.cuda.graph(gr_nf):
                for _ in range(10):
                    keep_nf.append(c_abi_step())
            _replay_ms(gr_nf, 2)
            ms_nf = _replay_ms(gr_nf, 7) / 10
            del keep_nf, gr_nf
            roof["whole_step"] = {
                "algorithmic_bytes": 459.3e6, "hbm_floor_ms": 459.3e6 / (hbm_peak * 1e9) * 1e3,
                "tensor_floor_ms_at_burst_peak": job_flops / (tf_burst * 1e12) * 1e3,
                "frac_of_hbm_roofline": (459.3e6 / (hbm_peak * 1e9) * 1e3) / ms,
                "frac_of_bf16_burst_peak": value / tf_burst,
                "no_flush": {"ms_per_step": ms_nf, "tflops": job_flops / (ms_nf * 1e-3) / 1e12,
                             "frac_of_hbm_roofline": (459.3e6 / (hbm_peak * 1e9) * 1e3) / ms_nf,
                             "timing": "CUDA graph of 10 back-to-back steps, no flush (306 MB touched per step, L2 = 126 MB)"},
            }
            # ---- decode (BASELINE configs[3]): HBM GB/s, each KV byte counted once
            Bd, Nkv = C3["B"], C3["Nkv"]
            qq = torch.randn(Bd, Hq, 1, D, device=dev, generator=g).to(dt)
            kk = torch.randn(Bd, Hkv, Nkv, D, device=dev, generator=g).to(dt)
            vv = torch.randn(Bd, Hkv, Nkv, D, device=dev, generator=g).to(dt)
            dms = graph_timed(lambda: sa.sink_decode_attention(qq, kk, vv, sd))
            dbytes = 2 * Bd * Hkv * Nkv * D * 2 + 2 * Bd * Hq * D * 2
            # second reading WITHOUT the flush: the 538 MB of K / V are 4x the L2, so back-to-back launches cannot reuse
            # it either -- and the kernel is not charged for HBM read / write turn-arounds against the write-back of the
            # flush buffer's dirty lines (ncu, cache control on: 86.6 us)
            gr_nf = torch.cuda.CUDAGraph()
            keep_nf = []
            with torch.cuda.graph(gr_nf):
                for _ in range(20):
                    keep_nf.append(sa.sink_decode_attention(qq, kk, vv, sd))
            _replay_ms(gr_nf, 2)
            dms_nf = _replay_ms(gr_nf, 7) / 20
            del keep_nf, gr_nf
            decode = {"workload": "KV-cache decode (BASELINE configs[3]): batch 64, sink 4 + window 4096, Hq=64 Hkv=8 D=64 "
                                  "s_aux bf16", "ms_per_step": dms, "hbm_GBps": dbytes / (dms * 1e-3) / 1e9,
                      "frac_of_hbm_peak": dbytes / (dms * 1e-3) / 1e9 / hbm_peak, "tokens_per_s": Bd / (dms * 1e-3),
                      "algorithmic_bytes": dbytes, "impl": _lib.last_impl(),
                      "timing": "CUDA graph of 10 x (L2 flush, decode) minus flush-only graph, CUDA events, median of 7 replays",
                      "no_flush": {"ms_per_step": dms_nf, "hbm_GBps": dbytes / (dms_nf * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": dbytes / (dms_nf * 1e-3) / 1e9 / hbm_peak,
                                   "timing": "CUDA graph of 20 back-to-back decode launches, no flush: inputs (538 MB) "
                                             "are larger than L2 (126 MB); CUDA events, median of 7 replays"}}
            del qq, kk, vv
            # ---- BASELINE configs[2] (Llama-style StreamingLLM layer, the tensor-bound config): fwd+bwd device time
            c2 = None
            if os.environ.get("SFA_BENCH_SKIP_C2") is None:
                B2, N2, Hq2, Hkv2, D2, S2, W2 = (C2[k_] for k_ in ("B", "N", "Hq", "Hkv", "D", "S", "W"))
                q2 = torch.randn(B2, Hq2, N2, D2, device=dev, generator=g).to(dt)
                k2 = torch.randn(B2, Hkv2, N2, D2, device=dev, generator=g).to(dt)
                v2 = torch.randn(B2, Hkv2, N2, D2, device=dev, generator=g).to(dt)
                do2 = torch.randn(B2, Hq2, N2, D2, device=dev, generator=g).to(dt)

                def c2_fwd():
                    return _lib.fwd(q2, k2, v2, S2, W2, None)
                o2, lse2 = c2_fwd()

                def c2_bwd():
                    return _lib.bwd(q2, k2, v2, o2, do2, lse2, S2, W2, None)
                t_f2 = graph_timed(c2_fwd, reps=5, inner=2)
                fwd2_impl = _lib.last_impl()
                t_b2 = graph_timed(c2_bwd, reps=5, inner=2)
                pairs2 = attended_pairs(N2, S2, W2) * B2 * Hq2
                fl2 = 14 * D2 * pairs2
                c2 = {"workload": "Llama-3-8B-style StreamingLLM layer fwd+bwd (BASELINE configs[2]): B=4 N=16384 Hq=32 Hkv=8 "
                                  "D=128 num_sink=4 window=4096 bf16",
                      "fwd_ms": t_f2, "bwd_ms": t_b2, "ms_per_step": t_f2 + t_b2, "masked_flops_per_step": fl2,
                      "tflops": fl2 / ((t_f2 + t_b2) * 1e-3) / 1e12,
                      "fwd_tflops": 4 * D2 * pairs2 / (t_f2 * 1e-3) / 1e12, "bwd_tflops": 10 * D2 * pairs2 / (t_b2 * 1e-3) / 1e12,
                      "frac_of_bf16_burst_peak": fl2 / ((t_f2 + t_b2) * 1e-3) / 1e12 / tf_burst,
                      # a 20 ms tensor-bound step runs at the power-limited clock: the sustained cuBLAS figure is its roof
                      "frac_of_bf16_sustained_peak": fl2 / ((t_f2 + t_b2) * 1e-3) / 1e12 / tf_sust,
                      "executed_tflops": 18 * D2 * pairs2 / ((t_f2 + t_b2) * 1e-3) / 1e12,
                      "executed_note": "the dQ and dK/dV kernels each recompute S and dP: 18 D flops per pair executed for 14 D counted",
                      "bound": "tensor", "impl": {"fwd": fwd2_impl, "bwd": _lib.last_impl()},
                      "timing": "CUDA graph of 2 x (L2 flush, call) minus flush-only graph, CUDA events, median of 5 replays"}
                del q2, k2, v2, do2, o2, lse2
                torch.cuda.empty_cache()
            # ---- e2e: the public API from pinned HOST buffers, H2D + D2H inside the timed region
            hq_, hk_, hv_, hdo_ = (t.detach().cpu().pin_memory() for t in (q, k, v, do))
            hs_ = s_aux.detach().cpu().pin_memory()
            ho = torch.empty_like(hq_).pin_memory()
            hdq, hdk, hdv = torch.empty_like(hq_).pin_memory(), torch.empty_like(hk_).pin_memory(), torch.empty_like(hv_).pin_memory()
            hds = torch.empty_like(hs_).pin_memory()

            def run_api(dev_in):
                dq_, dk_, dv_, ddo_, ds_ = dev_in
                o = sa.sink_flash_attention(dq_, dk_, dv_, S, W, ds_)
                o.backward(ddo_.detach())
                return [o.detach(), dq_.grad, dk_.grad, dv_.grad, ds_.grad]
            ems = e2e_pipelined([hq_, hk_, hv_, hdo_, hs_], [ho, hdq, hdk, hdv, hds], run_api, max(5, min(args.steps, 20)))
            h2d = sum(t.numel() * t.element_size() for t in (hq_, hk_, hv_, hdo_, hs_))
            d2h = sum(t.numel() * t.element_size() for t in (ho, hdq, hdk, hdv, hds))
            e2e = {"value": job_flops / (ems * 1e-3) / 1e12, "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": ems,
                   "how": "public API (sink_flash_attention + backward) from pinned host buffers; every step copies q, k, v, "
                          "dO, s_aux in and O, dQ, dK, dV, ds_aux out; copies of neighbouring steps overlap the kernels "
                          "(double-buffered inputs, three streams); wall clock between device synchronisations"}
            # ---- CPU baseline beside it (bounded sample, ~10-30 s)
            threads = os.cpu_count() or 1
            n_s, h_s = 2048, 16
            fl, tm = cpu_eager_sample(n_s, h_s, 3, 1, threads)
            cpu_base = {"value": fl / (sum(tm) / len(tm)) / 1e12, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"first {n_s} positions x {h_s} q heads of the C1 layer, fp32 eager masked-softmax "
                                  f"fwd+autograd bwd, {len(tm)} timed passes"}
        else:
            # multi-GPU e2e: the same Ulysses step fed from pinned host chunks
            hq_, hk_, hv_, hdo_ = (t.detach().cpu().pin_memory() for t in (q, k, v, do))
            ho = torch.empty_like(hq_).pin_memory()
            e2e = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            # N = 1 is configs[1] (one 8192-token layer); N > 1 is configs[4], whose 131072 tokens are fixed as N grows
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "parallelism": parallelism, "global_tokens": n_total * B,
                       "masked_flops_per_step": job_flops,
                       "l2": "256 MiB buffer written between timed steps; per-step CUDA events summed",
                       "step": ("CUDA-graph replay of the C-ABI launches of one fwd+bwd (sfa_fwd + sfa_bwd)" if world == 1
                                else step_mode)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if roof is not None:
            line["roofline"] = roof
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if decode is not None:
            line["decode"] = decode
        if world == 1 and c2 is not None:
            line["c2"] = c2
        if world > 1:
            line["parity_check"] = parity_check
            line["halo_sp"] = halo_line
        if e2e is not None:
            line["e2e"] = e2e
    if world > 1:
        # decode at N GPUs (BASELINE configs[3] per rank: the batch shards with no collective -> aggregate GB/s)
        Bd, Nkv = C3["B"], C3["Nkv"]
        qq = torch.randn(Bd, Hq, 1, D, device=dev, generator=g).to(dt)
        kk = torch.randn(Bd, Hkv, Nkv, D, device=dev, generator=g).to(dt)
        vv = torch.randn(Bd, Hkv, Nkv, D, device=dev, generator=g).to(dt)
        sdd = s_aux.detach()
        dist.barrier()
        tdm = torch.tensor([graph_timed(lambda: sa.sink_decode_attention(qq, kk, vv, sdd))], device=dev)
        dist.all_reduce(tdm, op=dist.ReduceOp.MAX)
        dbytes = (2 * Bd * Hkv * Nkv * D * 2 + 2 * Bd * Hq * D * 2) * world
        if rank == 0:
            hbm_peak_ = load_peaks()[0]
            line["decode"] = {"workload": f"KV-cache decode (BASELINE configs[3]) on every rank: batch 64 per GPU ({Bd * world} total), "
                                          "sink 4 + window 4096, Hq=64 Hkv=8 D=64 s_aux bf16; batch-sharded, no collective",
                              "ms_per_step": tdm.item(), "hbm_GBps": dbytes / (tdm.item() * 1e-3) / 1e9,
                              "frac_of_hbm_peak": dbytes / (tdm.item() * 1e-3) / 1e9 / (hbm_peak_ * world),
                              "tokens_per_s": Bd * world / (tdm.item() * 1e-3), "algorithmic_bytes": dbytes,
                              "impl": _lib.last_impl(), "scaling": "weak"}
        del qq, kk, vv
        # e2e at N GPUs: every rank feeds its chunk from pinned host memory and reads its O chunk back
        hq_, hk_, hv_, hdo_ = (t.detach().cpu().pin_memory() for t in (q, k, v, do))
        hs_ = s_aux.detach().cpu().pin_memory()
        ho = torch.empty_like(hq_).pin_memory()
        hdq, hdk, hdv = torch.empty_like(hq_).pin_memory(), torch.empty_like(hk_).pin_memory(), torch.empty_like(hv_).pin_memory()

        def run_uly(dev_in):
            dq_, dk_, dv_, ddo_, ds_ = dev_in
            o = uly(dq_, dk_, dv_, ds_)
            o.backward(ddo_.detach())
            return [o.detach(), dq_.grad, dk_.grad, dv_.grad]
        ems_n = e2e_pipelined([hq_, hk_, hv_, hdo_, hs_], [ho, hdq, hdk, hdv], run_uly, max(5, min(args.steps, 20)))
        t = torch.tensor([ems_n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            h2d = sum(x.numel() * x.element_size() for x in (hq_, hk_, hv_, hdo_, hs_)) * world
            d2h = sum(x.numel() * x.element_size() for x in (ho, hdq, hdk, hdv)) * world
            line["e2e"] = {"value": job_flops / (t.item() * 1e-3) / 1e12, "unit": UNIT, "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": d2h, "ms_per_step": t.item()}
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
