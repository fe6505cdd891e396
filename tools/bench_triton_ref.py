#!/usr/bin/env python
"""The real competitor on the same B200: the reference's own Triton kernels (JIT-compiled for sm_100a by the Triton in
this image) against libsinkfa, same inputs, same timing method.

  python tools/bench_triton_ref.py [--out gpurun_out/triton_ref_b200.json] [--configs c1,c2,c3]

The reference is the UNMODIFIED package installed by tools/install_reference.sh into baseline/_ref (git-ignored; it
travels to the GPU box with the snapshot).  It is loaded under the module name `ref_sink_attention` so that it cannot
shadow this repo's `sink_attention`; nothing in the product imports it.

Timing (both arms): a CUDA graph of INNER x (256 MiB L2 flush, call) minus a graph of INNER x (flush), / INNER, CUDA
events around the replay on the launching stream, median of 7 replays after 2 warm-up replays -- the launch overhead of
either host path is excluded, the kernels (and, for the reference, its torch glue: delta, GQA group sum, ds_aux, the
decode phase 2) are what is measured.  Parity: bf16 max-abs difference between the two implementations' outputs and
gradients (reference anchors: sink_flash_attention.py:491-689, decode_kernel.py:120-226).
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sink-flash-attention-kernel_b200"))
sys.path.insert(0, ROOT)
import sink_attention as sa  # noqa: E402
from sink_attention import _lib  # noqa: E402
from bench import attended_pairs, load_peaks  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1], [2], [3]
    "c1": dict(kind="prefill", B=1, N=8192, Hq=64, Hkv=8, D=64, S=0, W=128, s_aux=True),
    "c2": dict(kind="prefill", B=4, N=16384, Hq=32, Hkv=8, D=128, S=4, W=4096, s_aux=False),
    # gpt-oss full-attention layer (every second layer of the model): same shape, window = N
    "c1full": dict(kind="prefill", B=1, N=8192, Hq=64, Hkv=8, D=64, S=0, W=8192, s_aux=True),
    "c3": dict(kind="decode", B=64, Nkv=4100, Hq=64, Hkv=8, D=64, s_aux=True),
}


def load_reference():
    pkg = os.path.join(ROOT, "baseline", "_ref", "sink_attention")
    if not os.path.isdir(pkg):
        raise SystemExit(f"{pkg} missing: run tools/install_reference.sh in the build container first")
    spec = importlib.util.spec_from_file_location("ref_sink_attention", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_sink_attention"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "triton_ref_b200.json"))
    ap.add_argument("--configs", default="c1,c2,c3")
    ap.add_argument("--inner", type=int, default=5)
    args = ap.parse_args()
    ref = load_reference()
    import triton
    dev = torch.device("cuda", 0)
    hbm_peak, tf_burst, _, peak_src = load_peaks()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    INNER = args.inner
    flush_ms = {}

    def replay_ms(gr, reps):
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gr.replay()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    def capture(body):
        gr = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(gr):
            for it in range(INNER):
                flush_buf.fill_(it)
                keep.append(body())
        return gr, keep

    def graph_timed(fn, reps=7):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if not flush_ms:
            g0, _ = capture(lambda: None)
            replay_ms(g0, 2)
            flush_ms["ms"] = replay_ms(g0, reps)
        try:
            gr, keep = capture(fn)
        except Exception as e:      # noqa: BLE001 -- not capturable: fall back to per-call events (noted in the output)
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps + 2):
                flush_buf.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            return statistics.median(ts[2:]), f"eager events ({type(e).__name__})"
        replay_ms(gr, 2)
        ms = (replay_ms(gr, reps) - flush_ms["ms"]) / INNER
        del keep
        return ms, "graph"

    def maxabs(a, b):
        return (a.float() - b.float()).abs().max().item()

    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "triton": triton.__version__,
           "peaks": {"hbm_gbs": hbm_peak, "bf16_tflops_burst": tf_burst, "source": peak_src},
           "timing": f"CUDA graph of {INNER} x (256 MiB L2 flush, call) minus flush-only graph; CUDA events; median of 7",
           "configs": {}}
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                          "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    out["clocks_before"] = clk
    dt = torch.bfloat16
    for name in args.configs.split(","):
        c = CONFIGS[name]
        g = torch.Generator(device=dev).manual_seed(42)
        res = {"config": c}
        if c["kind"] == "prefill":
            B, N, Hq, Hkv, D, S, W = (c[k] for k in ("B", "N", "Hq", "Hkv", "D", "S", "W"))
            q = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)
            k = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt)
            v = torch.randn(B, Hkv, N, D, device=dev, generator=g).to(dt)
            do = torch.randn(B, Hq, N, D, device=dev, generator=g).to(dt)
            s_aux = (torch.randn(Hq, device=dev, generator=g) * 0.5) if c["s_aux"] else None
            pairs = attended_pairs(N, S, W) * B * Hq
            f_fwd, f_all = 4 * D * pairs, 14 * D * pairs

            def run(fn_attn, with_bwd):
                def body():
                    if not with_bwd:
                        with torch.no_grad():
                            return fn_attn(q, k, v, S, W, s_aux)
                    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
                    ss = s_aux.detach().requires_grad_(True) if s_aux is not None else None
                    o = fn_attn(qq, kk, vv, S, W, ss)
                    o.backward(do)
                    return o.detach(), qq.grad, kk.grad, vv.grad, (ss.grad if ss is not None else None)
                return body

            for arm, fn_attn in (("ours", sa.sink_flash_attention), ("triton_ref", ref.sink_flash_attention)):
                r = {}
                ms, how = graph_timed(run(fn_attn, False))
                r["fwd_ms"], r["fwd_tflops"], r["fwd_timing"] = ms, f_fwd / (ms * 1e-3) / 1e12, how
                ms, how = graph_timed(run(fn_attn, True))
                r["fwd_bwd_ms"], r["fwd_bwd_tflops"], r["fwd_bwd_timing"] = ms, f_all / (ms * 1e-3) / 1e12, how
                r["fwd_bwd_frac_of_bf16_burst_peak"] = r["fwd_bwd_tflops"] / tf_burst
                res[arm] = r
                torch.cuda.synchronize()
            res["ours"]["impl"] = _lib.last_impl()
            a = run(sa.sink_flash_attention, True)()
            b = run(ref.sink_flash_attention, True)()
            torch.cuda.synchronize()
            res["bf16_max_abs_ours_vs_triton"] = {n: maxabs(x, y) for n, x, y in zip(("o", "dq", "dk", "dv", "ds_aux"), a, b)
                                                  if x is not None}
            res["speedup_fwd"] = res["triton_ref"]["fwd_ms"] / res["ours"]["fwd_ms"]
            res["speedup_fwd_bwd"] = res["triton_ref"]["fwd_bwd_ms"] / res["ours"]["fwd_bwd_ms"]
            del q, k, v, do, a, b
        else:
            B, Nkv, Hq, Hkv, D = (c[k] for k in ("B", "Nkv", "Hq", "Hkv", "D"))
            q = torch.randn(B, Hq, 1, D, device=dev, generator=g).to(dt)
            k = torch.randn(B, Hkv, Nkv, D, device=dev, generator=g).to(dt)
            v = torch.randn(B, Hkv, Nkv, D, device=dev, generator=g).to(dt)
            s_aux = torch.randn(Hq, device=dev, generator=g) * 0.5
            nbytes = 2 * B * Hkv * Nkv * D * 2 + 2 * B * Hq * D * 2
            for arm, fn in (("ours", sa.sink_decode_attention), ("triton_ref", ref.sink_decode_attention)):
                ms, how = graph_timed(lambda fn=fn: fn(q, k, v, s_aux))
                res[arm] = {"ms": ms, "hbm_GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                            "timing": how}
            res["bf16_max_abs_ours_vs_triton"] = {"o": maxabs(sa.sink_decode_attention(q, k, v, s_aux),
                                                              ref.sink_decode_attention(q, k, v, s_aux))}
            res["speedup"] = res["triton_ref"]["ms"] / res["ours"]["ms"]
            del q, k, v
        out["configs"][name] = res
        torch.cuda.empty_cache()
        print(name, json.dumps(res, default=str), flush=True)
    out["clocks_after"] = subprocess.run(
        ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
        capture_output=True, text=True).stdout.strip()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1, default=str)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
