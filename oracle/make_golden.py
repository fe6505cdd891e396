"""Pin oracle/sink_oracle.py against the real reference and emit golden fixtures.

Runs ONLY in the build container (needs /root/reference, which never travels to the GPU
box).  For every seeded case it

  1. runs the reference's eager oracles (tests/test_sink_attention.py:15,
     tests/test_s_aux.py:16, tests/test_decode_kernel.py:19) incl. autograd backward,
  2. runs the reference's Triton kernels under TRITON_INTERPRET=1 on the CPU
     (sink_flash_attention.py:491-689, decode_kernel.py:120-226) -- fp32 / fp16 only,
     power-of-two D only,
  3. asserts oracle/sink_oracle.py agrees with both, and
  4. stores inputs + reference outputs (fp32) in tests/golden/*.npz.

Usage:  TRITON_INTERPRET=1 python oracle/make_golden.py
"""
import os
import sys

os.environ.setdefault("TRITON_INTERPRET", "1")
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import numpy as np
import torch

import sink_oracle as orc

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def _import_ref():
    import importlib.util

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    # the eager oracles live in test modules that import the (Triton) package at top level
    import sink_attention as ref_pkg  # noqa: F401  (reference package, Triton interpreter)
    t_sa = load("ref_test_sink_attention", os.path.join(REF, "tests", "test_sink_attention.py"))
    t_aux = load("ref_test_s_aux", os.path.join(REF, "tests", "test_s_aux.py"))
    t_dec = load("ref_test_decode", os.path.join(REF, "tests", "test_decode_kernel.py"))
    return ref_pkg, t_sa, t_aux, t_dec


def maxdiff(a, b):
    return (a.double() - b.double()).abs().max().item()


from golden_cases import PREFILL_CASES, DECODE_CASES, prefill_inputs, decode_inputs, checksum


def main():
    ref_pkg, t_sa, t_aux, t_dec = _import_ref()
    report = []
    for case in PREFILL_CASES:
        (name, B, Hq, Hkv, N, D, S, W, use_aux, run_triton, store_grads) = case
        q, k, v, do, s_aux = prefill_inputs(case)

        # --- reference eager (+ autograd) ---
        qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
        sr = s_aux.clone().requires_grad_(True) if use_aux else None
        o_ref = t_aux.reference_attention_with_s_aux(qr, kr, vr, s_aux=sr, window_size=W, num_sink=S)
        o_ref.backward(do)
        if not use_aux and W >= 1:
            o_naive = t_sa.naive_sink_attention(q, k, v, S, W)
            assert maxdiff(o_naive, o_ref) < 1e-5, name
        # --- oracle ---
        o, lse = orc.sink_attention_fwd(q, k, v, S, W, s_aux)
        dq, dk, dv, dsa = orc.sink_attention_bwd(q, k, v, do, S, W, s_aux)
        e = {
            "o": maxdiff(o, o_ref), "dq": maxdiff(dq, qr.grad), "dk": maxdiff(dk, kr.grad),
            "dv": maxdiff(dv, vr.grad),
        }
        if use_aux:
            e["ds_aux"] = maxdiff(dsa, sr.grad)
        # eager restatement in the oracle == reference eager, op for op
        o_eager = orc.eager_sink_attention(q, k, v, S, W, s_aux)
        e["eager"] = maxdiff(o_eager, o_ref)
        assert max(e.values()) < 2e-5, (name, e)
        # --- reference Triton kernels in the interpreter (fp32 in -> fp32 math on CPU) ---
        lse_tri = None
        if run_triton:
            qt, kt, vt = (t.clone().requires_grad_(True) for t in (q, k, v))
            st = s_aux.clone().requires_grad_(True) if use_aux else None
            from sink_attention.sink_flash_attention import SinkFlashAttentionFunc
            o_tri = SinkFlashAttentionFunc.apply(qt, kt, vt, S, W, st)
            # LSE saved by the kernel for backward (sink_flash_attention.py:556)
            lse_tri = o_tri.grad_fn.saved_tensors[4].detach().clone()
            o_tri.backward(do)
            e["tri_o"] = maxdiff(o, o_tri)
            e["tri_dq"] = maxdiff(dq, qt.grad)
            e["tri_dk"] = maxdiff(dk, kt.grad)
            e["tri_dv"] = maxdiff(dv, vt.grad)
            if use_aux:
                e["tri_ds_aux"] = maxdiff(dsa, st.grad)
            e["tri_lse"] = maxdiff(lse, lse_tri)
            assert max(e.values()) < 5e-5, (name, e)
        report.append((name, e))
        z = np.zeros(0, np.float32)
        np.savez_compressed(
            os.path.join(OUT, f"prefill_{name}.npz"),
            in_checksum=checksum(q, k, v, do, s_aux),
            o=o_ref.detach().numpy(),
            lse=lse.float().numpy(),
            dq=(qr.grad.numpy() if store_grads else z), dk=(kr.grad.numpy() if store_grads else z),
            dv=(vr.grad.numpy() if store_grads else z),
            ds_aux=(sr.grad.numpy() if use_aux else z),
            grad_checksum=checksum(qr.grad, kr.grad, vr.grad),
            lse_triton=(lse_tri.numpy() if lse_tri is not None else z),
            src=np.array("o/dq/dk/dv/ds_aux: reference eager tests/test_s_aux.py:16-72 + autograd; "
                         "lse: oracle restatement of sink_flash_attention.py:192"
                         + ("; all cross-checked vs the reference Triton kernels in the interpreter"
                            if run_triton else "")),
        )

    for case in DECODE_CASES:
        (name, B, Hq, Hkv, Nkv, D, use_aux, run_triton) = case
        q, k, v, s_aux = decode_inputs(case)
        o_ref = t_dec.reference_decode_attention(q, k, v, s_aux)
        o = orc.decode_attention(q, k, v, s_aux)
        e = {"o": maxdiff(o, o_ref)}
        if run_triton:
            o_tri = ref_pkg.sink_decode_attention(q, k, v, s_aux)
            e["tri_o"] = maxdiff(o, o_tri)
        assert max(e.values()) < 2e-5, (name, e)
        if use_aux == "big":
            assert o_ref.abs().max().item() < 0.01
        report.append((name, e))
        np.savez_compressed(
            os.path.join(OUT, f"decode_{name}.npz"),
            in_checksum=checksum(q, k, v, s_aux),
            o=o_ref.numpy(),
            src=np.array("reference eager tests/test_decode_kernel.py:19-55"
                         + ("; Triton interpreter cross-checked" if run_triton else "")),
        )

    # cache known-answers: drive the reference SinkCacheLayer with token-id valued K and
    # record which ids it returns (cache.py:149-216); tests replay against ours + the model.
    from sink_attention.cache import SinkCacheLayer
    cache_cases = []
    for (S, W, n_prefill, n_decode) in [(4, 8, 6, 10), (4, 8, 20, 13), (2, 4, 2, 9), (0, 5, 3, 12), (4, 8, 3, 4), (1, 1, 5, 3)]:
        layer = SinkCacheLayer(S, W)
        ids = torch.arange(n_prefill, dtype=torch.float32).view(1, 1, -1, 1)
        layer.update(ids.clone(), ids.clone())
        trace = []
        model = orc.RingCacheModel(S, W)
        model.prefill(n_prefill)
        for t in range(n_decode):
            tok = torch.full((1, 1, 1, 1), float(n_prefill + t))
            k_out, _ = layer.update(tok.clone(), tok.clone())
            got = [int(x) for x in k_out.flatten().tolist()]
            model.decode()
            assert got == model.linear(), (S, W, n_prefill, t, got, model.linear())
            trace.append(got)
        cache_cases.append({"S": S, "W": W, "n_prefill": n_prefill, "trace": trace})
    import json
    with open(os.path.join(OUT, "cache_traces.json"), "w") as f:
        json.dump({"src": "reference SinkCacheLayer cache.py:29-238 driven with token-id keys",
                   "cases": cache_cases}, f)

    for name, e in report:
        print(name, {kk: f"{vv:.2e}" for kk, vv in e.items()})
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
