"""Shared helpers for the parity tests (tests may import oracle/; the product never does)."""
import os

import numpy as np
import torch

import golden_cases as gc
import sink_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_prefill(case):
    name = case[0]
    z = np.load(os.path.join(GOLDEN, f"prefill_{name}.npz"))
    q, k, v, do, s_aux = gc.prefill_inputs(case)
    assert abs(gc.checksum(q, k, v, do, s_aux) - float(z["in_checksum"])) < 1e-6 * max(1.0, float(z["in_checksum"])), \
        "seeded inputs drifted from the ones the golden outputs were generated with"
    return (q, k, v, do, s_aux), z


def load_decode(case):
    name = case[0]
    z = np.load(os.path.join(GOLDEN, f"decode_{name}.npz"))
    q, k, v, s_aux = gc.decode_inputs(case)
    assert abs(gc.checksum(q, k, v, s_aux) - float(z["in_checksum"])) < 1e-6 * max(1.0, float(z["in_checksum"]))
    return (q, k, v, s_aux), z


def maxdiff(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    fin = torch.isfinite(a) & torch.isfinite(b)
    same_inf = (~fin) & (a == b)
    assert bool((fin | same_inf).all()), "non-finite mismatch"
    return (a[fin] - b[fin]).abs().max().item() if fin.any() else 0.0


def excess(a, b, atol, rtol):
    """max over elements of |a-b| / (atol + rtol*|b|): <= 1 means `a` is close to `b` in the sense of
    torch.testing.assert_close(atol, rtol), the form the reference's gradient tests use
    (tests/test_sink_attention.py:94-96 with atol = rtol = 5e-2)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert bool(torch.isfinite(a).all()) and bool(torch.isfinite(b).all()), "non-finite value"
    return ((a - b).abs() / (atol + rtol * b.abs())).max().item()


def to_dev(t, dtype=None, dev="cuda"):
    if t is None:
        return None
    return t.to(device=dev, dtype=dtype if dtype is not None else t.dtype)
