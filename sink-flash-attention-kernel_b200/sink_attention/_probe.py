"""ctypes binding of libsinkfa_probe.so (include/sinkfa_probe.h): tcgen05 / TMA / TMEM micro-probes.

Diagnostics for performance work (tools/probe_*.py) and the UMMA descriptor self-test of the GPU suite; not part of
the operator API and not linked into libsinkfa.so.
"""
from __future__ import annotations

import ctypes
import os

import torch

from ._lib import DTYPE_CODE, SinkFAError, _require_cuda, _stream

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFA_PROBE_LIB") or os.path.join(_HERE, "libsinkfa_probe.so")
EXPORTS = ("sfa_probe_last_error", "sfa_probe_umma", "sfa_probe_tma_bw", "sfa_probe_mma_rate", "sfa_probe_mma_desc",
           "sfa_probe_math_rate", "sfa_probe_tmem_rate")
_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SinkFAError(f"{LIB_PATH} not found: build it with `make -C {os.path.dirname(_HERE)}`")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    p, i = c.c_void_p, c.c_int
    lib.sfa_probe_last_error.restype = c.c_char_p
    lib.sfa_probe_math_rate.argtypes = [p, p, i, i, i, p]
    lib.sfa_probe_tmem_rate.argtypes = [p, p, i, i, i, p]
    lib.sfa_probe_mma_rate.argtypes = [p, i, i, i, i, p]
    lib.sfa_probe_mma_desc.argtypes = [p, c.POINTER(c.c_int), p]
    lib.sfa_probe_tma_bw.argtypes = [p, i, i, i, i, i, i, i, p]
    lib.sfa_probe_umma.argtypes = [p, p, p, i, i, i, i, p]
    for name in EXPORTS[1:]:
        getattr(lib, name).restype = i
    _lib = lib
    return lib


def probe_umma(a: torch.Tensor, b: torch.Tensor, n: int, k: int, mode: int) -> torch.Tensor:
    lib = load()
    _require_cuda(a, b)
    c = torch.empty((128, n), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        rc = lib.sfa_probe_umma(a.data_ptr(), b.data_ptr(), c.data_ptr(), n, k, mode, DTYPE_CODE[a.dtype], _stream(a))
    if rc != 0:
        raise SinkFAError(f"sfa_probe_umma: {lib.sfa_probe_last_error().decode(errors='replace')} (cudaError {rc})")
    return c
