"""ctypes binding of libsinkfa.so (C ABI declared in include/sinkfa.h).

There is no fallback: if the CUDA library is missing or a tensor is not on a CUDA device the
call raises.  PyTorch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFA_LIB: developer override used by tools/trace_*.py to load the timeline build (`make trace`)
LIB_PATH = os.environ.get("SFA_LIB") or os.path.join(_HERE, "libsinkfa.so")

DTYPE_CODE = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}
OP_FWD, OP_BWD, OP_DECODE = 0, 1, 2
IMPL_AUTO, IMPL_SIMT = 0, 1

EXPORTS = (
    "sfa_version", "sfa_last_error", "sfa_set_impl", "sfa_set_bwd_stages", "sfa_set_trace_buffer", "sfa_last_impl", "sfa_workspace_bytes",
    "sfa_fwd", "sfa_bwd", "sfa_decode", "sfa_decode_ring", "sfa_decode_paged", "sfa_ulysses_scatter", "sfa_fwd_sp", "sfa_bwd_sp", "sfa_set_debug", "sfa_cache_append", "sfa_fwd_ex", "sfa_bwd_ex",
)

_lib = None


class SpRoute(ctypes.Structure):
    """sfa_sp_route of include/sinkfa.h: output routing of the Ulysses layout (peer receive buffers)."""
    _fields_ = [("P", ctypes.c_int), ("n_local", ctypes.c_int), ("heads_total", ctypes.c_int),
                ("head_off", ctypes.c_int), ("peer", ctypes.c_void_p * 8)]


def make_route(peer_ptrs: Sequence[int], n_local: int, heads_total: int, head_off: int) -> SpRoute:
    if len(peer_ptrs) > 8:
        raise ValueError("at most 8 ranks per route")
    r = SpRoute()
    r.P, r.n_local, r.heads_total, r.head_off = len(peer_ptrs), int(n_local), int(heads_total), int(head_off)
    for i, ptr in enumerate(peer_ptrs):
        r.peer[i] = int(ptr)
    return r


class AttnExt(ctypes.Structure):
    """sfa_attn_ext of include/sinkfa.h: packed (varlen) sequences and chunked prefill / halo keys."""
    _fields_ = [("seq_lo", ctypes.c_void_p), ("seq_hi", ctypes.c_void_p), ("seq_batch_stride", ctypes.c_int64),
                ("n_kv", ctypes.c_int), ("q_off", ctypes.c_int)]


def make_ext(n_q: int, n_kv: int, q_off: int = 0, seq_lo: Optional[torch.Tensor] = None,
             seq_hi: Optional[torch.Tensor] = None) -> Optional["AttnExt"]:
    """seq_lo [N_q] or [B, N_q], seq_hi [N_kv] or [B, N_kv]: int32 CUDA tensors (both or none)."""
    if seq_lo is None and q_off == 0 and n_kv == n_q:
        return None
    e = AttnExt()
    e.n_kv, e.q_off = int(n_kv), int(q_off)
    e.seq_lo = e.seq_hi = None
    e.seq_batch_stride = 0
    if seq_lo is not None:
        if seq_hi is None or seq_lo.dtype != torch.int32 or seq_hi.dtype != torch.int32:
            raise ValueError("seq_lo and seq_hi must both be int32 tensors")
        if not (seq_lo.is_contiguous() and seq_hi.is_contiguous()):
            raise ValueError("seq_lo / seq_hi must be contiguous")
        if seq_lo.dim() == 2:
            # one allocation per array, rows n_kv apart: the C ABI takes ONE batch stride for both
            if seq_lo.shape[1] != seq_hi.shape[1]:
                raise ValueError("batched seq_lo / seq_hi need equal row lengths (q_off == 0)")
            e.seq_batch_stride = seq_lo.shape[1] if seq_lo.shape[0] > 1 else 0
        e.seq_lo, e.seq_hi = seq_lo.data_ptr(), seq_hi.data_ptr()
    return e


class SinkFAError(RuntimeError):
    pass


def _i64(vals: Sequence[int]):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def load() -> ctypes.CDLL:
    """Load libsinkfa.so (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SinkFAError(
            f"{LIB_PATH} not found: build it with `make -C {os.path.dirname(_HERE)}` "
            "(hand-written sm_100a kernels; there is no CPU or Triton fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    p, i, i64p, f32p = c.c_void_p, c.c_int, c.POINTER(c.c_int64), c.c_void_p
    lib.sfa_version.restype = i
    lib.sfa_last_error.restype = c.c_char_p
    lib.sfa_last_impl.restype = c.c_char_p
    lib.sfa_set_impl.argtypes = [i]
    lib.sfa_set_impl.restype = i
    lib.sfa_set_bwd_stages.argtypes = [i]
    lib.sfa_set_bwd_stages.restype = i
    lib.sfa_set_trace_buffer.argtypes = [p]
    lib.sfa_set_trace_buffer.restype = i
    lib.sfa_workspace_bytes.argtypes = [i] * 7
    lib.sfa_workspace_bytes.restype = c.c_size_t
    lib.sfa_fwd.argtypes = [p, p, p, p, f32p, f32p] + [i] * 8 + [i64p] * 4 + [p, c.c_size_t, p]
    lib.sfa_fwd.restype = i
    lib.sfa_bwd.argtypes = [p] * 5 + [f32p, f32p] + [p] * 3 + [f32p] + [i] * 8 + [i64p] * 8 + [p, c.c_size_t, p]
    lib.sfa_bwd.restype = i
    lib.sfa_fwd_sp.argtypes = lib.sfa_fwd.argtypes + [c.POINTER(SpRoute)]
    lib.sfa_fwd_sp.restype = i
    lib.sfa_bwd_sp.argtypes = [p] * 5 + [f32p, f32p] + [p] * 2 + [f32p] + [i] * 8 + [i64p] * 7 + [p, c.c_size_t, p, c.POINTER(SpRoute)]
    lib.sfa_bwd_sp.restype = i
    lib.sfa_fwd_ex.argtypes = lib.sfa_fwd.argtypes + [c.POINTER(AttnExt)]
    lib.sfa_fwd_ex.restype = i
    lib.sfa_bwd_ex.argtypes = lib.sfa_bwd.argtypes + [c.POINTER(AttnExt)]
    lib.sfa_bwd_ex.restype = i
    lib.sfa_decode.argtypes = [p, p, p, p, f32p] + [i] * 6 + [i64p] * 4 + [p, c.c_size_t, p]
    lib.sfa_decode.restype = i
    lib.sfa_decode_ring.argtypes = [p] * 6 + [f32p] + [i] * 7 + [i64p] * 4 + [p, c.c_size_t, p]
    lib.sfa_decode_ring.restype = i
    lib.sfa_decode_paged.argtypes = [p, p, p, p, f32p, p, p] + [i] * 7 + [c.c_int64] + [i64p] * 4 + [p, c.c_size_t, p]
    lib.sfa_decode_paged.restype = i
    lib.sfa_ulysses_scatter.argtypes = [p, c.POINTER(c.c_void_p)] + [i] * 8 + [i64p, i, i, p]
    lib.sfa_ulysses_scatter.restype = i
    lib.sfa_cache_append.argtypes = [p] * 4 + [i] * 4 + [i64p] * 2 + [i, i, p]
    lib.sfa_cache_append.restype = i
    lib.sfa_set_debug.argtypes = [i, i]
    lib.sfa_set_debug.restype = i
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().sfa_last_error().decode(errors="replace")
        if rc < 0:
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise SinkFAError(f"{what}: {msg} (cudaError {rc})")


def _require_cuda(*ts: torch.Tensor):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise SinkFAError(
                "sink_attention runs only on CUDA tensors (sm_100a kernels; no CPU fallback); "
                f"got a tensor on {t.device}")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _unit_last(t: torch.Tensor) -> torch.Tensor:
    """Kernels accept any (batch, head, position) strides but need channel stride 1.  Broadcast views (stride 0 on
    a dim of size > 1: ``k.expand(B, ...)``, a ``dO`` that autograd expanded from a reduction) are materialised, as
    the reference's ``.contiguous()`` does (sink_flash_attention.py:507-509,581): a TMA tensor map cannot walk them."""
    if t.stride(-1) != 1 or any(st == 0 and sz > 1 for st, sz in zip(t.stride(), t.shape)):
        return t.contiguous()
    return t


def _dense_non_overlapping(t: torch.Tensor) -> bool:
    sizes_strides = sorted(zip(t.stride(), t.shape))
    expect = 1
    for st, sz in sizes_strides:
        if sz == 1:
            continue
        if st != expect:
            return False
        expect *= sz
    return True


def last_impl() -> str:
    return load().sfa_last_impl().decode()


def set_impl(code: int):
    _check(load().sfa_set_impl(int(code)), "sfa_set_impl")


def _s_aux_f32(s_aux: Optional[torch.Tensor], hq: int) -> Optional[torch.Tensor]:
    if s_aux is None:
        return None
    if tuple(s_aux.shape) != (hq,):
        raise AssertionError(f"s_aux shape must be [H_q={hq}], got {tuple(s_aux.shape)}")
    return s_aux.detach().contiguous().float()


def fwd(q, k, v, num_sink: int, window_size: int, s_aux_f32, o_route: Optional[SpRoute] = None,
        ext: Optional[AttnExt] = None):
    """-> (o, lse).  q [B,Hq,N,D] (any strides with unit channel stride), k/v [B,Hkv,N,D] ([B,Hkv,n_kv,D] with ext).
    o_route: also store O into the peers' receive buffers (sfa_fwd_sp); ValueError if the shape cannot route.
    ext: packed sequences / chunked prefill (sfa_fwd_ex)."""
    lib = load()
    _require_cuda(q, k, v, s_aux_f32)
    B, Hq, N, D = q.shape
    Hkv = k.shape[1]
    q, k, v = _unit_last(q), _unit_last(k), _unit_last(v)
    o = torch.empty_like(q)   # keeps q's memory layout: an HF [B,N,H,D] view gets an HF-layout output
    if o.stride(-1) != 1:
        o = torch.empty((B, Hq, N, D), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, Hq, N), device=q.device, dtype=torch.float32)
    args = (q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
            s_aux_f32.data_ptr() if s_aux_f32 is not None else None,
            B, Hq, Hkv, N, D, int(num_sink), int(window_size), DTYPE_CODE[q.dtype],
            _i64(q.stride()), _i64(k.stride()), _i64(v.stride()), _i64(o.stride()),
            None, 0, _stream(q))
    with torch.cuda.device(q.device):
        if ext is not None:
            rc = lib.sfa_fwd_ex(*args, ctypes.byref(ext))
        elif o_route is None:
            rc = lib.sfa_fwd(*args)
        else:
            rc = lib.sfa_fwd_sp(*args, ctypes.byref(o_route))
    _check(rc, "sfa_fwd")
    return o, lse


def bwd(q, k, v, o, do, lse, num_sink: int, window_size: int, s_aux_f32, dq_route: Optional[SpRoute] = None,
        ext: Optional[AttnExt] = None):
    """-> (dq, dk, dv, ds_aux|None); dk/dv are already reduced over the GQA group (fp32 accumulate).
    dq_route: dQ is stored ONLY into the peers' receive buffers (sfa_bwd_sp) and returned as None."""
    lib = load()
    _require_cuda(q, k, v, o, do, lse, s_aux_f32)
    B, Hq, N, D = q.shape
    Hkv = k.shape[1]
    q, k, v, o, do = (_unit_last(t) for t in (q, k, v, o, do))
    # The tensor-core kernels load Q and dO tiles with one row order, which is set by whether the head stride is
    # smaller than the position stride (HF [B,N,H,D] views) or not ([B,H,N,D]).  Only that ORDER has to agree.
    def _hf_order(t):
        return t.shape[1] > 1 and t.shape[2] > 1 and t.stride(1) < t.stride(2)
    if _hf_order(do) != _hf_order(q):
        do = do.transpose(1, 2).contiguous().transpose(1, 2) if _hf_order(q) else do.contiguous()
    if ext is not None and (ext.q_off != 0 or ext.n_kv != N):
        dk, dv = torch.zeros_like(k), torch.zeros_like(v)      # keys no query row attends keep a zero gradient
    else:
        dk, dv = torch.empty_like(k), torch.empty_like(v)
    dq = torch.empty_like(q) if dq_route is None else None
    ds_aux = torch.empty((Hq,), device=q.device, dtype=torch.float32) if s_aux_f32 is not None else None
    code = DTYPE_CODE[q.dtype]
    ws_bytes = lib.sfa_workspace_bytes(OP_BWD, B, Hq, Hkv, N, D, code)
    ws = torch.empty((ws_bytes,), device=q.device, dtype=torch.uint8)
    if dq_route is not None:
        with torch.cuda.device(q.device):
            rc = lib.sfa_bwd_sp(
                q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                s_aux_f32.data_ptr() if s_aux_f32 is not None else None,
                dk.data_ptr(), dv.data_ptr(), ds_aux.data_ptr() if ds_aux is not None else None,
                B, Hq, Hkv, N, D, int(num_sink), int(window_size), code,
                _i64(q.stride()), _i64(k.stride()), _i64(v.stride()), _i64(o.stride()), _i64(do.stride()),
                _i64(dk.stride()), _i64(dv.stride()),
                ws.data_ptr(), ws_bytes, _stream(q), ctypes.byref(dq_route))
        _check(rc, "sfa_bwd_sp")
        return None, dk, dv, ds_aux
    bargs = (q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
             s_aux_f32.data_ptr() if s_aux_f32 is not None else None,
             dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ds_aux.data_ptr() if ds_aux is not None else None,
             B, Hq, Hkv, N, D, int(num_sink), int(window_size), code,
             _i64(q.stride()), _i64(k.stride()), _i64(v.stride()), _i64(o.stride()), _i64(do.stride()),
             _i64(dq.stride()), _i64(dk.stride()), _i64(dv.stride()),
             ws.data_ptr(), ws_bytes, _stream(q))
    with torch.cuda.device(q.device):
        rc = lib.sfa_bwd(*bargs) if ext is None else lib.sfa_bwd_ex(*bargs, ctypes.byref(ext))
    _check(rc, "sfa_bwd")
    return dq, dk, dv, ds_aux


def decode(q, k, v, s_aux_f32):
    """q [B,Hq,1,D]; k,v [B,Hkv,Nkv,D] -> o [B,Hq,1,D]."""
    lib = load()
    _require_cuda(q, k, v, s_aux_f32)
    B, Hq, _, D = q.shape
    Hkv, Nkv = k.shape[1], k.shape[2]
    q, k, v = _unit_last(q), _unit_last(k), _unit_last(v)
    o = torch.empty((B, Hq, 1, D), device=q.device, dtype=q.dtype)
    code = DTYPE_CODE[q.dtype]
    with torch.cuda.device(q.device):       # the work decomposition (hence the workspace) depends on the device's SM count
        ws_bytes = lib.sfa_workspace_bytes(OP_DECODE, B, Hq, Hkv, Nkv, D, code)
    ws = torch.empty((max(ws_bytes, 1),), device=q.device, dtype=torch.uint8)
    with torch.cuda.device(q.device):
        rc = lib.sfa_decode(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
            s_aux_f32.data_ptr() if s_aux_f32 is not None else None,
            B, Hq, Hkv, Nkv, D, code,
            _i64(q.stride()[:2]), _i64(k.stride()[:3]), _i64(v.stride()[:3]), _i64(o.stride()[:2]),
            ws.data_ptr(), ws_bytes, _stream(q))
    _check(rc, "sfa_decode")
    return o


def decode_paged(q, k_cache, v_cache, s_aux_f32, seq_lens=None, block_table=None, max_len=None):
    """Decode with per-batch cache lengths and / or a paged cache (sfa_decode_paged).

    block_table is None: k_cache, v_cache [B,Hkv,Nkv,D] (any batch / head / position strides).
    block_table [B, max_pages] int32: k_cache, v_cache are page pools [num_pages, page_size, Hkv, D] (the vLLM layout;
    any page / position / head strides).  seq_lens: int32 [B] on the device or None."""
    lib = load()
    _require_cuda(q, k_cache, v_cache, s_aux_f32)
    B, Hq, _, D = q.shape
    q = _unit_last(q)
    if k_cache.stride(-1) != 1 or v_cache.stride(-1) != 1:
        raise ValueError("cache tensors must have unit channel stride")
    if block_table is not None:
        if block_table.dtype != torch.int32 or not block_table.is_cuda or block_table.dim() != 2 or block_table.stride(1) != 1:
            raise ValueError("block_table must be a CUDA int32 [B, max_pages] tensor with unit inner stride")
        page_size, Hkv = k_cache.shape[1], k_cache.shape[2]
        ks = (k_cache.stride(0), k_cache.stride(2), k_cache.stride(1))        # (page, head, position)
        vs = (v_cache.stride(0), v_cache.stride(2), v_cache.stride(1))
        cap = block_table.shape[1] * page_size
        max_len = cap if max_len is None else int(max_len)
        if max_len > cap:
            raise ValueError(f"max_len {max_len} exceeds the block table's capacity {cap}")
        plan_len = (max_len + page_size - 1) // page_size * page_size
        bt_ptr, bt_stride = block_table.data_ptr(), block_table.stride(0)
    else:
        Hkv, page_size = k_cache.shape[1], 0
        ks, vs = k_cache.stride()[:3], v_cache.stride()[:3]
        max_len = k_cache.shape[2] if max_len is None else int(max_len)
        plan_len = max_len
        bt_ptr, bt_stride = None, 0
    if seq_lens is not None and (seq_lens.dtype != torch.int32 or not seq_lens.is_cuda or seq_lens.numel() != B
                                 or not seq_lens.is_contiguous()):
        raise ValueError("seq_lens must be a contiguous CUDA int32 tensor of B entries")
    o = torch.empty((B, Hq, 1, D), device=q.device, dtype=q.dtype)
    code = DTYPE_CODE[q.dtype]
    with torch.cuda.device(q.device):
        ws_bytes = lib.sfa_workspace_bytes(OP_DECODE, B, Hq, Hkv, plan_len, D, code)
    ws = torch.empty((max(ws_bytes, 1),), device=q.device, dtype=torch.uint8)
    with torch.cuda.device(q.device):
        rc = lib.sfa_decode_paged(
            q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), o.data_ptr(),
            s_aux_f32.data_ptr() if s_aux_f32 is not None else None,
            bt_ptr, seq_lens.data_ptr() if seq_lens is not None else None,
            B, Hq, Hkv, max_len, D, code, page_size, bt_stride,
            _i64(q.stride()[:2]), _i64(ks), _i64(vs), _i64(o.stride()[:2]),
            ws.data_ptr(), ws_bytes, _stream(q))
    _check(rc, "sfa_decode_paged")
    return o


def decode_ring(q, sink_k, sink_v, win_k, win_v, sink_len: int, window_len: int, s_aux_f32):
    """Decode straight from a SinkCacheLayer's buffers (no linearisation copy)."""
    lib = load()
    _require_cuda(q, sink_k, sink_v, win_k, win_v, s_aux_f32)
    B, Hq, _, D = q.shape
    Hkv = win_k.shape[1]
    q = _unit_last(q)
    for t in (sink_k, sink_v, win_k, win_v):
        if t.stride(-1) != 1:
            raise ValueError("cache buffers must have unit channel stride")
    if sink_k.stride() != sink_v.stride() or win_k.stride() != win_v.stride():
        raise ValueError("K and V cache buffers must share strides")
    o = torch.empty((B, Hq, 1, D), device=q.device, dtype=q.dtype)
    code = DTYPE_CODE[q.dtype]
    with torch.cuda.device(q.device):
        ws_bytes = lib.sfa_workspace_bytes(OP_DECODE, B, Hq, Hkv, sink_len + window_len, D, code)
    ws = torch.empty((max(ws_bytes, 1),), device=q.device, dtype=torch.uint8)
    with torch.cuda.device(q.device):
        rc = lib.sfa_decode_ring(
            q.data_ptr(), sink_k.data_ptr(), sink_v.data_ptr(), win_k.data_ptr(), win_v.data_ptr(), o.data_ptr(),
            s_aux_f32.data_ptr() if s_aux_f32 is not None else None,
            B, Hq, Hkv, int(sink_len), int(window_len), D, code,
            _i64(q.stride()[:2]), _i64(sink_k.stride()[:3]), _i64(win_k.stride()[:3]), _i64(o.stride()[:2]),
            ws.data_ptr(), ws_bytes, _stream(q))
    _check(rc, "sfa_decode_ring")
    return o


def cache_append(k_new, v_new, win_k, win_v, write_pos: int):
    """One decoded token ([B,Hkv,1,D]) -> ring slot `write_pos` of the window buffers, K and V in one launch."""
    lib = load()
    _require_cuda(k_new, v_new, win_k, win_v)
    B, H, W, D = win_k.shape
    k_new, v_new = _unit_last(k_new), _unit_last(v_new)
    if k_new.stride()[:2] != v_new.stride()[:2]:
        v_new = v_new.contiguous()
        k_new = k_new.contiguous()
    if win_k.stride() != win_v.stride() or win_k.stride(-1) != 1:
        raise ValueError("K and V ring buffers must share strides (unit channel stride)")
    with torch.cuda.device(win_k.device):
        rc = lib.sfa_cache_append(k_new.data_ptr(), v_new.data_ptr(), win_k.data_ptr(), win_v.data_ptr(), B, H, D,
                                  DTYPE_CODE[win_k.dtype], _i64(k_new.stride()[:2]), _i64(win_k.stride()[:3]), W,
                                  int(write_pos), _stream(win_k))
    _check(rc, "sfa_cache_append")


def ulysses_scatter(src: torch.Tensor, peer_ptrs: Sequence[int], rank: int, mode: int, dst_heads: int, head_off: int):
    """src [B, L, H, D] (unit channel stride) -> rows stored into the peers' receive buffers (see include/sinkfa.h)."""
    lib = load()
    _require_cuda(src)
    src = _unit_last(src)
    B, L, H, D = src.shape
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    with torch.cuda.device(src.device):
        rc = lib.sfa_ulysses_scatter(src.data_ptr(), arr, len(peer_ptrs), int(rank), int(mode), B, L, H, D,
                                     DTYPE_CODE[src.dtype], _i64(src.stride()[:3]), int(dst_heads), int(head_off),
                                     _stream(src))
    _check(rc, "sfa_ulysses_scatter")


def set_debug(knob: int, value: int):
    """Test / diagnostics knobs of the library (include/sinkfa.h: sfa_set_debug)."""
    _check(load().sfa_set_debug(int(knob), int(value)), "sfa_set_debug")
