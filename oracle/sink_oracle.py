"""CPU oracle for the sink-flash-attention hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch on the CPU, the algorithm of the reference
(RulinShao/sink-flash-attention-kernel).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the product
package (``sink_attention``) never does and raises if its CUDA library is missing.

Parity status: PINNED.  ``oracle/make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) checks every function below against

  * the reference's own eager oracles ``naive_sink_attention``
    (tests/test_sink_attention.py:15-50), ``reference_attention_with_s_aux``
    (tests/test_s_aux.py:16-72), ``reference_decode_attention``
    (tests/test_decode_kernel.py:19-55), and
  * the reference's Triton kernels run under ``TRITON_INTERPRET=1``
    (sink_flash_attention.py:93-484, decode_kernel.py:28-226) incl. their LSE,

and stores seeded input/output vectors under ``tests/golden/`` which the CPU test
suite replays against this file (tests/test_oracle.py).

Conventions (all cited lines are in /root/reference/sink_attention/):
  q [B,H_q,N,D]; k,v [B,H_kv,N,D]; kv head of q head h is h // (H_q // H_kv)
  (sink_flash_attention.py:119-121); scale = 1/sqrt(D) (:505); s_aux [H_q] fp32.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


# ---------------------------------------------------------------------------
# mask
# ---------------------------------------------------------------------------
def attended_mask(n: int, num_sink: int, window_size: int, n_kv: Optional[int] = None) -> torch.Tensor:
    """Boolean [N, N_kv] attended-set predicate.

    valid(i, j) = (j <= i) & (j < num_sink | j >= i - window_size + 1)
    -- sink_flash_attention.py:30-39 (== tests/test_sink_attention.py:35-41).
    """
    n_kv = n if n_kv is None else n_kv
    i = torch.arange(n).unsqueeze(1)
    j = torch.arange(n_kv).unsqueeze(0)
    return (j <= i) & ((j < num_sink) | (j >= i - window_size + 1))


def attended_pairs(n: int, num_sink: int, window_size: int) -> int:
    """Number of attended (i, j) pairs per (batch, q-head): the masked-FLOP numerator
    P(N,S,W) = sum_i [min(i+1,W) + min(S, max(0, i-W+1))]  (SURVEY.md section 8d)."""
    w = max(window_size, 0)
    s = max(num_sink, 0)
    total = 0
    # closed form would do; N <= 2^20 keeps this loop cheap and obviously right
    for i in range(n):
        total += min(i + 1, w) + min(s, max(0, i - w + 1))
    return total


# ---------------------------------------------------------------------------
# prefill / training forward
# ---------------------------------------------------------------------------
def _expand_kv(x: torch.Tensor, groups: int) -> torch.Tensor:
    return x if groups == 1 else x.repeat_interleave(groups, dim=1)


def sink_attention_fwd(
    q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
    num_sink: int, window_size: int, s_aux: Optional[torch.Tensor] = None,
    dtype: torch.dtype = torch.float64,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """O [B,H_q,N,D] and natural-log LSE [B,H_q,N] in ``dtype``.

    Semantics follow the kernel, sink_flash_attention.py:93-194: online softmax seeded
    with (m, l) = (s_aux, 1) (:139-146) so exp(s_aux) joins the denominator only;
    rows with nothing attended and no s_aux give O = 0, LSE = -inf (:183,192).
    Numerically this equals tests/test_s_aux.py:16-72 (s_aux as an extra column).
    """
    B, Hq, N, D = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    scale = 1.0 / math.sqrt(D)
    qf, kf, vf = q.to(dtype), _expand_kv(k.to(dtype), g), _expand_kv(v.to(dtype), g)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale                     # [B,Hq,N,N]
    mask = attended_mask(N, num_sink, window_size).to(s.device)
    s = s.masked_fill(~mask, float("-inf"))
    if s_aux is not None:
        col = s_aux.to(dtype).reshape(1, Hq, 1, 1).expand(B, Hq, N, 1)
        s_all = torch.cat([s, col], dim=-1)
    else:
        s_all = s
    lse = torch.logsumexp(s_all, dim=-1)                                   # -inf on empty rows
    p = torch.exp(s - lse.unsqueeze(-1))
    p = torch.nan_to_num(p, nan=0.0)                                       # empty rows -> 0
    o = torch.matmul(p, vf)
    return o, lse


def sink_attention_bwd(
    q, k, v, do, num_sink: int, window_size: int, s_aux=None, dtype=torch.float64,
):
    """Closed-form gradients (dq, dk, dv, ds_aux) in ``dtype`` -- no autograd.

    dV = P^T dO; dP = dO V^T; dS = P * (dP - delta), delta = rowsum(dO * O)
    (sink_flash_attention.py:242-251, 582); dQ = dS K * scale (:449,481);
    dK = dS^T Q * scale (:227,251); GQA group sum (:648-651);
    ds_aux[h] = -sum_{b,n} exp(s_aux[h] - lse) * delta (:653-665).
    """
    B, Hq, N, D = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    scale = 1.0 / math.sqrt(D)
    qf, dof = q.to(dtype), do.to(dtype)
    kf, vf = _expand_kv(k.to(dtype), g), _expand_kv(v.to(dtype), g)
    o, lse = sink_attention_fwd(q, k, v, num_sink, window_size, s_aux, dtype)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    mask = attended_mask(N, num_sink, window_size).to(s.device)
    p = torch.exp(s - lse.unsqueeze(-1))
    p = torch.where(mask, p, torch.zeros_like(p))
    p = torch.nan_to_num(p, nan=0.0)
    delta = (dof * o).sum(-1)                                              # [B,Hq,N]
    dv_e = torch.matmul(p.transpose(-1, -2), dof)
    dp = torch.matmul(dof, vf.transpose(-1, -2))
    ds = p * (dp - delta.unsqueeze(-1))
    dq = torch.matmul(ds, kf) * scale
    dk_e = torch.matmul(ds.transpose(-1, -2), qf) * scale
    dk = dk_e.view(B, Hkv, g, N, D).sum(2)
    dv = dv_e.view(B, Hkv, g, N, D).sum(2)
    ds_aux = None
    if s_aux is not None:
        sink_prob = torch.exp(s_aux.to(dtype)[None, :, None] - lse)
        ds_aux = -(sink_prob * delta).sum(dim=(0, 2))
    return dq, dk, dv, ds_aux


def eager_sink_attention(q, k, v, num_sink: int = 0, window_size: Optional[int] = None, s_aux=None):
    """The reference's EAGER masked-softmax path, restated op for op so that autograd
    reproduces its backward.  Follows tests/test_s_aux.py:16-72 (additive -1e9 mask,
    s_aux appended as an extra logit column, max-subtract, softmax, drop column) which
    itself mirrors HF gpt-oss eager attention.  This is what bench.py times as the CPU
    baseline ("the reference's eager CPU path").  Differentiable; dtype follows inputs.
    """
    B, Hq, N, D = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    scale = 1.0 / math.sqrt(D)
    ke, ve = _expand_kv(k, g), _expand_kv(v, g)
    w = torch.matmul(q, ke.transpose(-2, -1)) * scale
    if window_size is None:
        window_size = N
    valid = attended_mask(N, num_sink, window_size).to(q.device)
    w = w + ((~valid).to(w.dtype) * (-1e9)).unsqueeze(0).unsqueeze(0)
    if s_aux is not None:
        col = s_aux.to(w.dtype).reshape(1, Hq, 1, 1).expand(B, Hq, N, 1)
        comb = torch.cat([w, col], dim=-1)
        comb = comb - comb.max(dim=-1, keepdim=True).values
        probs = torch.softmax(comb, dim=-1)[..., :-1]
    else:
        probs = torch.softmax(w, dim=-1)
    return torch.matmul(probs, ve)


# ---------------------------------------------------------------------------
# row-sampled oracle: the same arithmetic for SELECTED query rows / key rows, O((S + W) * D) per row and no N x N
# matrix -- lets the parity tests compare the CUDA path with the oracle directly at the full BASELINE sizes
# (C1 / C2 / C4), where sink_attention_fwd / _bwd above would need terabytes.  Pinned against the full functions on
# small shapes by tests/test_oracle.py (test_sampled_oracle_matches_full_oracle).
# ---------------------------------------------------------------------------
def _row_keys(i: torch.Tensor, num_sink: int, window_size: int, n_kv: int):
    """For query positions i [R]: candidate key indices [R, C] (sinks then the window band) and their validity
    under sink_flash_attention.py:36-39 (sinks that fall inside the band are counted once, as band keys)."""
    S, W = max(num_sink, 0), max(window_size, 0)
    dev = i.device
    cand = []
    if S > 0:
        cand.append(torch.arange(S, device=dev).unsqueeze(0).expand(i.numel(), S))
    if W > 0:
        cand.append(i.unsqueeze(1) - (W - 1) + torch.arange(W, device=dev).unsqueeze(0))
    if not cand:
        z = torch.zeros(i.numel(), 1, dtype=torch.long, device=dev)
        return z, torch.zeros_like(z, dtype=torch.bool)
    j = torch.cat(cand, dim=1)
    valid = (j >= 0) & (j < n_kv) & (j <= i.unsqueeze(1))
    if S > 0:
        is_sink_col = torch.zeros_like(valid)
        is_sink_col[:, :S] = True
        lo = i.unsqueeze(1) - W + 1
        valid &= torch.where(is_sink_col, j < lo if W > 0 else torch.ones_like(valid), j >= 0)
        if W > 0:
            valid &= torch.where(is_sink_col, torch.ones_like(valid), (j >= S) | (j >= lo))   # band keys keep j >= lo
    return j.clamp(0, n_kv - 1), valid


def sampled_fwd(q, k, v, num_sink: int, window_size: int, s_aux, rows: torch.Tensor, dtype=torch.float64,
                chunk: int = 128) -> Tuple[torch.Tensor, torch.Tensor]:
    """O [R, D] and natural-log LSE [R] for the sampled rows ``rows`` = int64 [R, 3] of (batch, q head, position).
    Same definition as sink_attention_fwd (sink_flash_attention.py:93-194), evaluated row by row."""
    B, Hq, N, D = q.shape
    g = Hq // k.shape[1]
    scale = 1.0 / math.sqrt(D)
    outs, lses = [], []
    for c0 in range(0, rows.shape[0], chunk):
        r = rows[c0:c0 + chunk]
        b, h, i = r[:, 0], r[:, 1], r[:, 2]
        j, valid = _row_keys(i, num_sink, window_size, N)
        kk = k[b.unsqueeze(1), (h // g).unsqueeze(1), j].to(dtype)                 # [R, C, D]
        vv = v[b.unsqueeze(1), (h // g).unsqueeze(1), j].to(dtype)
        qq = q[b, h, i].to(dtype)                                                  # [R, D]
        sc = torch.einsum("rd,rcd->rc", qq, kk) * scale
        sc = sc.masked_fill(~valid, float("-inf"))
        if s_aux is not None:
            sc_all = torch.cat([sc, s_aux.to(dtype)[h].unsqueeze(1)], dim=1)
        else:
            sc_all = sc
        lse = torch.logsumexp(sc_all, dim=1)
        p = torch.nan_to_num(torch.exp(sc - lse.unsqueeze(1)), nan=0.0)
        outs.append(torch.einsum("rc,rcd->rd", p, vv))
        lses.append(lse)
    return torch.cat(outs), torch.cat(lses)


def sampled_dq(q, k, v, do, o, lse, num_sink: int, window_size: int, rows: torch.Tensor, dtype=torch.float64,
               chunk: int = 128) -> torch.Tensor:
    """dQ [R, D] of the sampled rows.  ``o`` and ``lse`` are the forward's outputs as the backward receives them
    (delta = rowsum(dO * O) uses the stored O, sink_flash_attention.py:582; P = exp(S - LSE), :242)."""
    B, Hq, N, D = q.shape
    g = Hq // k.shape[1]
    scale = 1.0 / math.sqrt(D)
    outs = []
    for c0 in range(0, rows.shape[0], chunk):
        r = rows[c0:c0 + chunk]
        b, h, i = r[:, 0], r[:, 1], r[:, 2]
        j, valid = _row_keys(i, num_sink, window_size, N)
        kk = k[b.unsqueeze(1), (h // g).unsqueeze(1), j].to(dtype)
        vv = v[b.unsqueeze(1), (h // g).unsqueeze(1), j].to(dtype)
        qq, dd, oo = q[b, h, i].to(dtype), do[b, h, i].to(dtype), o[b, h, i].to(dtype)
        sc = torch.einsum("rd,rcd->rc", qq, kk) * scale
        p = torch.exp(sc - lse[b, h, i].to(dtype).unsqueeze(1))
        p = torch.nan_to_num(torch.where(valid, p, torch.zeros_like(p)), nan=0.0)
        delta = (dd * oo).sum(1, keepdim=True)
        dp = torch.einsum("rd,rcd->rc", dd, vv)
        ds = p * (dp - delta)
        outs.append(torch.einsum("rc,rcd->rd", ds, kk) * scale)
    return torch.cat(outs)


def sampled_dkdv(q, k, v, do, o, lse, num_sink: int, window_size: int, keys: torch.Tensor, dtype=torch.float64,
                 max_rows: int = 8192) -> Tuple[torch.Tensor, torch.Tensor]:
    """dK, dV [R, D] for the sampled key rows ``keys`` = int64 [R, 3] of (batch, kv head, position): sums over every
    query row of the GQA group that attends the key (sink keys: all later rows; others: the W rows from the key on).
    dV = P^T dO, dK = scale * dS^T Q, GQA group sum (sink_flash_attention.py:242-251, 648-651)."""
    B, Hq, N, D = q.shape
    g = Hq // k.shape[1]
    scale = 1.0 / math.sqrt(D)
    S, W = max(num_sink, 0), max(window_size, 0)
    dks, dvs = [], []
    for b, y, j in keys.tolist():
        hi = N - 1 if j < S else min(j + W - 1, N - 1)
        dk = torch.zeros(D, dtype=dtype, device=q.device)
        dv = torch.zeros(D, dtype=dtype, device=q.device)
        kj, vj = k[b, y, j].to(dtype), v[b, y, j].to(dtype)
        if hi >= j and (j < S or W > 0):
            for i0 in range(j, hi + 1, max_rows):
                i1 = min(i0 + max_rows, hi + 1)
                hs = slice(y * g, (y + 1) * g)
                qq = q[b, hs, i0:i1].to(dtype)                                     # [g, r, D]
                dd, oo = do[b, hs, i0:i1].to(dtype), o[b, hs, i0:i1].to(dtype)
                p = torch.nan_to_num(torch.exp(torch.einsum("grd,d->gr", qq, kj) * scale - lse[b, hs, i0:i1].to(dtype)), nan=0.0)
                ds = p * (torch.einsum("grd,d->gr", dd, vj) - (dd * oo).sum(-1))
                dv += torch.einsum("gr,grd->d", p, dd)
                dk += torch.einsum("gr,grd->d", ds, qq) * scale
        dks.append(dk)
        dvs.append(dv)
    return torch.stack(dks), torch.stack(dvs)


# ---------------------------------------------------------------------------
# decode
# ---------------------------------------------------------------------------
def decode_attention(q, k, v, s_aux=None, dtype=torch.float64):
    """Single-query attention over a (sink+window) cache: q [B,H_q,1,D], k/v [B,H_kv,N_kv,D].

    All cached keys are attended; s_aux is one extra softmax column without a value
    (decode_kernel.py:205-226; tests/test_decode_kernel.py:19-55).  Returns [B,H_q,1,D].
    """
    B, Hq, _, D = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    scale = 1.0 / math.sqrt(D)
    kf, vf = _expand_kv(k.to(dtype), g), _expand_kv(v.to(dtype), g)
    s = torch.matmul(q.to(dtype), kf.transpose(-1, -2)) * scale            # [B,Hq,1,Nkv]
    if s_aux is not None:
        col = s_aux.to(dtype)[None, :, None, None].expand(B, -1, 1, 1)
        p = torch.softmax(torch.cat([col, s], dim=-1), dim=-1)[..., 1:]
    else:
        p = torch.softmax(s, dim=-1)
    return torch.matmul(p, vf)


def decode_attention_paged(q, k_cache, v_cache, block_table, seq_lens, s_aux=None, dtype=torch.float64):
    """Checker for the paged / per-batch-length decode (SURVEY section 8 row f4 -- beyond the reference, whose cache
    shares ONE length across the batch and is contiguous, cache.py:11-13): row b attends the first seq_lens[b] keys of
    the logical sequence block_table[b] spells out over the page pool [num_pages, page_size, H_kv, D]; per row it is
    exactly decode_attention (decode_kernel.py:205-226 restated).  block_table None: k_cache / v_cache are
    [B, H_kv, N_max, D].  A row without keys returns zeros (only the s_aux column is attended)."""
    B = q.shape[0]
    outs = []
    for b in range(B):
        n = int(seq_lens[b]) if seq_lens is not None else None
        if block_table is None:
            kb, vb = k_cache[b:b + 1], v_cache[b:b + 1]
            if n is not None:
                kb, vb = kb[:, :, :n], vb[:, :, :n]
        else:
            page = k_cache.shape[1]
            n = block_table.shape[1] * page if n is None else n
            npg = (n + page - 1) // page
            if npg:
                kb = torch.cat([k_cache[int(block_table[b, j])] for j in range(npg)], dim=0)[:n].transpose(0, 1)[None]
                vb = torch.cat([v_cache[int(block_table[b, j])] for j in range(npg)], dim=0)[:n].transpose(0, 1)[None]
            else:
                kb = vb = k_cache.new_zeros(1, k_cache.shape[2], 0, k_cache.shape[3])
        if kb.shape[2] == 0:
            outs.append(torch.zeros(1, q.shape[1], 1, q.shape[3], dtype=dtype))
        else:
            outs.append(decode_attention(q[b:b + 1], kb, vb, s_aux, dtype=dtype))
    return torch.cat(outs, dim=0)


# ---------------------------------------------------------------------------
# cache (sink buffer + ring window buffer) -- pure-Python model of cache.py
# ---------------------------------------------------------------------------
class RingCacheModel:
    """Index-level model of SinkCacheLayer (cache.py:29-238): which absolute token
    positions are resident, and in which order get_kv() returns them."""

    def __init__(self, num_sink: int, window_size: int):
        self.S, self.W = num_sink, window_size
        self.sink: list[int] = []
        self.ring: list[Optional[int]] = [None] * window_size
        self.window_len = 0
        self.write_pos = 0
        self.seen = 0
        self.prefilled = False

    def prefill(self, n: int):                       # cache.py:80-127
        self.seen = n
        if n <= self.S:
            self.sink = list(range(n))
            self.window_len, self.write_pos = 0, 0
        else:
            self.sink = list(range(self.S))
            non_sink = n - self.S
            if non_sink <= self.W:
                for t in range(non_sink):
                    self.ring[t] = self.S + t
                self.window_len = non_sink
                self.write_pos = non_sink % self.W
            else:
                for t in range(self.W):
                    self.ring[t] = n - self.W + t
                self.window_len, self.write_pos = self.W, 0
        self.prefilled = True

    def decode(self):                                # cache.py:129-147
        self.ring[self.write_pos] = self.seen
        self.seen += 1
        self.write_pos = (self.write_pos + 1) % self.W
        self.window_len = min(self.window_len + 1, self.W)

    def linear(self) -> list[int]:                   # cache.py:185-216
        out = list(self.sink)
        if self.window_len > 0:
            if self.window_len < self.W:
                out += self.ring[: self.window_len]
            else:
                out += self.ring[self.write_pos:] + self.ring[: self.write_pos]
        return out
