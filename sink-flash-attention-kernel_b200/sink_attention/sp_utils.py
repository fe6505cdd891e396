"""Sequence parallelism for sink attention.

Three schemes live here (the third, HaloSinkAttention, exchanges only the W - 1 keys a narrow window needs).

1. Ulysses (what the north-star layout asks for; the reference leaves the all-to-all to verl,
   verl_patch.py:15-20, and only slices ``s_aux``, :132-154).  Activations arrive sequence-sharded
   ``[B, N/P, H, D]``; one all-to-all turns them into head-sharded ``[B, N, H/P, D]``; every rank runs
   the unmodified attention op on the full sequence for its heads (heads are independent, dK/dV reduce
   inside a GQA group, ds_aux is per head -> no collective inside the op); the inverse all-to-all
   returns ``O``.  ``UlyssesSinkAttention`` wires that up with autograd, pipelines the exchange in
   head chunks on a side stream so NVLink traffic overlaps the tensor-core work, and takes ``s_aux``
   for ALL heads, slicing the local part with the reference's rule.

2. The reference's sequence-chunk helpers ``prepare_sink_kv_for_sp`` / ``reduce_sink_kv_grads`` /
   ``get_local_position_offset`` / ``SinkAttentionSPWrapper`` (sp_utils.py:28-180), kept with the same
   names and collective semantics (broadcast of the sink K/V from group-rank 0, all-reduce of their
   gradients).  As in the reference this scheme has no window halo; it is API surface, not the
   recommended path.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


# =============================================================================================
# Ulysses all-to-all
# =============================================================================================
def _group_size_rank(group) -> Tuple[int, int]:
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _a2a_seq_to_head(x: torch.Tensor, group) -> torch.Tensor:
    """[B, n, H, D] (this rank's sequence chunk, all heads) -> [B, n*P, H/P, D] (all positions, local heads)."""
    P, _ = _group_size_rank(group)
    if P == 1:
        return x
    B, n, H, D = x.shape
    assert H % P == 0, f"heads ({H}) must be divisible by the sequence-parallel size ({P})"
    hl = H // P
    # send[r] = heads of rank r: [P, B, n, hl, D]
    send = x.reshape(B, n, P, hl, D).permute(2, 0, 1, 3, 4).contiguous()
    recv = torch.empty_like(send)                  # recv[s] = sequence chunk s
    dist.all_to_all_single(recv, send, group=group)
    if B == 1:                                     # [P,1,n,hl,D] is already [1, P*n, hl, D] in memory
        return recv.view(1, P * n, hl, D)
    return recv.permute(1, 0, 2, 3, 4).reshape(B, P * n, hl, D)


def _a2a_head_to_seq(x: torch.Tensor, group) -> torch.Tensor:
    """[B, N, H/P, D] (all positions, local heads) -> [B, N/P, H, D] (local positions, all heads)."""
    P, _ = _group_size_rank(group)
    if P == 1:
        return x
    B, N, hl, D = x.shape
    assert N % P == 0, f"sequence length ({N}) must be divisible by the sequence-parallel size ({P})"
    n = N // P
    send = x.reshape(B, P, n, hl, D).permute(1, 0, 2, 3, 4).contiguous()   # send[s] = chunk s
    recv = torch.empty_like(send)                                         # recv[r] = heads of rank r
    dist.all_to_all_single(recv, send, group=group)
    return recv.permute(1, 2, 0, 3, 4).reshape(B, n, P * hl, D)


class _SeqToHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _a2a_seq_to_head(x, group)

    @staticmethod
    def backward(ctx, g):
        return _a2a_head_to_seq(g.contiguous(), ctx.group), None


class _HeadToSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _a2a_head_to_seq(x, group)

    @staticmethod
    def backward(ctx, g):
        return _a2a_seq_to_head(g.contiguous(), ctx.group), None


class _QKVSeqToHead(torch.autograd.Function):
    """q, k, v in ONE all-to-all each way: ``[B, n, H, D]`` chunks -> ``[B, n*P, H/P, D]`` views of one receive
    buffer (heads of this rank: its q heads, then its k heads, then its v heads).  The views have a position
    stride of (hq_l + 2*hkv_l)*D -- the kernels take arbitrary (batch, head, position) strides, so nothing is
    copied after the exchange.  Backward packs dq/dk/dv the same way and sends them home in one all-to-all."""

    @staticmethod
    def forward(ctx, q, k, v, group):
        P, _ = _group_size_rank(group)
        B, n, Hq, D = q.shape
        Hkv = k.shape[2]
        hq_l, hkv_l = Hq // P, Hkv // P
        tot = hq_l + 2 * hkv_l
        send = torch.empty(P, B, n, tot, D, device=q.device, dtype=q.dtype)
        send[:, :, :, :hq_l].copy_(q.reshape(B, n, P, hq_l, D).permute(2, 0, 1, 3, 4))
        send[:, :, :, hq_l:hq_l + hkv_l].copy_(k.reshape(B, n, P, hkv_l, D).permute(2, 0, 1, 3, 4))
        send[:, :, :, hq_l + hkv_l:].copy_(v.reshape(B, n, P, hkv_l, D).permute(2, 0, 1, 3, 4))
        recv = torch.empty_like(send)                       # recv[s] = sequence chunk s, this rank's heads
        dist.all_to_all_single(recv, send, group=group)
        full = recv.view(1, P * n, tot, D) if B == 1 else recv.permute(1, 0, 2, 3, 4).reshape(B, P * n, tot, D)
        ctx.group, ctx.dims = group, (P, B, n, Hq, Hkv, D)
        return full[:, :, :hq_l], full[:, :, hq_l:hq_l + hkv_l], full[:, :, hq_l + hkv_l:]

    @staticmethod
    def backward(ctx, dq, dk, dv):
        P, B, n, Hq, Hkv, D = ctx.dims
        hq_l, hkv_l = Hq // P, Hkv // P
        tot = hq_l + 2 * hkv_l
        send = torch.empty(P, B, n, tot, D, device=dq.device, dtype=dq.dtype)     # send[s] = chunk s, my heads
        send[:, :, :, :hq_l].copy_(dq.reshape(B, P, n, hq_l, D).permute(1, 0, 2, 3, 4))
        send[:, :, :, hq_l:hq_l + hkv_l].copy_(dk.reshape(B, P, n, hkv_l, D).permute(1, 0, 2, 3, 4))
        send[:, :, :, hq_l + hkv_l:].copy_(dv.reshape(B, P, n, hkv_l, D).permute(1, 0, 2, 3, 4))
        recv = torch.empty_like(send)                                             # recv[r] = heads of rank r
        dist.all_to_all_single(recv, send, group=ctx.group)
        gq = recv[:, :, :, :hq_l].permute(1, 2, 0, 3, 4).reshape(B, n, Hq, D)
        gk = recv[:, :, :, hq_l:hq_l + hkv_l].permute(1, 2, 0, 3, 4).reshape(B, n, Hkv, D)
        gv = recv[:, :, :, hq_l + hkv_l:].permute(1, 2, 0, 3, 4).reshape(B, n, Hkv, D)
        return gq, gk, gv, None


def ulysses_qkv_seq_to_head(q, k, v, group=None):
    """Differentiable fused exchange of q, k and v (one NCCL all-to-all each direction)."""
    P, _ = _group_size_rank(group)
    if P == 1:
        return q, k, v
    return _QKVSeqToHead.apply(q, k, v, group)


def ulysses_seq_to_head(x: torch.Tensor, group=None) -> torch.Tensor:
    """Differentiable all-to-all ``[B, N/P, H, D] -> [B, N, H/P, D]`` (HF layout)."""
    return _SeqToHead.apply(x, group)


def ulysses_head_to_seq(x: torch.Tensor, group=None) -> torch.Tensor:
    """Differentiable inverse all-to-all ``[B, N, H/P, D] -> [B, N/P, H, D]``."""
    return _HeadToSeq.apply(x, group)


# =============================================================================================
# Ulysses over peer memory: the exchange fused into single scatter kernels (no NCCL, no pack / unpack copies)
# =============================================================================================
class _P2PBuffers:
    """Symmetric-memory receive buffers of ONE attention layer (torch symmetric memory gives every rank peer
    mappings of the same allocation on all ranks of the group; the scatter kernel of libsinkfa stores into them
    over NVLink).  Four regions, all in the layout their consumer reads in place:

      qkv_full [B, P*n, hq_l + 2*hkv_l, D]   q | k | v of this rank's heads for all positions   (forward input)
      o_seq    [B, n, Hq, D]                 O of all heads for this rank's positions           (forward output)
      do_full  [B, P*n, hq_l, D]             dO of this rank's heads for all positions          (backward input)
      g_seq    [B, n, Hq + 2*Hkv, D]         dq | dk | dv of all heads for this rank's positions (backward output)

    One cross-rank barrier follows every exchange.  A region is overwritten by the peers only in the next
    forward / backward of the same layer, and at least one such barrier (which every rank reaches only after its
    own reads of the region were issued on the stream) lies in between -- so no barrier is needed in front."""

    def __init__(self, group, B, n, Hq, Hkv, D, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        P, rank = _group_size_rank(group)
        self.P, self.rank, self.key = P, rank, (B, n, Hq, Hkv, D, dtype)
        hq_l, hkv_l = Hq // P, Hkv // P
        self.tot = hq_l + 2 * hkv_l
        sizes = [B * P * n * self.tot * D, B * n * Hq * D, B * P * n * hq_l * D, B * n * (Hq + 2 * Hkv) * D]
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 127) // 128 * 128                  # regions stay 256-byte aligned
        self.buf = symm_mem.empty(total, dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD if group is None else group)
        es = self.buf.element_size()
        ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        self.peer = [[ptr + off * es for ptr in ptrs] for off in offs]     # [region][rank] -> device pointer
        self.qkv_full = self.buf[offs[0]:offs[0] + sizes[0]].view(B, P * n, self.tot, D)
        self.o_seq = self.buf[offs[1]:offs[1] + sizes[1]].view(B, n, Hq, D)
        self.do_full = self.buf[offs[2]:offs[2] + sizes[2]].view(B, P * n, hq_l, D)
        self.g_seq = self.buf[offs[3]:offs[3] + sizes[3]].view(B, n, Hq + 2 * Hkv, D)

        # fused output side: O and dQ are stored into the peers' buffers by the attention kernels themselves (TMA /
        # vector stores to peer mappings): no scatter pass over O / dQ.  (Round 1 shipped routed dQ switched off: on 2
        # GPUs dq and dk differed from run to run.  The cause was a shared-memory race INSIDE the fused backward that
        # the slower remote stores merely exposed -- csrc/bwdf_sm100.cu, "Pass 2(n) reads P back from the image" --
        # fixed in round 2; tools/check_ulysses_p2p.py runs the multi-round bit-exactness check with skewed ranks.)
        # SFA_ULY_NO_ROUTE=1 disables both, SFA_ULY_ROUTE_DQ=0 only the dQ side.
        self.route_o = P <= 8 and os.environ.get("SFA_ULY_NO_ROUTE") is None
        self.route_dq = self.route_o and os.environ.get("SFA_ULY_ROUTE_DQ", "1") != "0"
        if os.environ.get("SFA_ULY_NO_ROUTE_O"):          # diagnostics: routed dQ without routed O
            self.route_o = False
        # The autograd node saves VIEWS of qkv_full for its backward, and the peers overwrite that region in this
        # buffer set's next forward -- remote writes, which autograd's version counters cannot see.  `pending` marks a
        # set whose backward has not run yet; the layer then takes another set (fwd, fwd, bwd, bwd: pipeline
        # micro-batches, weight-shared layers, two loss terms through one module).
        self.pending = False

    def barrier(self):
        self.hdl.barrier(channel=0)


class _UlyssesP2PAttention(torch.autograd.Function):
    """seq->head exchange, attention, head->seq exchange as one autograd node; forward and backward are each
    [scatter kernels, barrier, attention kernels, scatter kernels, barrier] on the current stream."""

    @staticmethod
    def forward(ctx, q, k, v, s_loc, bufs, num_sink, window_size):
        from . import _lib
        P, rank = bufs.P, bufs.rank
        B, n, Hq, D = q.shape
        Hkv = k.shape[2]
        hq_l, hkv_l = Hq // P, Hkv // P
        _lib.ulysses_scatter(q, bufs.peer[0], rank, 0, bufs.tot, 0)
        _lib.ulysses_scatter(k, bufs.peer[0], rank, 0, bufs.tot, hq_l)
        _lib.ulysses_scatter(v, bufs.peer[0], rank, 0, bufs.tot, hq_l + hkv_l)
        bufs.barrier()
        full = bufs.qkv_full
        qh = full[:, :, :hq_l].transpose(1, 2)               # [B, hq_l, P*n, D] views, consumed in place
        kh = full[:, :, hq_l:hq_l + hkv_l].transpose(1, 2)
        vh = full[:, :, hq_l + hkv_l:].transpose(1, 2)
        s32 = _lib._s_aux_f32(s_loc, hq_l)
        o = None
        if bufs.route_o:
            # the forward kernel stores every O tile into the sequence owner's buffer as well (TMA store to a peer
            # mapping): no scatter pass over O.  Shapes the routed kernels do not cover fall back for good.
            try:
                o, lse = _lib.fwd(qh, kh, vh, num_sink, window_size, s32,
                                  o_route=_lib.make_route(bufs.peer[1], n, Hq, rank * hq_l))
            except ValueError:
                bufs.route_o = False
        if o is None:
            o, lse = _lib.fwd(qh, kh, vh, num_sink, window_size, s32)
            _lib.ulysses_scatter(o.transpose(1, 2), bufs.peer[1], rank, 1, Hq, 0)
        bufs.barrier()
        out = bufs.o_seq.clone()                              # the region is reused by the next step
        bufs.pending = any(ctx.needs_input_grad[:4])         # a backward will read the saved views of qkv_full
        ctx.save_for_backward(qh, kh, vh, o, lse, s32 if s32 is not None else torch.empty(0, device=q.device))
        ctx.bufs, ctx.cfg, ctx.has_aux = bufs, (num_sink, window_size, Hq, Hkv), s_loc is not None
        ctx.s_dtype = s_loc.dtype if s_loc is not None else None
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import _lib
        qh, kh, vh, o, lse, s32 = ctx.saved_tensors
        bufs = ctx.bufs
        num_sink, window_size, Hq, Hkv = ctx.cfg
        P, rank = bufs.P, bufs.rank
        hq_l = Hq // P
        _lib.ulysses_scatter(dout, bufs.peer[2], rank, 0, hq_l, 0)
        bufs.barrier()
        do_h = bufs.do_full.transpose(1, 2)
        n = dout.shape[1]
        dq = None
        if bufs.route_dq:
            try:                                             # dQ goes straight to the sequence owners
                dq, dk, dv, ds = _lib.bwd(qh, kh, vh, o, do_h, lse, num_sink, window_size, s32 if ctx.has_aux else None,
                                          dq_route=_lib.make_route(bufs.peer[3], n, Hq + 2 * Hkv, rank * hq_l))
            except ValueError:
                bufs.route_dq = False
        if not bufs.route_dq:
            dq, dk, dv, ds = _lib.bwd(qh, kh, vh, o, do_h, lse, num_sink, window_size, s32 if ctx.has_aux else None)
            _lib.ulysses_scatter(dq.transpose(1, 2), bufs.peer[3], rank, 1, Hq + 2 * Hkv, 0)
        _lib.ulysses_scatter(dk.transpose(1, 2), bufs.peer[3], rank, 1, Hq + 2 * Hkv, Hq)
        _lib.ulysses_scatter(dv.transpose(1, 2), bufs.peer[3], rank, 1, Hq + 2 * Hkv, Hq + Hkv)
        bufs.barrier()
        g = bufs.g_seq
        gq, gk, gv = g[:, :, :Hq].contiguous(), g[:, :, Hq:Hq + Hkv].contiguous(), g[:, :, Hq + Hkv:].contiguous()
        gs = ds.to(ctx.s_dtype) if ctx.has_aux else None
        bufs.pending = False                                  # qkv_full may be overwritten again
        return gq, gk, gv, gs, None, None, None


def slice_s_aux_for_rank(s_aux: Optional[torch.Tensor], local_heads: int, rank: int) -> Optional[torch.Tensor]:
    """Reference rule (verl_patch.py:140-151): rank r owns s_aux[r*H_local : (r+1)*H_local]."""
    if s_aux is None or s_aux.shape[0] == local_heads:
        return s_aux
    assert s_aux.shape[0] % local_heads == 0
    return s_aux[rank * local_heads:(rank + 1) * local_heads]


class UlyssesSinkAttention(torch.nn.Module):
    """Sequence-parallel sink attention over one NVLink/NVSwitch node.

    Inputs are this rank's sequence chunk in HF layout: ``q [B, N/P, H_q, D]``, ``k, v [B, N/P, H_kv, D]``
    (``P`` must divide ``H_kv``); ``s_aux`` holds ALL ``H_q`` logits.  Returns ``O [B, N/P, H_q, D]``.
    """

    MAX_PENDING_FORWARDS = 8

    def __init__(self, num_sink: int = 0, window_size: int = 4096, sp_group=None, head_chunks: int = 1,
                 p2p: bool = False):
        super().__init__()
        self.num_sink = num_sink
        self.window_size = window_size
        self.sp_group = sp_group
        self.head_chunks = max(1, int(head_chunks))
        # p2p=True: the exchange runs as peer-memory scatter kernels of libsinkfa over NVLink instead of NCCL
        # all-to-alls (one node, 16-bit / fp32 CUDA tensors).  The layer then owns symmetric receive buffers sized
        # for its input shape; its saved q/k/v live in them until the layer's backward -- a second forward before that
        # backward takes a second buffer set (see _P2PBuffers.pending).
        self.p2p = bool(p2p)
        self._bufs = []       # symmetric buffer sets; more than one only while several forwards await their backward

    def forward(self, q, k, v, s_aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        from .sink_flash_attention import sink_flash_attention
        P, rank = _group_size_rank(self.sp_group)
        H_q, H_kv = q.shape[2], k.shape[2]
        assert H_kv % P == 0 and H_q % P == 0, "the sequence-parallel size must divide H_kv and H_q"
        hq_l, hkv_l = H_q // P, H_kv // P
        s_loc = slice_s_aux_for_rank(s_aux, hq_l, rank)
        if self.p2p and P > 1:
            B, n, _, D = q.shape
            key = (B, n, H_q, H_kv, D, q.dtype)
            self._bufs = [b for b in self._bufs if b.key == key]
            # every rank runs the same program, so every rank picks (or collectively allocates) the same set
            bufs = next((b for b in self._bufs if not b.pending), None)
            if bufs is None:
                if len(self._bufs) >= self.MAX_PENDING_FORWARDS:
                    raise RuntimeError(
                        f"UlyssesSinkAttention(p2p=True): {len(self._bufs)} forwards of this layer are waiting for their "
                        "backward; each holds a symmetric receive buffer set (raise MAX_PENDING_FORWARDS, or run the "
                        "forwards that need no gradient under torch.no_grad())")
                bufs = _P2PBuffers(self.sp_group, B, n, H_q, H_kv, D, q.dtype, q.device)
                self._bufs.append(bufs)
            return _UlyssesP2PAttention.apply(q, k, v, s_loc, bufs, self.num_sink, self.window_size)
        # head chunks are whole GQA groups so every chunk is an independent attention problem
        nchunk = min(self.head_chunks, hkv_l) if P > 1 else 1
        while hkv_l % nchunk:
            nchunk -= 1
        if nchunk == 1:
            qh, kh, vh = ulysses_qkv_seq_to_head(q, k, v, self.sp_group)
            oh = sink_flash_attention(qh.transpose(1, 2), kh.transpose(1, 2), vh.transpose(1, 2),
                                      self.num_sink, self.window_size, s_loc).transpose(1, 2)
            return ulysses_head_to_seq(oh, self.sp_group)
        # Pipelined: chunk c's attention overlaps chunk c+1's all-to-all (NCCL runs on its own stream).
        g = H_q // H_kv
        kv_per = hkv_l // nchunk
        outs = []
        B, n = q.shape[0], q.shape[1]
        qv = q.reshape(B, n, P, hkv_l, g, -1)
        kv_ = k.reshape(B, n, P, hkv_l, -1)
        vv = v.reshape(B, n, P, hkv_l, -1)
        for c in range(nchunk):
            sl = slice(c * kv_per, (c + 1) * kv_per)
            qc = qv[:, :, :, sl].reshape(B, n, P * kv_per * g, -1)
            kc = kv_[:, :, :, sl].reshape(B, n, P * kv_per, -1)
            vc = vv[:, :, :, sl].reshape(B, n, P * kv_per, -1)
            qh = ulysses_seq_to_head(qc, self.sp_group)
            kh = ulysses_seq_to_head(kc, self.sp_group)
            vh = ulysses_seq_to_head(vc, self.sp_group)
            sc = None if s_loc is None else s_loc[c * kv_per * g:(c + 1) * kv_per * g]
            oh = sink_flash_attention(qh.transpose(1, 2), kh.transpose(1, 2), vh.transpose(1, 2),
                                      self.num_sink, self.window_size, sc).transpose(1, 2)
            outs.append(ulysses_head_to_seq(oh, self.sp_group).reshape(B, n, P, kv_per * g, -1))
        return torch.cat(outs, dim=3).reshape(B, n, H_q, -1)


# =============================================================================================
# Halo-exchange sequence parallelism for narrow windows (no sink tokens)
# =============================================================================================
# With a sliding window of W keys a sequence chunk only needs the last W - 1 keys / values of the chunk before it.
# At the gpt-oss shape (W = 128, 8 KV heads, D = 64) that is 127 x 8 x 64 x 2 B x (K, V) = 260 KB per boundary, against
# the 264 MB per rank (P = 8) that the Ulysses all-to-all moves for the same layer: the exchange drops out of the
# step time and every rank runs the attention kernels on its own chunk with `halo` extra keys in front
# (sink_flash_attention_chunk: the queries are the last rows of a longer key axis).  Backward sends the halo keys'
# dK / dV back to the rank that owns them.  This is what the reference's sequence-chunk helpers (sp_utils.py:28-129)
# are missing -- they broadcast the sink K/V but have no window halo -- and it replaces the
# all-to-all only where the window is narrow and there are no sink tokens; otherwise use UlyssesSinkAttention.
def _halo_rows(window_size: int, n_local: int) -> int:
    halo = (max(window_size - 1, 0) + 15) // 16 * 16        # whole 16-key blocks: the fused backward's granularity
    return min(halo, n_local)


class _HaloFromPrev(torch.autograd.Function):
    """Portable exchange (any torch.distributed backend): my last `halo` rows of [K | V] go to rank + 1, I receive
    rank - 1's (rank 0 receives zeros it never attends).  Backward runs the mirror image with the gradients."""

    @staticmethod
    def forward(ctx, tail, group):
        P, rank = _group_size_rank(group)
        ctx.group = group
        recv = torch.zeros_like(tail)
        ops = []
        gr = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        if rank + 1 < P:
            ops.append(dist.P2POp(dist.isend, tail.contiguous(), gr(rank + 1), group))
        if rank > 0:
            ops.append(dist.P2POp(dist.irecv, recv, gr(rank - 1), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv

    @staticmethod
    def backward(ctx, g):
        group = ctx.group
        P, rank = _group_size_rank(group)
        recv = torch.zeros_like(g)
        ops = []
        gr = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, g.contiguous(), gr(rank - 1), group))
        if rank + 1 < P:
            ops.append(dist.P2POp(dist.irecv, recv, gr(rank + 1), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv, None


class _HaloBuffers:
    """Symmetric-memory buffers of one layer for the peer-memory halo exchange:
         kv_ext [2][B, halo + n, Hkv, D]   rows [0, halo): written by rank - 1 (its last keys / values);
                                           rows [halo, halo + n): this rank's K / V  -> consumed in place by the kernels
         g_halo [2][B, halo, Hkv, D]       dK / dV of this rank's last `halo` keys, written by rank + 1 in backward"""

    def __init__(self, group, B, n, Hkv, D, halo, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        P, rank = _group_size_rank(group)
        self.P, self.rank, self.halo, self.key = P, rank, halo, (B, n, Hkv, D, halo, dtype)
        ext, gh = B * (halo + n) * Hkv * D, B * halo * Hkv * D
        sizes = [ext, ext, gh, gh]
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 127) // 128 * 128
        self.buf = symm_mem.empty(total, dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD if group is None else group)
        shapes = [(B, halo + n, Hkv, D)] * 2 + [(B, halo, Hkv, D)] * 2

        def views(r):
            if r == rank:
                base = self.buf
                return [base[o:o + sz].view(sh) for o, sz, sh in zip(offs, sizes, shapes)]
            return [self.hdl.get_buffer(r, sh, dtype, storage_offset=o) for o, sh in zip(offs, shapes)]
        self.k_ext, self.v_ext, self.gk_halo, self.gv_halo = views(rank)
        self.next = views(rank + 1) if rank + 1 < P else None      # peer mappings: stores go over NVLink
        self.prev = views(rank - 1) if rank > 0 else None
        self.pending = False

    def barrier(self):
        self.hdl.barrier(channel=0)


class _HaloP2PAttention(torch.autograd.Function):
    """[copy K/V into the extended buffer + halo rows into the next rank's buffer, barrier, attention kernels] and the
    mirror image in backward: plain kernels on one stream (CUDA-graph capturable), no NCCL call."""

    @staticmethod
    def forward(ctx, q, k, v, s_aux, bufs, window_size):
        from . import _lib
        B, n, Hq, D = q.shape
        halo, rank = bufs.halo, bufs.rank
        bufs.k_ext[:, halo:].copy_(k)
        bufs.v_ext[:, halo:].copy_(v)
        if bufs.next is not None:
            bufs.next[0][:, :halo].copy_(k[:, n - halo:])           # peer stores
            bufs.next[1][:, :halo].copy_(v[:, n - halo:])
        bufs.barrier()
        s32 = _lib._s_aux_f32(s_aux, Hq)
        qh = q.transpose(1, 2)
        if rank == 0:
            kh, vh = bufs.k_ext[:, halo:].transpose(1, 2), bufs.v_ext[:, halo:].transpose(1, 2)
            ext = None
        else:
            kh, vh = bufs.k_ext.transpose(1, 2), bufs.v_ext.transpose(1, 2)
            ext = _lib.make_ext(n, halo + n, halo)
        o, lse = _lib.fwd(qh, kh, vh, 0, window_size, s32, ext=ext)
        bufs.pending = any(ctx.needs_input_grad[:4])
        ctx.save_for_backward(qh, kh, vh, o, lse, s32 if s32 is not None else torch.empty(0, device=q.device))
        ctx.bufs, ctx.cfg = bufs, (window_size, n, s_aux is not None, s_aux.dtype if s_aux is not None else None)
        return o.transpose(1, 2)

    @staticmethod
    def backward(ctx, dout):
        from . import _lib
        qh, kh, vh, o, lse, s32 = ctx.saved_tensors
        bufs = ctx.bufs
        window_size, n, has_aux, s_dtype = ctx.cfg
        halo, rank = bufs.halo, bufs.rank
        ext = None if rank == 0 else _lib.make_ext(n, halo + n, halo)
        dq, dk, dv, ds = _lib.bwd(qh, kh, vh, o, dout.transpose(1, 2), lse, 0, window_size, s32 if has_aux else None, ext=ext)
        dk, dv = dk.transpose(1, 2), dv.transpose(1, 2)              # [B, (halo +) n, Hkv, D]
        if bufs.prev is not None:                                   # the halo keys belong to rank - 1
            bufs.prev[2].copy_(dk[:, :halo])
            bufs.prev[3].copy_(dv[:, :halo])
        bufs.barrier()
        if rank > 0:
            dk, dv = dk[:, halo:], dv[:, halo:]
        if bufs.next is not None:
            dk[:, n - halo:] += bufs.gk_halo
            dv[:, n - halo:] += bufs.gv_halo
        bufs.pending = False
        return dq.transpose(1, 2), dk, dv, (ds.to(s_dtype) if has_aux else None), None, None


class HaloSinkAttention(torch.nn.Module):
    """Sequence-chunk parallel sink attention for narrow sliding windows without sink tokens.

    Inputs are this rank's sequence chunk in HF layout: ``q [B, n, H_q, D]``, ``k, v [B, n, H_kv, D]`` (rank r holds
    positions ``[r n, (r + 1) n)``; ``window_size - 1 <= n``); ``s_aux`` holds all ``H_q`` logits.  Returns
    ``O [B, n, H_q, D]``.  ``s_aux.grad`` is this rank's partial sum over its positions: sum it over the group (as
    the trainer does for every replicated parameter), or pass ``reduce_s_aux_grad=True``.

    ``p2p=True`` runs the exchange as stores into the neighbours' symmetric-memory buffers (one node, CUDA); the
    default uses ``torch.distributed`` point-to-point ops and works with any backend.
    """

    MAX_PENDING_FORWARDS = 8

    def __init__(self, window_size: int = 128, sp_group=None, p2p: bool = False, reduce_s_aux_grad: bool = False):
        super().__init__()
        self.window_size = int(window_size)
        self.sp_group = sp_group
        self.p2p = bool(p2p)
        self.reduce_s_aux_grad = bool(reduce_s_aux_grad)
        self._bufs = []

    def forward(self, q, k, v, s_aux: Optional[torch.Tensor] = None) -> torch.Tensor:
        from .sink_flash_attention import sink_flash_attention, sink_flash_attention_chunk
        P, rank = _group_size_rank(self.sp_group)
        B, n, H_q, D = q.shape
        H_kv = k.shape[2]
        W = self.window_size
        if s_aux is not None and self.reduce_s_aux_grad and P > 1:
            s_aux = _AllReduceGrad.apply(s_aux, self.sp_group)
        if P == 1:
            return sink_flash_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), 0, W, s_aux).transpose(1, 2)
        assert W - 1 <= n, f"halo exchange needs window_size - 1 ({W - 1}) <= chunk length ({n}); use UlyssesSinkAttention"
        halo = _halo_rows(W, n)
        if self.p2p:
            key = (B, n, H_kv, D, halo, q.dtype)
            self._bufs = [b for b in self._bufs if b.key == key]
            bufs = next((b for b in self._bufs if not b.pending), None)
            if bufs is None:
                if len(self._bufs) >= self.MAX_PENDING_FORWARDS:
                    raise RuntimeError("HaloSinkAttention(p2p=True): too many forwards waiting for their backward")
                bufs = _HaloBuffers(self.sp_group, B, n, H_kv, D, halo, q.dtype, q.device)
                self._bufs.append(bufs)
            return _HaloP2PAttention.apply(q, k, v, s_aux, bufs, W)
        tail = torch.cat([k[:, n - halo:], v[:, n - halo:]], dim=2)          # [B, halo, 2 H_kv, D]
        recv = _HaloFromPrev.apply(tail, self.sp_group)
        if rank == 0:
            o = sink_flash_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), 0, W, s_aux).transpose(1, 2)
            return o + (recv.sum() * 0).to(o.dtype)     # keeps the exchange in the graph: its backward must run on every rank
        k_ext = torch.cat([recv[:, :, :H_kv], k], dim=1)
        v_ext = torch.cat([recv[:, :, H_kv:], v], dim=1)
        return sink_flash_attention_chunk(q.transpose(1, 2), k_ext.transpose(1, 2), v_ext.transpose(1, 2), 0, W,
                                          s_aux).transpose(1, 2)


class _AllReduceGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


# =============================================================================================
# Reference sequence-chunk helpers (same names / semantics as sp_utils.py:28-180)
# =============================================================================================
def prepare_sink_kv_for_sp(k: torch.Tensor, v: torch.Tensor, num_sink: int, sp_group,
                           rank: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Broadcast the first ``num_sink`` K/V rows from group-rank 0 and prepend them on the other ranks."""
    if num_sink == 0:
        return k, v
    if rank is None:
        rank = dist.get_rank(sp_group)
    B, H, _, D = k.shape
    if rank == 0:
        sk, sv = k[:, :, :num_sink].contiguous(), v[:, :, :num_sink].contiguous()
    else:
        sk = torch.empty(B, H, num_sink, D, device=k.device, dtype=k.dtype)
        sv = torch.empty(B, H, num_sink, D, device=v.device, dtype=v.dtype)
    src = dist.get_global_rank(sp_group, 0) if sp_group is not None else 0
    dist.broadcast(sk, src=src, group=sp_group)
    dist.broadcast(sv, src=src, group=sp_group)
    if rank == 0:
        return k, v
    return torch.cat([sk, k], dim=2), torch.cat([sv, v], dim=2)


def reduce_sink_kv_grads(dk: torch.Tensor, dv: torch.Tensor, num_sink: int, sp_group,
                         rank: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sum the sink rows' gradients over the group; rank 0 keeps them, the others strip the prefix."""
    if num_sink == 0:
        return dk, dv
    if rank is None:
        rank = dist.get_rank(sp_group)
    sdk, sdv = dk[:, :, :num_sink].contiguous(), dv[:, :, :num_sink].contiguous()
    dist.all_reduce(sdk, op=dist.ReduceOp.SUM, group=sp_group)
    dist.all_reduce(sdv, op=dist.ReduceOp.SUM, group=sp_group)
    if rank == 0:
        dk, dv = dk.clone(), dv.clone()
        dk[:, :, :num_sink] = sdk
        dv[:, :, :num_sink] = sdv
        return dk, dv
    return dk[:, :, num_sink:], dv[:, :, num_sink:]


def get_local_position_offset(rank: int, n_local: int, num_sink: int) -> int:
    """Global position of the first local (non-prepended) token of ``rank``'s chunk."""
    return rank * n_local


class SinkAttentionSPWrapper(torch.nn.Module):
    """Reference-compatible wrapper (sp_utils.py:151-180).  Without a group (or with a group of one)
    it is plain ``sink_flash_attention``.  With a group it runs the Ulysses path on kernel-layout
    ``[B, H, N/P, D]`` chunks -- the reference's own multi-rank branch cannot run (it feeds a K longer
    than Q into an op that asserts equal lengths)."""

    def __init__(self, num_sink: int = 4, window_size: int = 4096, sp_group=None):
        super().__init__()
        self.num_sink = num_sink
        self.window_size = window_size
        self.sp_group = sp_group

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        from .sink_flash_attention import sink_flash_attention
        if self.sp_group is None or dist.get_world_size(self.sp_group) == 1:
            return sink_flash_attention(q, k, v, self.num_sink, self.window_size)
        uly = UlyssesSinkAttention(self.num_sink, self.window_size, self.sp_group)
        return uly(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
