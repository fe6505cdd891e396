"""World-size-2 `gloo` tests of the Ulysses sequence-parallel path (CPU, no GPU needed).

The all-to-all plumbing, the head/sequence bookkeeping, the s_aux slice rule (reference verl_patch.py:140-151)
and the autograd wiring are exercised with the CPU oracle standing in for the CUDA operator (the product
has no CPU kernels); the result of the sharded run must equal the un-sharded oracle on the full sequence."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "sink-flash-attention-kernel_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sink_oracle as orc
        import sink_attention  # noqa: F401
        from sink_attention import sp_utils
        sfa_mod = sys.modules["sink_attention.sink_flash_attention"]   # the package re-exports a function of the same name

        # CPU stand-in for the CUDA operator: the oracle's differentiable eager path
        sfa_mod.sink_flash_attention = lambda q, k, v, num_sink=4, window_size=512, s_aux=None: \
            orc.eager_sink_attention(q, k, v, num_sink, window_size, s_aux)

        B, N, Hq, Hkv, D, S, W = 2, 48, 8, 4, 16, 2, 11
        g = torch.Generator().manual_seed(3)
        q = torch.randn(B, N, Hq, D, generator=g, dtype=torch.float64)
        k = torch.randn(B, N, Hkv, D, generator=g, dtype=torch.float64)
        v = torch.randn(B, N, Hkv, D, generator=g, dtype=torch.float64)
        do = torch.randn(B, N, Hq, D, generator=g, dtype=torch.float64)
        s_aux = torch.randn(Hq, generator=g, dtype=torch.float64)
        n = N // world
        sl = slice(rank * n, (rank + 1) * n)

        # 1. plumbing: seq->head gives the full sequence of this rank's heads; head->seq inverts it
        qh = sp_utils.ulysses_seq_to_head(q[:, sl].contiguous())
        hq_l = Hq // world
        assert torch.equal(qh, q[:, :, rank * hq_l:(rank + 1) * hq_l])
        assert torch.equal(sp_utils.ulysses_head_to_seq(qh), q[:, sl])
        q3, k3, v3 = sp_utils.ulysses_qkv_seq_to_head(q[:, sl].contiguous(), k[:, sl].contiguous(), v[:, sl].contiguous())
        hkv_l = Hkv // world
        assert torch.equal(q3, q[:, :, rank * hq_l:(rank + 1) * hq_l])
        assert torch.equal(k3, k[:, :, rank * hkv_l:(rank + 1) * hkv_l])
        assert torch.equal(v3, v[:, :, rank * hkv_l:(rank + 1) * hkv_l])

        # 2. the sharded operator (fused q/k/v exchange, and the head-chunk pipelined variant) == un-sharded oracle
        ref_in = [t.clone().requires_grad_(True) for t in (q, k, v, s_aux)]
        o_ref = orc.eager_sink_attention(ref_in[0].transpose(1, 2), ref_in[1].transpose(1, 2), ref_in[2].transpose(1, 2),
                                         S, W, ref_in[3]).transpose(1, 2)
        o_ref.backward(do)
        for chunks in (1, 2):
            loc = [q[:, sl].clone().requires_grad_(True), k[:, sl].clone().requires_grad_(True),
                   v[:, sl].clone().requires_grad_(True), s_aux.clone().requires_grad_(True)]
            uly = sp_utils.UlyssesSinkAttention(num_sink=S, window_size=W, sp_group=None, head_chunks=chunks)
            o = uly(loc[0], loc[1], loc[2], loc[3])
            o.backward(do[:, sl])
            assert torch.allclose(o, o_ref[:, sl], atol=1e-10), f"O mismatch (chunks={chunks})"
            for got, ref in zip(loc[:3], ref_in[:3]):
                assert torch.allclose(got.grad, ref.grad[:, sl], atol=1e-10), f"grad mismatch (chunks={chunks})"
            # ds_aux: every rank holds the gradient of ITS heads (the slice rule); the others stay zero
            gs = loc[3].grad
            assert torch.allclose(gs[rank * hq_l:(rank + 1) * hq_l], ref_in[3].grad[rank * hq_l:(rank + 1) * hq_l], atol=1e-10)
            other = torch.cat([gs[:rank * hq_l], gs[(rank + 1) * hq_l:]])
            assert float(other.abs().max()) == 0.0
        out[rank] = "ok"
    except Exception as e:  # surface the failure in the parent
        out[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ulysses_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}, dict(out)


def _halo_worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "sink-flash-attention-kernel_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import math
        import sink_oracle as orc
        import sink_attention  # noqa: F401
        from sink_attention import sp_utils
        sfa_mod = sys.modules["sink_attention.sink_flash_attention"]

        def eager_chunk(q, k, v, num_sink=4, window_size=512, s_aux=None, q_offset=None):
            """CPU stand-in for the CUDA operator with N_q <= N_kv: the oracle's eager masked softmax with the
            queries at the LAST positions of the key axis (differentiable)."""
            B, Hq, Nq, D = q.shape
            Nkv = k.shape[2]
            off = Nkv - Nq if q_offset is None else q_offset
            g = Hq // k.shape[1]
            ke, ve = k.repeat_interleave(g, dim=1), v.repeat_interleave(g, dim=1)
            w = torch.matmul(q, ke.transpose(-2, -1)) / math.sqrt(D)
            valid = orc.attended_mask(Nkv, num_sink, window_size)[off:off + Nq]
            w = w + ((~valid).to(w.dtype) * (-1e9))
            if s_aux is not None:
                col = s_aux.to(w.dtype).reshape(1, Hq, 1, 1).expand(B, Hq, Nq, 1)
                comb = torch.cat([w, col], dim=-1)
                probs = torch.softmax(comb - comb.max(dim=-1, keepdim=True).values, dim=-1)[..., :-1]
            else:
                probs = torch.softmax(w, dim=-1)
            return torch.matmul(probs, ve)
        sfa_mod.sink_flash_attention = lambda q, k, v, num_sink=4, window_size=512, s_aux=None: \
            orc.eager_sink_attention(q, k, v, num_sink, window_size, s_aux)
        sfa_mod.sink_flash_attention_chunk = eager_chunk

        B, N, Hq, Hkv, D, W = 2, 96, 4, 2, 8, 20            # halo = 32 rows (W - 1 rounded up to 16-key blocks)
        g = torch.Generator().manual_seed(5)
        q = torch.randn(B, N, Hq, D, generator=g, dtype=torch.float64)
        k = torch.randn(B, N, Hkv, D, generator=g, dtype=torch.float64)
        v = torch.randn(B, N, Hkv, D, generator=g, dtype=torch.float64)
        do = torch.randn(B, N, Hq, D, generator=g, dtype=torch.float64)
        s_aux = torch.randn(Hq, generator=g, dtype=torch.float64)
        n = N // world
        sl = slice(rank * n, (rank + 1) * n)
        ref_in = [t.clone().requires_grad_(True) for t in (q, k, v, s_aux)]
        o_ref = orc.eager_sink_attention(ref_in[0].transpose(1, 2), ref_in[1].transpose(1, 2), ref_in[2].transpose(1, 2),
                                         0, W, ref_in[3]).transpose(1, 2)
        o_ref.backward(do)
        for reduce_s in (False, True):
            loc = [q[:, sl].clone().requires_grad_(True), k[:, sl].clone().requires_grad_(True),
                   v[:, sl].clone().requires_grad_(True), s_aux.clone().requires_grad_(True)]
            halo = sp_utils.HaloSinkAttention(window_size=W, sp_group=None, reduce_s_aux_grad=reduce_s)
            o = halo(loc[0], loc[1], loc[2], loc[3])
            o.backward(do[:, sl])
            assert torch.allclose(o, o_ref[:, sl], atol=1e-10), "O mismatch"
            for name, got, ref in zip("qkv", loc[:3], ref_in[:3]):
                assert torch.allclose(got.grad, ref.grad[:, sl], atol=1e-10), f"d{name} mismatch (rank {rank})"
            gs = loc[3].grad.clone()
            if not reduce_s:
                dist.all_reduce(gs)                       # the partial sums over the chunks add up to the full gradient
            assert torch.allclose(gs, ref_in[3].grad, atol=1e-10), "ds_aux mismatch"
        out[rank] = "ok"
    except Exception as e:  # surface the failure in the parent
        import traceback
        out[rank] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_halo_exchange_world_size_3_gloo():
    """Halo-exchange sequence parallelism (narrow window, no sink tokens): three ranks, so that one rank both sends and
    receives a halo in each direction; the sharded result must equal the un-sharded oracle incl. all gradients."""
    world = 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok", 2: "ok"}, dict(out)
