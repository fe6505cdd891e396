"""Seeded case table shared by oracle/make_golden.py (writer, build container only) and
tests/ (reader).  TEST INFRASTRUCTURE ONLY.  Inputs are regenerated from the CPU generator seed
(same torch build in the image on both boxes); the fixtures carry an input checksum plus the
REFERENCE's outputs, so a generator drift is detected rather than silently compared."""
import torch

# (name, B, Hq, Hkv, N, D, S, W, s_aux?, run_triton_interpreter, store_grads)
PREFILL_CASES = [
    ("c0_small",    1, 8, 8, 256, 64, 4, 128, False, False, False),  # BASELINE.json configs[0]
    ("mha_s4_w32",  1, 2, 2, 128, 64, 4, 32,  False, True,  True),   # tests/test_sink_attention.py:187
    ("gqa_s4_w64",  1, 4, 1, 160, 64, 4, 64,  False, False, True),   # :189 (GQA 4:1)
    ("b2_s1_w64",   2, 2, 2, 96,  64, 1, 64,  False, False, True),   # :190
    ("d128_s4_w64", 1, 2, 2, 96, 128, 4, 64,  False, False, True),   # :191
    ("saux_full",   1, 4, 1, 128, 64, 0, 128, True,  True,  True),   # tests/test_s_aux.py:80-99
    ("saux_win",    1, 4, 1, 256, 64, 0, 128, True,  False, True),   # :103-122
    ("saux_d80",    1, 4, 1, 64,  80, 0, 64,  True,  False, True),   # :294-314 (eager only)
    ("ragged_n96",  1, 4, 2, 96,  64, 4, 40,  True,  True,  True),   # N not a tile multiple
    ("w1_sink",     1, 2, 2, 128, 64, 4, 1,   False, False, True),   # tests/test_sink_attention.py:119-131
    ("s0_wN",       1, 2, 2, 128, 64, 0, 128, False, False, True),   # :99-116 full causal
    ("sink_gt_n",   1, 2, 2, 40,  16, 50, 8,  True,  True,  True),   # SURVEY 4.5 edge: num_sink > N
    ("sinks_span",  1, 2, 1, 150, 32, 70, 16, True,  False, True),   # sinks spanning several KV tiles
    ("mqa_4_1",     1, 4, 1, 128, 64, 4, 32,  True,  False, True),   # benchmark.py MQA
    ("w0_saux",     1, 2, 2, 33,  16, 0, 0,   True,  True,  True),   # window 0, no sinks: O=0, LSE=s_aux
]

# (name, B, Hq, Hkv, Nkv, D, s_aux: False|True|"big", run_triton_interpreter)
DECODE_CASES = [
    ("dec_mha_64",    1, 8, 8, 64,   64,  False, True),
    ("dec_gqa_300",   2, 16, 4, 300, 64,  True,  True),    # ragged N_kv
    ("dec_gqa8_1028", 1, 64, 8, 1028, 64, True,  False),   # gpt-oss head ratio
    ("dec_d128",      1, 32, 8, 516, 128, True,  False),
    ("dec_d256",      1, 4, 4, 256, 256,  False, False),
    ("dec_mqa",       1, 8, 1, 512, 64,   True,  False),
    ("dec_saux100",   1, 8, 8, 128, 64,   "big", False),   # tests/test_decode_kernel.py:165-186
    ("dec_d32",       1, 4, 2, 77,  32,   True,  True),
]


def prefill_inputs(case):
    name, B, Hq, Hkv, N, D, S, W, use_aux = case[:9]
    g = torch.Generator().manual_seed(1234 + 7 * len(name) + N + 13 * D)
    q = torch.randn(B, Hq, N, D, generator=g)
    k = torch.randn(B, Hkv, N, D, generator=g)
    v = torch.randn(B, Hkv, N, D, generator=g)
    do = torch.randn(B, Hq, N, D, generator=g)
    s_aux = (torch.randn(Hq, generator=g) * 0.5 + 0.5) if use_aux else None
    return q, k, v, do, s_aux


def decode_inputs(case):
    name, B, Hq, Hkv, Nkv, D, use_aux = case[:7]
    g = torch.Generator().manual_seed(99 + Nkv + D + 3 * len(name))
    q = torch.randn(B, Hq, 1, D, generator=g)
    k = torch.randn(B, Hkv, Nkv, D, generator=g)
    v = torch.randn(B, Hkv, Nkv, D, generator=g)
    if use_aux == "big":
        s_aux = torch.full((Hq,), 100.0)
    elif use_aux:
        s_aux = torch.randn(Hq, generator=g) * 0.5 + 1.0
    else:
        s_aux = None
    return q, k, v, s_aux


def checksum(*tensors):
    return float(sum(t.double().abs().sum().item() for t in tensors if t is not None))
