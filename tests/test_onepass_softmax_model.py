"""Numerical model (numpy, CPU) of the provisional-shift one-pass softmax that `enc_attention_tc2_kernel<NKB, 2>` (B200RANK_ATTN=tc4,
llm-rankers_b200/csrc/attention_tc.cuh) implements, next to the exact-maximum two-pass form of the shipped kernel: the shift is the maximum of
the row's first 32 keys, p = 2^(v - shift) is rounded to bf16 for the P.V product, the row sum stays fp32, and a row whose sum leaves
[1/2, 2^100) is redone with its exact maximum. The kernel itself can only be checked on a GPU (tests/test_engine_gpu.py::
test_enc_attention_onepass_vs_numpy); this file pins the arithmetic argument the kernel rests on (modeling_t5.py:308-334 is the reference
softmax: fp32, exact maximum)."""
import numpy as np

LOG2E = np.float32(1.4426950408889634)


def _bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000          # round to nearest even
    return u.astype(np.uint32).view(np.float32)


def _attend(v, V, shift):
    with np.errstate(over="ignore", invalid="ignore"):
        p = np.exp2((v - shift[:, None]).astype(np.float32)).astype(np.float32)
        l = p.sum(axis=1, dtype=np.float32)
        o = (_bf16(p).astype(np.float64) @ V.astype(np.float64)).astype(np.float32)
    return o, l


def two_pass(v, V):
    o, l = _attend(v, V, v.max(axis=1))
    return o / l[:, None]


def one_pass(v, V):
    shift = v[:, :32].max(axis=1)
    o, l = _attend(v, V, shift)
    bad = ~((l >= 0.5) & (l < np.float32(2.0 ** 100)))
    if bad.any():
        o2, l2 = _attend(v[bad], V, v[bad].max(axis=1))
        o[bad], l[bad] = o2, l2
    return o / l[:, None], int(bad.sum())


def test_provisional_shift_matches_exact_maximum_softmax():
    rng = np.random.default_rng(4)
    for n_keys, scale in [(184, 3.0), (184, 30.0), (33, 3.0), (192, 80.0), (20, 5.0)]:
        v = (rng.standard_normal((128, n_keys)) * scale).astype(np.float32) * LOG2E
        V = _bf16(rng.standard_normal((n_keys, 64)).astype(np.float32))
        ref = two_pass(v, V)
        got, redone = one_pass(v, V)
        assert np.isfinite(got).all()
        # identical blocking and bf16 rounding of P; the only difference is the power-of-two-ish scale of p before rounding
        assert np.abs(got - ref).max() <= 2.0 ** -7 * np.abs(ref).max(), (n_keys, scale)
        if scale >= 80.0:
            assert redone > 0      # the exact-maximum redo is exercised
        if scale <= 5.0:
            assert redone == 0     # ordinary score ranges never leave the fast path


def test_rows_that_climb_after_the_first_chunk_are_redone():
    v = np.full((4, 184), -50.0, np.float32)
    v[:, 100] = 300.0                                   # 350 log2 units above everything in the first chunk: 2^350 overflows fp32
    V = _bf16(np.arange(184 * 64, dtype=np.float32).reshape(184, 64) / 1000)
    got, redone = one_pass(v, V)
    assert redone == 4
    assert np.allclose(got, np.broadcast_to(V[100], (4, 64)), rtol=2.0 ** -7)
