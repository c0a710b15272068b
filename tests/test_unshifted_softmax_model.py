"""Numerical model (numpy, CPU) of the UNSHIFTED one-pass softmax of the default encoder-attention kernel (csrc/attention_tc5.cuh):
p_j = 2^(v_j) with v = log2(e) * (score + bias), no row maximum, P rounded to bf16, l = sum_j p_j in fp32, O = (P V) / l, and a row is
left to the exact fix-up walk iff l is outside [2^-100, 2^100) or not finite. The claims checked here are the ones the kernel's
correctness argument rests on (the GPU test drives the kernel itself with scores in the hundreds):
  1. inside the window the unshifted form equals the max-shifted softmax to fp32 / bf16 rounding — fp32 and bf16 share one exponent
     range, so scaling every p by 2^max changes no mantissa;
  2. the [2^-100, 2^100) test on l never lets an inaccurate row through: whenever the row maximum is within +-100 - 8 (8 = log2 of the
     192-key row length, rounded up) the row passes and is accurate, and whenever the unshifted arithmetic would overflow or flush the
     dominant terms to zero the test fails (row goes to the fix-up walk);
  3. terms flushed to zero by ex2.approx.ftz (below 2^-126) are at least 2^-26 below the row sum: invisible at bf16 / fp32 precision."""
import numpy as np


def bf16_round(x):
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def ex2_ftz(v):
    with np.errstate(over="ignore", under="ignore"):
        p = np.exp2(np.asarray(v, np.float64)).astype(np.float32)
    p[np.abs(p) < np.float32(2.0 ** -126)] = 0.0          # .ftz: subnormal results flush to zero
    return p


def unshifted_row(v, vals):
    p = ex2_ftz(v)
    with np.errstate(over="ignore", invalid="ignore"):
        l = np.float32(p.astype(np.float32).sum(dtype=np.float32))
        ok = bool(np.isfinite(l) and l >= np.float32(2.0 ** -100) and l < np.float32(2.0 ** 100))
        o = (bf16_round(p).astype(np.float64) @ vals) / np.float64(l) if ok else None
    return ok, o


def exact_row(v, vals):
    v = np.asarray(v, np.float64)
    p = np.exp2(v - v.max())
    return (p @ vals) / p.sum()


def test_inside_the_window_unshifted_equals_shifted_softmax():
    rng = np.random.default_rng(0)
    for centre in (-92.0, -40.0, 0.0, 37.5, 84.0):      # row maximum <= 92 = 100 - log2(192 keys, rounded up)
        for spread in (0.5, 4.0, 30.0):
            v = (centre + spread * rng.standard_normal(184)).astype(np.float32)
            v = np.minimum(v, np.float32(centre + 7.9))      # keep the row maximum inside the stated window
            vals = rng.standard_normal((184, 8))
            ok, o = unshifted_row(v, vals)
            assert ok, (centre, spread)
            ref = exact_row(v, vals)
            assert np.abs(o - ref).max() <= 6e-3 * max(1.0, np.abs(ref).max()), (centre, spread)   # bf16 P: 2^-9 relative per term


def test_the_row_sum_test_rejects_every_row_the_unshifted_form_cannot_represent():
    rng = np.random.default_rng(1)
    vals = rng.standard_normal((184, 8))
    for centre in (-300.0, -135.0, -109.0, 101.0, 127.0, 128.5, 400.0, np.inf, np.nan):
        v = (centre + rng.standard_normal(184)).astype(np.float32)
        ok, _ = unshifted_row(v, vals)
        assert not ok, centre
    # one dominant key far above a sea of tiny ones: representable (the tiny ones flush to zero, 2^-26 below the sum)
    v = np.full(184, -140.0, np.float32)
    v[17] = -99.0
    ok, o = unshifted_row(v, vals)
    assert ok and np.abs(o - exact_row(v, vals)).max() <= 6e-3 * np.abs(vals[17]).max() + 1e-6


def test_both_half_row_threads_take_the_same_decision():
    """The two threads of a row add the same two partial sums in the same order (l = partial[0] + partial[1]) — the decision and the
    normaliser are bit-identical in both, so a row is either stored by both or left to the fix-up walk by both."""
    rng = np.random.default_rng(2)
    for _ in range(200):
        v = (rng.uniform(-110, 110) + 20 * rng.standard_normal(184)).astype(np.float32)
        p = ex2_ftz(v)
        with np.errstate(over="ignore"):
            a, b = np.float32(p[:96].sum(dtype=np.float32)), np.float32(p[96:].sum(dtype=np.float32))
            l0, l1 = np.float32(a + b), np.float32(a + b)
        assert (l0 == l1) or (np.isnan(l0) and np.isnan(l1))
