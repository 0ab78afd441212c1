"""The arithmetic claim behind k_conv_h / k_gin<true>, checked in numpy: the fp16 two-term split
x = hi + lo * 2^-11 carries 22 significant bits, and hi.Whi + (hi.Wlo + lo.Whi) * 2^-11 is as accurate as 3xTF32
(hi.hi + hi.lo + lo.hi on 10-bit mantissas) relative to sum |x||w| -- for values anywhere in the guarded range."""
import numpy as np


def split_h(a):
    hi = a.astype(np.float16)
    lo = ((a - hi.astype(np.float32)) * np.float32(2048)).astype(np.float16)
    return hi, lo


def tf32(a):
    return ((a.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_split_reconstructs_22_bits_over_the_guarded_range():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * 10.0 ** rng.uniform(-4, 4.7, 200000)).astype(np.float32)
    x = x[np.abs(x) <= 60000]                      # beyond that the producers raise the range flag (TG_H_LIMIT)
    hi, lo = split_h(x)
    assert np.isfinite(hi.astype(np.float32)).all()
    rec = hi.astype(np.float64) + lo.astype(np.float64) / 2048
    rel = np.abs(rec - x.astype(np.float64)) / np.abs(x.astype(np.float64))
    assert rel[np.abs(x) >= 2.0 ** -14].max() <= 2.0 ** -21          # normal fp16 range of hi
    assert np.abs(rec - x.astype(np.float64)).max() <= 60000 * 2.0 ** -21
    tiny = np.float32([1e-6, 3e-7, -5e-8])         # below fp16's normal range: absolute accuracy 2^-35
    h, l = split_h(tiny)
    assert np.abs(h.astype(np.float64) + l.astype(np.float64) / 2048 - tiny.astype(np.float64)).max() <= 2.0 ** -35


def test_three_products_match_3xtf32_accuracy():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((4096, 32)) * rng.choice([1e-3, 1.0, 30.0], size=(4096, 1))).astype(np.float32)
    w = (1 / (1 + np.exp(-2 * rng.standard_normal((32, 32))))).astype(np.float32)        # sigmoid weights in (0, 1)
    ref = x.astype(np.float64) @ w.astype(np.float64)
    den = np.abs(x).astype(np.float64) @ w.astype(np.float64)
    xh, xl = (a.astype(np.float64) for a in split_h(x))
    wh, wl = (a.astype(np.float64) for a in split_h(w))
    err_h = (np.abs(xh @ wh + (xh @ wl + xl @ wh) / 2048 - ref) / den).max()
    xt = tf32(x.copy()); xtl = tf32((x - xt).copy()); wt = tf32(w.copy()); wtl = tf32((w - wt).copy())
    err_t = (np.abs(xt.astype(np.float64) @ wt + xt.astype(np.float64) @ wtl + xtl.astype(np.float64) @ wt - ref) / den).max()
    assert err_h <= 2.5e-7 and err_h <= 1.5 * err_t, (err_h, err_t)


def test_sum_of_split_rows_in_half_precision_keeps_22_bits():
    """k_conv_z adds the split copies of several same-type neighbour rows WITHOUT unpacking them (conv_z.cu):
    hi' = fl16(hi_a + hi_b), err = (hi_a + hi_b) - hi' (TwoSum, exact in fp16), lo' = fl16(lo_a + lo_b + 2^11 err).
    The new pair must carry x_a + x_b (+ x_c) as accurately as a fresh split of the fp32 sum."""
    rng = np.random.default_rng(2)
    f16 = np.float16

    def two_sum(a, b):
        s = (a + b).astype(f16)
        bb = (s - a).astype(f16)
        err = ((a - (s - bb).astype(f16)).astype(f16) + (b - bb).astype(f16)).astype(f16)
        return s, err

    n = 200000
    xs = [(rng.standard_normal(n) * 10.0 ** rng.uniform(-2, 3.5, n)).astype(np.float32) for _ in range(3)]
    hi, lo = split_h(xs[0])
    exact = hi.astype(np.float64) + lo.astype(np.float64) / 2048
    mag = np.abs(exact)
    for x in xs[1:]:
        h2, l2 = split_h(x)
        exact = exact + h2.astype(np.float64) + l2.astype(np.float64) / 2048
        mag = mag + np.abs(h2.astype(np.float64))
        hi, err = two_sum(hi, h2)
        # fused multiply-add in fp16: evaluate in fp64 (exact for these operands), round once
        lo = (err.astype(np.float64) * 2048 + (lo + l2).astype(f16).astype(np.float64)).astype(f16)
        assert np.isfinite(hi.astype(np.float32)).all() and np.isfinite(lo.astype(np.float32)).all()
        # TwoSum is exact: hi' + err == hi_a + hi_b
    rec = hi.astype(np.float64) + lo.astype(np.float64) / 2048
    rel = np.abs(rec - exact) / mag                      # relative to sum |x|, as the product accuracy is stated
    assert rel.max() <= 2.0 ** -20, rel.max()
