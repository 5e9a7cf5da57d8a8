"""The 2-D path's distance chains replace x / d by x * (1 / d) when d is a power of two (csrc/ps2d.cu: pow2_reciprocal, DistSlot) and
claim bit-identical results: both are the correctly rounded value of the same real number.  Checked here on the CPU in IEEE double
arithmetic (numpy), over random bit patterns including subnormal operands and results, and the reciprocal's bit construction
(exponent field 2046 - e) over every exponent the kernel accepts."""
import numpy as np


def _pow2_reciprocal_bits(d_bits):
    """the kernel's test and construction, on raw bits: returns (is_pow2, reciprocal_bits)"""
    ex = (d_bits >> np.uint64(52)).astype(np.uint64)              # sign must be 0 and the mantissa zero, or the test below fails
    ok = ((d_bits & np.uint64(0x800FFFFFFFFFFFFF)) == 0) & (ex >= 2) & (ex <= 2044)
    inv = (np.uint64(2046) - np.minimum(ex, np.uint64(2046))) << np.uint64(52)
    return ok, inv


def test_reciprocal_construction_is_exact_for_every_accepted_exponent():
    ex = np.arange(0, 2048, dtype=np.uint64)
    d_bits = ex << np.uint64(52)
    ok, inv_bits = _pow2_reciprocal_bits(d_bits)
    assert ok.sum() == 2043 and not ok[0] and not ok[1] and not ok[2045] and not ok[2047]   # zero / subnormal / huge / inf are refused
    d = d_bits.view(np.float64)[ok]
    with np.errstate(all="ignore"):
        assert np.array_equal((1.0 / d).view(np.uint64), inv_bits[ok])
    # anything with a mantissa bit or a sign is refused
    rng = np.random.default_rng(0)
    noisy = (rng.integers(2, 2045, 1000, dtype=np.uint64) << np.uint64(52)) | rng.integers(1, 1 << 52, 1000, dtype=np.uint64)
    assert not _pow2_reciprocal_bits(noisy)[0].any()
    assert not _pow2_reciprocal_bits((np.uint64(1) << np.uint64(63)) | (np.uint64(1023) << np.uint64(52)))[0]   # -1.0


def test_division_by_a_power_of_two_equals_multiplication_by_its_reciprocal_bit_for_bit():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 1 << 64, 200_000, dtype=np.uint64).view(np.float64)          # every kind of double: normal, subnormal, inf, nan
    x = np.concatenate([x, np.array([0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308]), rng.standard_normal(50_000)])
    finite = np.isfinite(x)
    with np.errstate(all="ignore"):
        for e in list(range(-60, 61)) + [-1021, -1000, 1000, 1021]:                  # d = 2^e; the kernel's divisors are 1, 2, 4, 8 and sums of inverse masses
            d = np.ldexp(1.0, e)
            inv = np.ldexp(1.0, -e)
            q, p = x / d, x * inv
            same = (q.view(np.uint64) == p.view(np.uint64)) | (np.isnan(q) & np.isnan(p))
            assert same.all(), (e, x[~same][:3])
            assert np.array_equal(np.isfinite(q[finite]), np.isfinite(p[finite]))


def test_contact_search_square_root_is_only_needed_near_the_threshold():
    """d_find_contacts decides sqrt(d2) < DIAM - EPS from d2 alone outside a 1e-9 band around the threshold's square (csrc/ps2d.cu)"""
    T = 0.5 - 1e-4
    T2 = T * T
    rng = np.random.default_rng(2)
    rel = np.concatenate([rng.uniform(-1e-6, 1e-6, 400_000), rng.uniform(-3e-9, 3e-9, 400_000), np.array([-1e-9, 1e-9, 0.0])])
    d2 = T2 * (1.0 + rel)
    exact = np.sqrt(d2) < T
    fast = np.where(d2 < T2 * (1.0 - 1e-9), True, np.where(d2 > T2 * (1.0 + 1e-9), False, exact))
    assert np.array_equal(fast, exact)
    assert (d2 < T2 * (1.0 - 1e-9)).sum() > 100_000 and (d2 > T2 * (1.0 + 1e-9)).sum() > 100_000   # both shortcuts exercised
