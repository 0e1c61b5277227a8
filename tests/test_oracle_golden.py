"""Pins the CPU oracle against the literal vectors of the reference's own unit tests.

Every test here restates one `#[test]` of /root/reference/src/multivariate_taylor.rs or
src/univariate_taylor.rs (file:line in each docstring); comparisons are bit-exact
(`assert_eq!` on f64 in the reference).  CPU only.
"""
import numpy as np
import pytest

from oracle.oracle import OracleError, TaylorExpansion, TaylorPoly, taylor

T = TaylorPoly


def test_2d_derivative():
    """multivariate_taylor.rs:733-772"""
    t = taylor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0], [9.0, 10.0, 11.0, 12.0], [13.0, 14.0, 15.0, 16.0]])
    assert t.derivative(0, 1) == taylor([[5.0, 6.0, 7.0, 8.0], [18.0, 20.0, 22.0, 24.0], [39.0, 42.0, 45.0, 48.0]])
    assert t.derivative(1, 1) == taylor([[2.0, 6.0, 12.0], [6.0, 14.0, 24.0], [10.0, 22.0, 36.0], [14.0, 30.0, 48.0]])
    assert t.derivative(0, 2) == t.derivative(0, 1).derivative(0, 1)
    assert t.derivative(1, 2) == t.derivative(1, 1).derivative(1, 1)
    assert t.derivative(0, 3) == t.derivative(0, 1).derivative(0, 1).derivative(0, 1)


def test_2d_taylor_expansion_of_coeff():
    """multivariate_taylor.rs:775-803"""
    t = taylor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0], [9.0, 10.0, 11.0, 12.0], [13.0, 14.0, 15.0, 16.0]])
    assert t.taylor_expansion_of_coeff(0, 2) == taylor([[9.0, 10.0, 11.0, 12.0], [39.0, 42.0, 45.0, 48.0]])
    assert t.taylor_expansion_of_coeff(1, 3) == taylor([[4.0], [8.0], [12.0], [16.0]])
    expected = taylor([[11.0, 36.0], [45.0, 144.0]])
    assert t.taylor_expansion_of_coeff(0, 2).taylor_expansion_of_coeff(1, 2) == expected
    assert t.taylor_expansion_of_coeff(1, 2).taylor_expansion_of_coeff(0, 2) == expected


def test_2d_subst_var():
    """multivariate_taylor.rs:806-829"""
    t = taylor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    s = taylor([[10.0, 11.0, 12.0], [13.0, 14.0, 15.0], [16.0, 17.0, 18.0]])
    assert t.subst_var(0, s) == taylor([[741.0, 2436.0, 5353.0], [1872.0, 6163.0, 13516.0], [3487.0, 11452.0, 25030.0]])
    assert t.subst_var(1, s) == taylor([[321.0, 682.0, 1107.0], [1460.0, 3101.0, 5016.0], [4111.0, 8736.0, 14088.0]])
    assert t.subst_var(0, s).subst_var(1, s) != t.subst_var(1, s).subst_var(0, s)


def test_add_mismatched_shapes():
    """multivariate_taylor.rs:885-892"""
    a, b = T.var(0, 1.0, 5), T.var(1, 1.0, 4)
    assert (a + b).extend([5, 4]) == a.extend([5, 4]) + b.extend([5, 4])


def test_sub_mismatched_shapes():
    """multivariate_taylor.rs:940-947"""
    a, b = T.var(0, 1.0, 5), T.var(1, 1.0, 4)
    assert (a - b).extend([5, 4]) == a.extend([5, 4]) - b.extend([5, 4])


def test_mul_mismatched_shapes():
    """multivariate_taylor.rs:1081-1094"""
    a, b = T.var(0, 1.0, 5), T.var(1, 1.0, 4)
    assert (a * b).extend([5, 4]) == a.extend([5, 4]) * b.extend([5, 4])
    c, d = a * a * a, b * b
    assert (c * d).extend([5, 4]) == c.extend([5, 4]) * d.extend([5, 4])


def test_2d_mul():
    """multivariate_taylor.rs:1097-1102"""
    f, g = taylor([[1.0, 2.0], [3.0, 4.0]]), taylor([[5.0, 6.0], [7.0, 8.0]])
    assert f * g == taylor([[5.0, 16.0], [22.0, 60.0]])


def test_2d_mul_const():
    """multivariate_taylor.rs:1105-1127"""
    f, g = taylor([[1.0, 2.0], [3.0, 4.0]]), taylor([[5.0, 6.0], [7.0, 8.0]])
    assert f * g == taylor([[5.0, 16.0], [22.0, 60.0]])
    assert f * T.zero() == T.zero_with([2, 2])
    assert T.zero() * f == T.zero_with([2, 2])
    assert f * T.one() == f
    assert T.one() * f == f
    assert T.from_u32(2) * f == taylor([[2.0, 4.0], [6.0, 8.0]])
    assert f * T.from_u32(2) == taylor([[2.0, 4.0], [6.0, 8.0]])


def test_2d_mul_factor_linear():
    """multivariate_taylor.rs:1130-1160"""
    f = taylor([[1.0, 2.0], [3.0, 4.0]])
    g0 = T.from_u32(2) * T.var_at_zero(0, 2)
    assert g0.extract_linear() == (0.0, 2.0, 0)
    g1 = T.from_u32(3) * T.var_at_zero(1, 2)
    assert g1.extract_linear() == (0.0, 3.0, 1)
    assert f * g0 == taylor([[0.0, 0.0], [2.0, 4.0]])
    assert f * g1 == taylor([[0.0, 3.0], [0.0, 9.0]])
    assert g0 * f == taylor([[0.0, 0.0], [2.0, 4.0]])
    assert g1 * f == taylor([[0.0, 3.0], [0.0, 9.0]])
    assert g0 * g1 == taylor([[0.0, 0.0], [0.0, 6.0]])
    assert g1 * g0 == taylor([[0.0, 0.0], [0.0, 6.0]])
    g0 = taylor([3.0, 2.0])
    assert g0.extract_linear() == (3.0, 2.0, 0)
    g1 = taylor([[3.0, 2.0], [0.0, 0.0]])
    assert g1.extract_linear() == (3.0, 2.0, 1)
    assert f * g0 == taylor([[3.0, 6.0], [11.0, 16.0]])
    assert f * g1 == taylor([[3.0, 8.0], [9.0, 18.0]])
    assert g0 * f == taylor([[3.0, 6.0], [11.0, 16.0]])
    assert g1 * f == taylor([[3.0, 8.0], [9.0, 18.0]])
    assert g0 * g1 == taylor([[9.0, 6.0], [6.0, 4.0]])
    assert g1 * g0 == taylor([[9.0, 6.0], [6.0, 4.0]])


def test_div_mismatched_shapes():
    """multivariate_taylor.rs:1240-1253"""
    a, b = T.var(0, 1.0, 5), T.var(1, 1.0, 4)
    assert (a / b).extend([5, 4]) == a.extend([5, 4]) / b.extend([5, 4])
    c, d = a * a * a, b * b
    assert (c * d).extend([5, 4]) == c.extend([5, 4]) * d.extend([5, 4])


def test_2d_div():
    """multivariate_taylor.rs:1256-1268"""
    f, g = taylor([[1.0, 2.0], [3.0, 4.0]]), taylor([[5.0, 6.0], [7.0, 8.0]])
    r = f / g
    assert r == taylor([[0.2, 0.159_999_999_999_999_98], [0.319_999_999_999_999_95, -0.127_999_999_999_999_9]])
    assert r * g == f


def test_exp_mismatched_shapes():
    """multivariate_taylor.rs:1389-1402"""
    a = T.var(0, 1.0, 5)
    assert a.exp().extend([5, 4]) == a.extend([5, 4]).exp()
    c = a * a * a
    assert c.exp().extend([5, 4]) == c.extend([5, 4]).exp()
    a = taylor([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0]], [5, 4])
    assert a.exp().extend([5, 4]) == a.extend([5, 4]).exp()
    c = a * a * a
    assert c.exp().extend([5, 4]) == c.extend([5, 4]).exp()


def test_2d_exp():
    """multivariate_taylor.rs:1406-1437"""
    assert T.zero().exp() == T.one()
    f, g = taylor([[1.0, 2.0], [3.0, 4.0]]), taylor([[5.0, 6.0], [7.0, 8.0]])
    assert f.exp() == taylor([[2.718_281_828_459_045, 5.436_563_656_918_09],
                              [8.154_845_485_377_136, 27.182_818_284_590_454]])
    assert f.exp() * (-f).exp() == taylor([[1.0, 0.0], [0.0, 0.0]])
    assert f.exp() * g.exp() == taylor([[403.428_793_492_735_1, 3_227.430_347_941_881],
                                        [4_034.287_934_927_350_8, 37_115.449_001_331_624]])
    assert (f + g).exp() == taylor([[403.428_793_492_735_1, 3_227.430_347_941_881],
                                    [4_034.287_934_927_351, 37_115.449_001_331_63]])


def test_log_mismatched_shapes():
    """multivariate_taylor.rs:1440-1453"""
    a = T.var(0, 1.0, 5)
    assert a.log().extend([5, 4]) == a.extend([5, 4]).log()
    c = a * a * a
    assert c.log().extend([5, 4]) == c.extend([5, 4]).log()
    a = taylor([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0]], [5, 4])
    assert a.log().extend([5, 4]) == a.extend([5, 4]).log()
    c = a * a * a
    assert c.log().extend([5, 4]) == c.extend([5, 4]).log()


def test_2d_log():
    """multivariate_taylor.rs:1456-1513"""
    assert T.one().log() == T.zero()
    xp1 = T.var(0, 1.0, 5)
    assert xp1.log() == taylor([0.0, 1.0, -0.5, 0.333_333_333_333_333_3, -0.25])
    e = taylor([1.0, 2.0, 3.0])
    assert e.log() == taylor([0.0, 2.0, 1.0])
    assert e.log().exp() == e
    f = taylor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    g = taylor([[5.0, 6.0, 7.0], [7.0, 8.0, 9.0], [9.0, 10.0, 11.0]])
    assert f.log() == taylor([[0.0, 2.0, 1.0], [4.0, -3.0, 0.0], [-1.0, 6.0, -4.5]])
    assert f.log().exp() == taylor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    assert f.exp().log() == taylor([[1.0, 2.0, 3.000_000_000_000_001],
                                    [4.0, 4.999_999_999_999_999, 6.000_000_000_000_007],
                                    [6.999_999_999_999_999, 8.000_000_000_000_002, 8.999_999_999_999_991]])
    assert f.log() + (T.one() / f).log() == taylor([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    assert f.log() + g.log() == taylor([[1.609_437_912_434_100_3, 3.2, 1.680_000_000_000_000_2],
                                        [5.4, -3.079_999_999_999_999_6, -0.064_000_000_000_000_06],
                                        [-0.179_999_999_999_999_94, 5.952, -4.5416]])
    assert (f * g).log() == taylor([[1.609_437_912_434_100_3, 3.2, 1.679_999_999_999_999_7],
                                    [5.4, -3.080_000_000_000_001, -0.063_999_999_999_998_64],
                                    [-0.180_000_000_000_001_13, 5.952_000_000_000_003, -4.541_600_000_000_002_5]])


# ---- univariate (src/univariate_taylor.rs) -------------------------------------------------
E = TaylorExpansion


def test_uni_taylor_expansion_of_coeff():
    """univariate_taylor.rs:119-132"""
    x = E.var(2.0, 4)
    f = (x * x + E.one()).exp()
    g = f.taylor_expansion_of_coeff(2)
    assert g == E.from_coefficients([1_335.718_431_923_189_4, 6_530.179_000_513_37, 17_067.513_296_796_307])


def test_uni_subst():
    """univariate_taylor.rs:135-148"""
    x, y = E.var(1.0, 2), E.var(2.0, 2)
    assert x.subst(y) == E.from_coefficients([3.0, 1.0, 0.0])
    assert (x * x).subst(y * y) == E.from_coefficients([25.0, 40.0, 26.0])


def test_uni_e_x_squared_1():
    """univariate_taylor.rs:480-497"""
    x = E.var(0.0, 9)
    r = (x * x - E.one()).exp()
    assert r == E.from_coefficients([0.367_879_441_171_442_33, 0.0, 0.367_879_441_171_442_33, 0.0,
                                     0.183_939_720_585_721_17, 0.0, 0.061_313_240_195_240_39, 0.0,
                                     0.015_328_310_048_810_098, 0.0])


def test_uni_division():
    """univariate_taylor.rs:500-532"""
    x = E.var(0.0, 9)
    assert x / (x - E.one()) == E.from_coefficients([0.0] + [-1.0] * 9)
    assert x / x.exp() == E.from_coefficients([0.0, 1.0, -1.0, 0.5, -0.166_666_666_666_666_63,
                                               0.041_666_666_666_666_63, -0.008_333_333_333_333_31,
                                               0.001_388_888_888_888_877, -0.000_198_412_698_412_693_37,
                                               0.000_024_801_587_301_585_587])


def test_uni_division_constant():
    """univariate_taylor.rs:535-556"""
    x = E.var(0.0, 9)
    assert E.one() / (x - E.one()) == E.from_coefficients([-1.0] * 10)
    assert E.one() / x.exp() == E.from_coefficients([1.0, -1.0, 0.5, -0.166_666_666_666_666_63,
                                                     0.041_666_666_666_666_63, -0.008_333_333_333_333_31,
                                                     0.001_388_888_888_888_877, -0.000_198_412_698_412_693_37,
                                                     0.000_024_801_587_301_585_587, -2.755_731_922_398_079_3e-6])


def test_uni_log():
    """univariate_taylor.rs:559-578"""
    x = E.var(1.0, 4)
    assert x.log() == E.from_coefficients([0.0, 1.0, -0.5, 0.333_333_333_333_333_3, -0.25])
    assert x.exp().log() == x
    assert x.log().exp() == x
    e = E.from_coefficients([1.0, 2.0, 3.0])
    assert e.log() == E.from_coefficients([0.0, 2.0, 1.0])
    assert e.log().exp() == e


# ---- semantics listed in SURVEY.md Appendix A (not literal reference tests) ----------------
def test_zero_with_is_not_dense_zero():
    """Appendix A.1: derived PartialEq compares stored shape too (multivariate_taylor.rs:10, :1109-1116)."""
    assert T.zero_with([2, 2]) != taylor([[0.0, 0.0], [0.0, 0.0]])
    assert T.zero_with([2, 2]).array_shape() == (1, 1)


def test_shift_down_doc_example():
    """multivariate_taylor.rs:511-513: shifting 2 + 3v + v^2 down by 1 yields 5 + v."""
    assert taylor([2.0, 3.0, 1.0]).shift_down(0, 1) == taylor([5.0, 1.0], [2])


def test_coefficient_bounds():
    """multivariate_taylor.rs:314-339 (assert on index >= degrees, zero beyond stored shape)."""
    t = taylor([[1.0, 2.0]], [3, 4])
    assert t.coefficient([0, 1]) == 2.0
    assert t.coefficient([2, 3]) == 0.0
    with pytest.raises(OracleError):
        t.coefficient([3, 0])


def test_interval_encloses_point_arithmetic():
    """--bounds enclosure (interval.rs): Interval<F64> result contains the F64 result."""
    rng = np.random.default_rng(7)
    a, b = rng.uniform(0.5, 2.0, (4, 4)), rng.uniform(0.5, 2.0, (4, 4))
    fa, fb = taylor(a), taylor(b)
    ia, ib = taylor(np.stack([a, a], -1), kind="iv"), taylor(np.stack([b, b], -1), kind="iv")
    for op in (lambda x, y: x * y, lambda x, y: x / y, lambda x, y: (x * y).log(), lambda x, y: (x - y).exp()):
        p, q = op(fa, fb).array(), op(ia, ib).array()
        assert np.all(q[..., 0] <= p) and np.all(p <= q[..., 1])
