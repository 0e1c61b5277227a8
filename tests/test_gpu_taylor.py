"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(a) the literal vectors of the reference's unit tests and (b) the CPU oracle on seeded random inputs.

Bars: integer work (stored shapes, degrees_p1) bit-exact everywhere; f64 coefficients bit-exact for
the element-wise/gather family, the reference-order product kernel and the single-axis division;
within 1e-12 relative (north_star) wherever the summation order or libm differs (exp/log, the
register-tiled product, general division).
"""
import os

import numpy as np
import pytest

from helpers import RTOL, assert_close, assert_meta_equal, assert_same, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import genfer_b200
    ctx = genfer_b200.Context(0)
    genfer_b200.set_default_context(ctx)
    yield genfer_b200
    genfer_b200.set_default_context(None)
    ctx.close()


def O():
    from oracle import oracle
    return oracle


def both(G, a, degrees=None):
    a = np.asarray(a, dtype=np.float64)
    d = a.shape if degrees is None else degrees
    return G.TaylorPoly.new(a, d), O().TaylorPoly.new(a, d)


# ---------------------------------------------------------------------------------------------
# the reference's own unit tests, on the device
# ---------------------------------------------------------------------------------------------
def test_ref_2d_derivative(G):
    """multivariate_taylor.rs:733-772"""
    t = G.taylor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0], [9.0, 10.0, 11.0, 12.0], [13.0, 14.0, 15.0, 16.0]])
    assert t.derivative(0, 1) == G.taylor([[5.0, 6.0, 7.0, 8.0], [18.0, 20.0, 22.0, 24.0], [39.0, 42.0, 45.0, 48.0]])
    assert t.derivative(1, 1) == G.taylor([[2.0, 6.0, 12.0], [6.0, 14.0, 24.0], [10.0, 22.0, 36.0], [14.0, 30.0, 48.0]])
    assert t.derivative(0, 2) == t.derivative(0, 1).derivative(0, 1)
    assert t.derivative(1, 2) == t.derivative(1, 1).derivative(1, 1)
    assert t.derivative(0, 3) == t.derivative(0, 1).derivative(0, 1).derivative(0, 1)


def test_ref_2d_taylor_expansion_of_coeff(G):
    """multivariate_taylor.rs:775-803"""
    t = G.taylor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0], [9.0, 10.0, 11.0, 12.0], [13.0, 14.0, 15.0, 16.0]])
    assert t.taylor_expansion_of_coeff(0, 2) == G.taylor([[9.0, 10.0, 11.0, 12.0], [39.0, 42.0, 45.0, 48.0]])
    assert t.taylor_expansion_of_coeff(1, 3) == G.taylor([[4.0], [8.0], [12.0], [16.0]])
    expected = G.taylor([[11.0, 36.0], [45.0, 144.0]])
    assert t.taylor_expansion_of_coeff(0, 2).taylor_expansion_of_coeff(1, 2) == expected
    assert t.taylor_expansion_of_coeff(1, 2).taylor_expansion_of_coeff(0, 2) == expected


def test_ref_2d_subst_var(G):
    """multivariate_taylor.rs:806-829"""
    t = G.taylor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    s = G.taylor([[10.0, 11.0, 12.0], [13.0, 14.0, 15.0], [16.0, 17.0, 18.0]])
    assert t.subst_var(0, s) == G.taylor([[741.0, 2436.0, 5353.0], [1872.0, 6163.0, 13516.0], [3487.0, 11452.0, 25030.0]])
    assert t.subst_var(1, s) == G.taylor([[321.0, 682.0, 1107.0], [1460.0, 3101.0, 5016.0], [4111.0, 8736.0, 14088.0]])
    assert t.subst_var(0, s).subst_var(1, s) != t.subst_var(1, s).subst_var(0, s)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_ref_mismatched_shapes(G, op):
    """multivariate_taylor.rs:885-892, :940-947, :1081-1094, :1240-1253"""
    T = G.TaylorPoly
    f = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y, "div": lambda x, y: x / y}[op]
    a, b = T.var(0, 1.0, 5), T.var(1, 1.0, 4)
    assert f(a, b).extend([5, 4]) == f(a.extend([5, 4]), b.extend([5, 4]))
    c, d = a * a * a, b * b
    assert (c * d).extend([5, 4]) == c.extend([5, 4]) * d.extend([5, 4])


def test_ref_2d_mul_and_const(G):
    """multivariate_taylor.rs:1097-1127"""
    T = G.TaylorPoly
    f, g = G.taylor([[1.0, 2.0], [3.0, 4.0]]), G.taylor([[5.0, 6.0], [7.0, 8.0]])
    assert f * g == G.taylor([[5.0, 16.0], [22.0, 60.0]])
    assert f * T.zero() == T.zero_with([2, 2])
    assert T.zero() * f == T.zero_with([2, 2])
    assert f * T.one() == f
    assert T.one() * f == f
    assert T.from_u32(2) * f == G.taylor([[2.0, 4.0], [6.0, 8.0]])
    assert f * T.from_u32(2) == G.taylor([[2.0, 4.0], [6.0, 8.0]])


def test_ref_2d_mul_factor_linear(G):
    """multivariate_taylor.rs:1130-1160"""
    T, taylor = G.TaylorPoly, G.taylor
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


def test_ref_2d_div(G):
    """multivariate_taylor.rs:1256-1268 (general N-D divisor: wavefront kernel, tolerance)"""
    f, g = G.taylor([[1.0, 2.0], [3.0, 4.0]]), G.taylor([[5.0, 6.0], [7.0, 8.0]])
    r = f / g
    ref = np.array([[0.2, 0.159_999_999_999_999_98], [0.319_999_999_999_999_95, -0.127_999_999_999_999_9]])
    np.testing.assert_allclose(r.array(), ref, rtol=RTOL)
    np.testing.assert_allclose((r * g).array(), f.array(), rtol=RTOL)


def test_ref_exp(G):
    """multivariate_taylor.rs:1389-1437 (device exp() differs from glibc by <= 1 ulp: tolerance)"""
    T, taylor = G.TaylorPoly, G.taylor
    assert T.zero().exp() == T.one()
    a = T.var(0, 1.0, 5)
    assert a.exp().extend([5, 4]) == a.extend([5, 4]).exp()
    c = a * a * a
    assert c.exp().extend([5, 4]) == c.extend([5, 4]).exp()
    f, g = taylor([[1.0, 2.0], [3.0, 4.0]]), taylor([[5.0, 6.0], [7.0, 8.0]])
    np.testing.assert_allclose(f.exp().array(), [[2.718_281_828_459_045, 5.436_563_656_918_09],
                                                 [8.154_845_485_377_136, 27.182_818_284_590_454]], rtol=RTOL)
    np.testing.assert_allclose((f.exp() * (-f).exp()).array(), [[1.0, 0.0], [0.0, 0.0]], rtol=RTOL, atol=1e-14)
    np.testing.assert_allclose((f + g).exp().array(), [[403.428_793_492_735_1, 3_227.430_347_941_881],
                                                       [4_034.287_934_927_351, 37_115.449_001_331_63]], rtol=RTOL)
    np.testing.assert_allclose((f.exp() * g.exp()).array(), (f + g).exp().array(), rtol=RTOL)


def test_ref_log(G):
    """multivariate_taylor.rs:1440-1513"""
    T, taylor = G.TaylorPoly, G.taylor
    assert T.one().log() == T.zero()
    xp1 = T.var(0, 1.0, 5)
    assert xp1.log() == taylor([0.0, 1.0, -0.5, 0.333_333_333_333_333_3, -0.25])
    e = taylor([1.0, 2.0, 3.0])
    assert e.log() == taylor([0.0, 2.0, 1.0])
    assert e.log().exp() == e
    a = T.var(0, 1.0, 5)
    assert a.log().extend([5, 4]) == a.extend([5, 4]).log()
    f = taylor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    g = taylor([[5.0, 6.0, 7.0], [7.0, 8.0, 9.0], [9.0, 10.0, 11.0]])
    np.testing.assert_allclose(f.log().array(), [[0.0, 2.0, 1.0], [4.0, -3.0, 0.0], [-1.0, 6.0, -4.5]], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(f.log().exp().array(), f.array(), rtol=RTOL)
    np.testing.assert_allclose(f.exp().log().array(), f.array(), rtol=1e-11)   # a round trip through two series, not a parity check
    np.testing.assert_allclose((f * g).log().array(), (f.log() + g.log()).array(), rtol=1e-11, atol=1e-12)
    assert_close(f.log(), O().taylor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]).log(), rtol=1e-12)   # parity proper


# ---------------------------------------------------------------------------------------------
# seeded random parity against the oracle, every operator
# ---------------------------------------------------------------------------------------------
SHAPES = [((5,), (5,)), ((3, 4), (6, 5)), ((4, 1, 3), (4, 2, 5)), ((2, 3, 2, 3), (3, 3, 3, 3)), ((7, 6), (7, 6))]


def rand(rng, shape, positive=False):
    a = rng.uniform(0.5, 1.5, shape) if positive else rng.standard_normal(shape)
    return a


@pytest.mark.parametrize("sa,da", SHAPES)
@pytest.mark.parametrize("sb,db", SHAPES)
def test_binary_ops_match_oracle(G, sa, da, sb, db):
    rng = np.random.default_rng(hash((sa, sb)) % 2**32)
    a, b = rand(rng, sa), rand(rng, sb, positive=True)
    ga, oa = both(G, a, da)
    gb, ob = both(G, b, db)
    assert_same(ga + gb, oa + ob)
    assert_same(ga - gb, oa - ob)
    assert_same(gb - ga, ob - oa)
    assert_same(ga * gb, oa * ob)  # reference-order kernel: bit-exact
    assert_close(ga / gb, oa / ob, rtol=1e-12)  # general divisor: the reference's recurrence, device-resident (kernels_wave.cu)
    assert_same(-ga, -oa)


def test_scalar_fast_paths(G):
    """Add/Sub scalar paths (:862-869, :919-926) incl. signed zeros; Mul/Div by constants."""
    a = np.array([[1.5, -0.0, 2.0], [0.0, -3.0, 4.0]])
    ga, oa = both(G, a, (4, 5))
    for s in (0.0, 2.5, -1.0, 1.0):
        gs, os_ = G.TaylorPoly.from_scalar(s), O().TaylorPoly.from_scalar(s)
        assert_same(ga + gs, oa + os_)
        assert_same(gs + ga, os_ + oa)
        assert_same(ga - gs, oa - os_)
        assert_same(gs - ga, os_ - oa)
        assert_same(ga * gs, oa * os_)
        assert_same(gs * ga, os_ * oa)
        if s != 0.0:
            assert_same(ga / gs, oa / os_)


@pytest.mark.parametrize("shape,deg", [((6,), (8,)), ((4, 5), (6, 5)), ((3, 4, 5), (3, 6, 5)), ((2, 2, 3, 4), (4, 4, 4, 4))])
def test_gathers_match_oracle(G, shape, deg):
    rng = np.random.default_rng(1234)
    a = rand(rng, shape)
    g, o = both(G, a, deg)
    for v in range(len(shape)):
        for n in range(0, deg[v]):
            assert_same(g.derivative(v, n), o.derivative(v, n))
            assert_same(g.taylor_expansion_of_coeff(v, n), o.taylor_expansion_of_coeff(v, n))
            assert_same(g.shift_down(v, n), o.shift_down(v, n))
            assert_same(g.coefficients_of_term(v, n), o.coefficients_of_term(v, n))
            assert_same(g.taylor_polynomial(v, n), o.taylor_polynomial(v, n))
        assert_same(g.taylor_polynomial_terms(v, [0, 2]), o.taylor_polynomial_terms(v, [0, 2]))
        assert_same(g.taylor_polynomial_terms(v, [1]), o.taylor_polynomial_terms(v, [1]))
    assert_same(g.truncate_to_degree_p1(3), o.truncate_to_degree_p1(3))
    assert_same(g.truncate_to_degree_p1(2), o.truncate_to_degree_p1(2))
    assert_same(g.remove_last_variable(), o.remove_last_variable())
    assert_same(g.extend_to_dim(len(shape) + 2, 7), o.extend_to_dim(len(shape) + 2, 7))
    # beyond-ndim variable conventions (:343-348, :460-465)
    assert_same(g.coefficients_of_term(len(shape) + 1, 0), o.coefficients_of_term(len(shape) + 1, 0))
    assert_same(g.coefficients_of_term(len(shape) + 1, 1), o.coefficients_of_term(len(shape) + 1, 1))


def test_shift_down_long_lanes(G):
    """shift_down along the last axis uses ndarray's 8-way unrolled fold order (restated in the oracle)."""
    rng = np.random.default_rng(5)
    a = rng.standard_normal((5, 37))
    g, o = both(G, a)
    for n in (1, 7, 8, 9, 20, 35, 36):
        assert_same(g.shift_down(1, n), o.shift_down(1, n))
        assert_same(g.shift_down(0, min(n, 4)), o.shift_down(0, min(n, 4)))


@pytest.mark.parametrize("shape", [(40, 30, 16), (1100, 9), (33, 40, 24), (2048, 64), (1500, 2), (1200, 13)])
def test_shift_down_short_lanes_tiled(G, shape):
    """>= 1024 contiguous short lanes take the shared-memory tiled reduction kernel; same fold order, bit-exact."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal(shape)
    g, o = both(G, a)
    v = len(shape) - 1
    for n in sorted({0, 1, 7, 8, 9, shape[-1] - 2, shape[-1] - 1} & set(range(0, shape[-1]))):
        assert_same(g.shift_down(v, n), o.shift_down(v, n))


def test_readers_and_panics(G):
    a = np.arange(12, dtype=np.float64).reshape(3, 4) + 1
    g, o = both(G, a, (5, 6))
    assert g.constant_term() == o.constant_term() == 1.0
    assert g.coefficient([2, 3]) == o.coefficient([2, 3]) == 12.0
    assert g.coefficient([4, 5]) == o.coefficient([4, 5]) == 0.0
    with pytest.raises(G.TaylorPanic):
        g.coefficient([5, 0])  # index >= degrees_p1 -> the reference asserts (:318-322)
    with pytest.raises(G.TaylorPanic):
        g.derivative(0, 5)     # n >= len_of(v) (:459)
    with pytest.raises(G.TaylorPanic):
        g.shift_down(2, 0)     # v >= num_vars (:516)
    np.testing.assert_array_equal(g.gather_axis(0, 5), [1.0, 5.0, 9.0, 0.0, 0.0])
    np.testing.assert_array_equal(g.gather_axis(1, 6), [1.0, 2.0, 3.0, 4.0, 0.0, 0.0])
    assert g.evaluate_all_one() == o.evaluate_all_one() == 78.0
    assert g.extract_constant() is None and G.TaylorPoly.from_scalar(3.0).extract_constant() == 3.0
    assert not g.is_zero() and G.TaylorPoly.zero().is_zero() and G.TaylorPoly.one().is_one()


def test_extract_linear_matches_oracle(G):
    cases = [np.array([3.0, 2.0]), np.array([[3.0, 2.0], [0.0, 0.0]]), np.array([[3.0, 0.0], [2.0, 0.0]]),
             np.array([[3.0, 2.0], [1.0, 0.0]]), np.array([[0.0, 0.0], [0.0, 0.0]]), np.array([1.0, 2.0, 3.0]),
             np.array([1.0, 2.0, 0.0]), np.zeros((2, 3, 2)), np.array([[[1.0, 0.0]], [[5.0, 0.0]]])]
    for a in cases:
        g, o = both(G, a)
        assert g.extract_linear() == o.extract_linear(), a


@pytest.mark.parametrize("shape,deg", [((4,), (9,)), ((3, 3), (5, 4)), ((2, 3, 2), (4, 4, 3)), ((1, 4), (3, 6))])
def test_exp_log_pow_match_oracle(G, shape, deg):
    rng = np.random.default_rng(99)
    a = rng.uniform(0.5, 1.5, shape)
    g, o = both(G, a, deg)
    assert_close(g.exp(), o.exp(), rtol=1e-12)
    assert_close(g.log(), o.log(), rtol=1e-12)
    for e in (0, 1, 2, 5):
        assert_close(g.pow(e), o.pow(e), rtol=1e-12)


@pytest.mark.parametrize("n,deg", [(1, 12), (2, 7), (3, 5)])
def test_subst_var_matches_oracle(G, n, deg):
    rng = np.random.default_rng(7 + n)
    a = rng.uniform(0.1, 1.0, (deg,) * n)
    s = rng.uniform(0.1, 1.0, (deg,) * n)
    g, o = both(G, a)
    gs, os_ = both(G, s)
    for v in range(n):
        assert_same(g.subst_var(v, gs), o.subst_var(v, os_))             # Horner of ordered products: bit-exact
        lin_g, lin_o = G.TaylorPoly.from_scalar(0.3) * G.TaylorPoly.var_at_zero(v, deg), \
            O().TaylorPoly.from_scalar(0.3) * O().TaylorPoly.var_at_zero(v, deg)
        assert_same(g.subst_var(v, lin_g), o.subst_var(v, lin_o))        # m^i scaling path (:555-568)
        assert_same(g.subst_var(v, G.TaylorPoly.zero()), o.subst_var(v, O().TaylorPoly.zero()))


@pytest.mark.parametrize("shape,ylen,axis", [((40,), 3, 0), ((300,), 2, 0), ((6, 9), 4, 1), ((9, 6), 5, 0),
                                              ((3, 8, 4), 3, 1), ((70000, 5), 2, 1)])
def test_div_single_axis_bit_exact(G, shape, ylen, axis):
    """p / (1 - q v)-style divisors (semantics/gf.rs:465-519): single-axis recurrence, bit-exact."""
    rng = np.random.default_rng(31)
    x = rng.uniform(0.0, 1.0, shape)
    yshape = [1] * len(shape)
    yshape[axis] = ylen
    y = rng.uniform(0.2, 1.0, yshape)
    deg = list(shape)
    deg[axis] += 3  # the quotient is longer than the numerator along the divisor's axis
    gx, ox = both(G, x, deg)
    gy, oy = both(G, y, deg)
    assert_same(gx / gy, ox / oy)


# ---------------------------------------------------------------------------------------------
# univariate TaylorExpansion (src/univariate_taylor.rs tests :119-148, :480-578)
# ---------------------------------------------------------------------------------------------
def test_univariate_reference_tests(G):
    E = G.TaylorExpansion
    x = E.var(2.0, 4)
    g = (x * x + E.one()).exp().taylor_expansion_of_coeff(2)
    np.testing.assert_allclose(g.coeffs(), [1_335.718_431_923_189_4, 6_530.179_000_513_37, 17_067.513_296_796_307], rtol=RTOL)
    x, y = E.var(1.0, 2), E.var(2.0, 2)
    assert x.subst(y) == E.from_coefficients([3.0, 1.0, 0.0])
    assert (x * x).subst(y * y) == E.from_coefficients([25.0, 40.0, 26.0])
    x = E.var(0.0, 9)
    assert x / (x - E.one()) == E.from_coefficients([0.0] + [-1.0] * 9)
    assert E.one() / (x - E.one()) == E.from_coefficients([-1.0] * 10)
    np.testing.assert_allclose((E.one() / x.exp()).coeffs(),
                               [1.0, -1.0, 0.5, -0.166_666_666_666_666_63, 0.041_666_666_666_666_63,
                                -0.008_333_333_333_333_31, 0.001_388_888_888_888_877, -0.000_198_412_698_412_693_37,
                                0.000_024_801_587_301_585_587, -2.755_731_922_398_079_3e-6], rtol=1e-12)
    x = E.var(1.0, 4)
    assert x.log() == E.from_coefficients([0.0, 1.0, -0.5, 0.333_333_333_333_333_3, -0.25])
    e = E.from_coefficients([1.0, 2.0, 3.0])
    assert e.log() == E.from_coefficients([0.0, 2.0, 1.0])
    assert e.log().exp() == e


@pytest.mark.parametrize("n", [1, 5, 64, 300, 1000])
def test_univariate_matches_oracle(G, n):
    rng = np.random.default_rng(n)
    a, b = rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 1.5, n)
    E, OE = G.TaylorExpansion, O().TaylorExpansion
    ga, gb, oa, ob = E.from_coefficients(a), E.from_coefficients(b), OE.from_coefficients(a), OE.from_coefficients(b)
    gc, oc = E.constant(1.25), OE.constant(1.25)
    same = lambda g, o: np.testing.assert_array_equal(g.coeffs().view(np.uint64), o.coeffs().view(np.uint64))
    for f in (lambda x, y: x + y, lambda x, y: x - y, lambda x, y: x * y, lambda x, y: x / y):
        same(f(ga, gb), f(oa, ob))
        same(f(ga, gc), f(oa, oc))
        same(f(gc, ga), f(oc, oa))
        same(f(gc, gc), f(oc, oc))
    same(-ga, -oa)
    same(ga.pow(3), oa.pow(3))
    np.testing.assert_allclose(ga.exp().coeffs(), oa.exp().coeffs(), rtol=1e-12)
    np.testing.assert_allclose(ga.log().coeffs(), oa.log().coeffs(), rtol=1e-12, atol=0.0)
    assert ga.coeff(n - 1) == oa.coeff(n - 1)
    assert ga.derivative(min(n - 1, 5)) == oa.derivative(min(n - 1, 5))
    same(ga.taylor_expansion_of_coeff(n // 2), oa.taylor_expansion_of_coeff(n // 2))


# ---------------------------------------------------------------------------------------------
# fused Horner loop (kernels_horner.cu): subst_var with a substitution of <= 32 coefficients, the inner loop of the
# compound distributions / assignments of real programs (multivariate_taylor.rs:569-579) -- bit-identical to the
# reference's per-step Mul + Add, in a handful of launches instead of three per slice
# ---------------------------------------------------------------------------------------------
HORNER_CASES = [  # (self shape, degrees, substituted axis, subst shape)
    ((9, 8, 7), (12, 11, 10), 2, (2, 1, 2)), ((9, 8, 7), (12, 11, 10), 0, (2, 2, 1)), ((20, 6), (24, 9), 0, (2, 2)),
    ((6, 20), (9, 24), 1, (1, 3)), ((6, 20), (9, 24), 1, (3, 1)), ((30,), (40,), 0, (3,)), ((5, 4, 3, 6), (6, 6, 6, 6), 3, (2, 1, 1, 2)),
    ((40, 50, 45), (60, 60, 60), 2, (2, 1, 2)), ((7, 5), (7, 5), 1, (4, 5)), ((12, 3), (30, 30), 1, (2, 3)),
    ((33, 1, 9), (40, 5, 12), 2, (1, 2, 2)),
    # HBM / L2-sized: the row-staged bulk-copy (TMA) variant -- odd row lengths (shifted, odd-tailed rows), 2 and 4 groups
    ((70, 60, 80), (91, 85, 97), 2, (2, 1, 2)), ((64, 70, 66), (90, 90, 90), 0, (2, 2, 1)), ((50, 300), (400, 700), 0, (2, 2)),
    ((30, 20, 25, 40), (33, 33, 33, 47), 1, (2, 1, 2, 2))]


@pytest.mark.parametrize("shape,deg,v,sshape", HORNER_CASES)
def test_fused_horner_is_bit_exact(G, shape, deg, v, sshape):
    rng = np.random.default_rng(sum(shape) + v)
    a, sub = rng.standard_normal(shape), rng.standard_normal(sshape)
    g, o = both(G, a, deg)
    gs, os_ = both(G, sub, deg)
    ctx = g.ctx
    l0 = ctx.launch_count
    got = g.subst_var(v, gs)
    launches = ctx.launch_count - l0
    assert_same(got, o.subst_var(v, os_))
    assert launches <= 12, launches            # a few per-operator steps until res is non-scalar, then ONE kernel
    # A/B: the per-operator path and the other two fused kernels (row-staged bulk copies; one thread per coefficient) give the same bits
    try:
        for mode in (1 + 2048, 1 + 65536, 1 + 65536 + 16384):
            ctx.set_fast_mul(mode)
            assert_same(g.subst_var(v, gs), o.subst_var(v, os_))
    finally:
        ctx.set_fast_mul(1)


def test_row_walking_horner_kernel_on_every_case():
    """k_horner_direct is the default from 2^17 coefficients with rows of at least 48; here every Horner case that has such rows
    is forced onto it (GTP_DIRECT_MIN=1, read when the context is created), including short first steps (several rows per warp)."""
    import genfer_b200
    from oracle import oracle as O
    old = os.environ.get("GTP_DIRECT_MIN")
    os.environ["GTP_DIRECT_MIN"] = "1"
    try:
        ctx = genfer_b200.Context(0)
    finally:
        if old is None:
            del os.environ["GTP_DIRECT_MIN"]
        else:
            os.environ["GTP_DIRECT_MIN"] = old
    try:
        cases = HORNER_CASES + [((60, 3, 50), (70, 4, 64), 0, (2, 2, 2)), ((100, 49), (120, 80), 1, (3, 2)), ((4, 200), (6, 260), 0, (2, 3))]
        for shape, deg, v, sshape in cases:
            rng = np.random.default_rng(sum(shape) * 7 + v)
            a, sub = rng.standard_normal(shape), rng.standard_normal(sshape)
            g, gs = genfer_b200.TaylorPoly.new(a, deg, ctx), genfer_b200.TaylorPoly.new(sub, deg, ctx)
            o, os_ = O.TaylorPoly.new(a, deg), O.TaylorPoly.new(sub, deg)
            assert_same(g.subst_var(v, gs), o.subst_var(v, os_))
        # signed zeros and exact cancellations go through the branch-free arithmetic (absent terms add +-0): still the same bits
        rng = np.random.default_rng(99)
        a = rng.integers(-2, 3, size=(40, 6, 70)).astype(np.float64)
        a[a == 0] = np.where(rng.random(np.count_nonzero(a == 0)) < 0.5, -0.0, 0.0)
        sub = np.array([[[-0.0, 1.0]], [[-1.0, 0.0]]])
        deg = (50, 6, 90)
        g, gs = genfer_b200.TaylorPoly.new(a, deg, ctx), genfer_b200.TaylorPoly.new(sub, deg, ctx)
        o, os_ = O.TaylorPoly.new(a, deg), O.TaylorPoly.new(sub, deg)
        for v in (0, 2):
            assert_same(g.subst_var(v, gs), o.subst_var(v, os_))
        # a non-finite substitution coefficient takes the predicated form (0 * inf must not reach coefficients the reference
        # never multiplies): same NaN positions, same bits everywhere else
        sub2 = np.array([[[0.5, np.inf]], [[-1.0, 2.0]]])
        gs2, os2 = genfer_b200.TaylorPoly.new(sub2, deg, ctx), O.TaylorPoly.new(sub2, deg)
        ga, oa = g.subst_var(2, gs2).array(), o.subst_var(2, os2).array()
        assert ga.shape == oa.shape and np.array_equal(np.isnan(ga), np.isnan(oa))
        ok = ~np.isnan(oa)
        assert np.array_equal(ga[ok].view(np.uint64), oa[ok].view(np.uint64))
    finally:
        ctx.close()


def test_fused_horner_with_zero_slices_and_scalar_slices(G):
    """Top slices that are exactly zero keep res a scalar zero for a while (Mul's is_zero path, :1021-1023); a 1-d self
    has scalar slices (Add's scalar path, :862-865)."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((6, 9))
    a[:, 6:] = 0.0
    sub = rng.standard_normal((2, 2))
    g, o = both(G, a, (10, 12))
    gs, os_ = both(G, sub, (10, 12))
    assert_same(g.subst_var(1, gs), o.subst_var(1, os_))
    b = rng.standard_normal((25,))
    b[-3:] = 0.0
    g, o = both(G, b, (30,))
    gs, os_ = both(G, np.array([0.25, -1.5, 0.75]), (30,))
    assert_same(g.subst_var(0, gs), o.subst_var(0, os_))
    gs2, os2 = both(G, rng.standard_normal((1, 3)), (30, 8))       # substitution with more variables than self
    assert_same(g.subst_var(0, gs2), o.subst_var(0, os2))


# ---------------------------------------------------------------------------------------------
# device-resident recurrences (kernels_wave.cu): the reference's div / exp / log recurrences (:1162-1192, :1285-1386)
# as ONE cooperative kernel per call -- 1e-12 against the oracle at sizes where the old host loops needed hundreds of
# launches (and the old reciprocal-series division lost four digits), and at most 4 launches per call
# ---------------------------------------------------------------------------------------------
WAVE_SHAPES = [((32, 32, 32), None), ((16, 16, 16, 16), None), ((12, 40), None), ((40, 12), None), ((5, 6, 7), None),
               ((3, 70), None), ((9, 9, 9), (12, 10, 11)), ((2, 3, 2, 3, 2), (3, 4, 3, 4, 3))]


@pytest.mark.parametrize("shape,deg", WAVE_SHAPES)
def test_device_resident_recurrences_match_oracle(G, shape, deg):
    rng = np.random.default_rng(sum(shape))
    a, b = rng.uniform(0.5, 1.5, shape), rng.uniform(0.5, 1.5, shape)
    x = rng.standard_normal(shape)
    ga, oa = both(G, a, deg)
    gb, ob = both(G, b, deg)
    gx, ox = both(G, x, deg)
    ctx = ga.ctx
    for name, fg, fo in (("exp", lambda: ga.exp(), lambda: oa.exp()), ("log", lambda: ga.log(), lambda: oa.log()),
                         ("div", lambda: gx / gb, lambda: ox / ob)):
        l0 = ctx.launch_count
        got = fg()
        launches = ctx.launch_count - l0
        assert_close(got, fo(), rtol=1e-12)
        assert launches <= 4, (name, shape, launches)


@pytest.mark.parametrize("shape,deg", [((4, 5), (7, 6)), ((6, 6), None), ((3, 2, 4), (4, 4, 5)), ((2, 20), (3, 45)), ((3, 3, 2, 2), None)])
def test_small_nd_exp_is_bit_exact(G, shape, deg):
    """Below 2^20 MACs the exp recurrence runs in the reference's summation order (one warp per row segment, separate
    multiply and add, row sums from zero): bit-identical to the reference, like the reference-order product kernel.  (The
    seed exp(x[0]) comes from the host's libm; the 1-d recurrence of leaf 0 is sequential up to 32 terms per coefficient, as in
    the 1-d kernel -- hence the argument rows of at most 32 coefficients here.)"""
    rng = np.random.default_rng(5 + len(shape))
    a = rng.uniform(-1.0, 1.0, shape)
    g, o = both(G, a, deg)
    assert_same(g.exp(), o.exp())


def test_recurrences_with_ragged_operands(G):
    """Divisor / argument shorter than the result on some axes, constant along others (zero extension, uncoupled axes)."""
    rng = np.random.default_rng(77)
    for xs, ys, deg in (((5, 6, 4), (2, 1, 3), (6, 6, 6)), ((1, 7, 5), (3, 4, 1), (4, 8, 6)), ((6, 2), (2, 5), (9, 9)),
                        ((3, 1, 4, 2), (2, 2, 1, 2), (4, 3, 5, 3))):
        x, y = rng.standard_normal(xs), rng.uniform(0.5, 1.5, ys)
        gx, ox = both(G, x, deg)
        gy, oy = both(G, y, deg)
        assert_close(gx / gy, ox / oy, rtol=1e-12)
        assert_close(gy.exp(), oy.exp(), rtol=1e-12)
        assert_close(gy.log(), oy.log(), rtol=1e-12)


def test_recurrence_kernel_edge_shapes(G):
    """Shortest rows, seven leaf axes, a row longer than one CTA's threads, arguments shorter than the result."""
    rng = np.random.default_rng(2024)
    for shape, deg in (((2, 2), None), ((2,) * 8, None), ((2, 1500), (3, 1600)), ((3, 2, 2), (9, 9, 9)), ((1, 4, 1, 3), (2, 5, 2, 4))):
        a = rng.uniform(0.5, 1.5, shape)
        x = rng.standard_normal(shape)
        ga, oa = both(G, a, deg)
        gx, ox = both(G, x, deg)
        assert_close(ga.exp(), oa.exp(), rtol=1e-12)
        assert_close(ga.log(), oa.log(), rtol=1e-12)
        assert_close(gx / ga, ox / oa, rtol=1e-12)


def test_fused_horner_with_unbounded_degrees_and_32_terms(G):
    """`simplify` runs subst_var with degrees_p1 = usize::MAX (exact polynomial algebra, generating_function.rs:485, :569);
    a 2^5-coefficient substitution is the largest the fused loop takes."""
    from genfer_b200 import UNBOUNDED
    rng = np.random.default_rng(11)
    a = rng.standard_normal((3, 4, 5))
    sub = rng.standard_normal((2, 2, 2))
    g, o = G.TaylorPoly.new(a, [UNBOUNDED] * 3), O().TaylorPoly.new(a, [O().UMAX] * 3)
    gs, os_ = G.TaylorPoly.new(sub, [UNBOUNDED] * 3), O().TaylorPoly.new(sub, [O().UMAX] * 3)
    for v in range(3):
        assert_same(g.subst_var(v, gs), o.subst_var(v, os_))
    a5 = rng.standard_normal((3, 3, 3, 3, 4))
    s5 = rng.standard_normal((2, 2, 2, 2, 2))
    g, o = both(G, a5, (5, 5, 5, 5, 6))
    gs, os_ = both(G, s5, (5, 5, 5, 5, 6))
    assert_same(g.subst_var(4, gs), o.subst_var(4, os_))


def test_pinned_upload_and_trim(G):
    """gtp_from_host may be handed pinned memory that is overwritten as soon as the call returns (advisor finding, round 1);
    gtp_ctx_trim returns the cached blocks and the context keeps working."""
    import torch
    ctx = G.default_context()
    h = torch.arange(1 << 20, dtype=torch.float64).pin_memory()
    expect = h.numpy().copy()
    t = G.TaylorPoly.from_host_ptr(h.data_ptr(), (1 << 10, 1 << 10), (1 << 10, 1 << 10), ctx)
    h.fill_(-1.0)                                   # reuse the staging buffer at once
    assert np.array_equal(t.array().ravel(), expect)
    del t
    ctx.check(ctx.lib.gtp_ctx_trim(ctx.h))
    u = G.TaylorPoly.new(expect[:100], (100,), ctx)
    assert np.array_equal((u + u).array(), 2 * expect[:100])


# ---------------------------------------------------------------------------------------------
# north_star check 2: the f64 results lie inside the --bounds (Interval<F64>) enclosure
# ---------------------------------------------------------------------------------------------
ENCLOSED_OPS = {
    "mul": lambda x, y: x * y,
    "div": lambda x, y: x / y,
    "add_sub": lambda x, y: (x + y) - (y * x),
    "log_of_product": lambda x, y: (x * y).log(),
    "exp_of_difference": lambda x, y: (x - y).exp(),
    "pow5": lambda x, y: x.pow(5),
    "subst_var": lambda x, y: x.subst_var(1, y),
    "shift_down": lambda x, y: (x * y).shift_down(0, 2),
    "derivative": lambda x, y: (x * y).derivative(1, 2),
    "taylor_expansion_of_coeff": lambda x, y: (x / y).taylor_expansion_of_coeff(0, 1),
}


@pytest.mark.parametrize("name", sorted(ENCLOSED_OPS))
@pytest.mark.parametrize("shape", [(5, 5), (3, 4, 3), (12, 12)])
def test_gpu_results_inside_interval_enclosure(G, name, shape):
    """The device f64 result of every operator family must fall inside the result of the same operator over
    Interval<F64> (src/interval.rs, restated in the oracle) started from point intervals: the property the reference's
    --bounds mode guarantees for its own f64 path.  Covers the DFMA product kernels too (mode 2 forces them): fused
    multiply-add rounds once, the enclosure is of the exact result, so it must still contain it."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((name, shape)).encode()))
    a, b = rng.uniform(0.5, 2.0, shape), rng.uniform(0.5, 2.0, shape)
    op = ENCLOSED_OPS[name]
    iv = op(O().TaylorPoly.new(np.stack([a, a], -1), shape, kind="iv"),
            O().TaylorPoly.new(np.stack([b, b], -1), shape, kind="iv")).array()
    for mode in (0, 1, 2):
        G.default_context().set_fast_mul(mode)
        got = op(G.TaylorPoly.new(a, shape), G.TaylorPoly.new(b, shape)).array()
        G.default_context().set_fast_mul(1)
        assert got.shape == iv.shape[:-1]
        assert np.all(iv[..., 0] <= got) and np.all(got <= iv[..., 1]), (name, mode)


@pytest.mark.parametrize("shape,deg", [((6,), (6,)), ((6,), (9,)), ((4, 5), (4, 5)), ((4, 5), (6, 7)), ((3, 4, 5), (3, 5, 5)),
                                       ((2, 3, 4, 5), (4, 4, 4, 6)), ((33, 17), (40, 17)), ((1, 7), (3, 8))])
def test_mul_linear_fused_bit_exact(G, shape, deg):
    """Mul by `c + m eps_v` (:1052-1061 -> mul_linear :611-623): the one-pass kernel, the reference's composition
    (mode bit 128 switches the fused kernel off) and the oracle agree bit for bit, for every variable, for c == 0
    (mul_var alone), with the result capped by degrees_p1 or one slice longer, and with signed zeros in the data."""
    rng = np.random.default_rng(len(shape) * 1000 + sum(shape))
    a = rng.standard_normal(shape)
    a.flat[rng.integers(0, a.size, max(1, a.size // 7))] = -0.0
    a.flat[rng.integers(0, a.size, max(1, a.size // 9))] = 0.0
    ctx = G.default_context()
    for v in range(len(shape)):
        for c0, m in ((0.5, -1.5), (0.0, 2.0), (1.0, 1.0), (-2.0, 0.25)):
            lin_shape = [1] * len(shape)
            lin_shape[v] = 2
            lin = np.zeros(lin_shape)
            lin.flat[0], lin.flat[1] = c0, m
            results = []
            for mode in (1, 129):
                ctx.set_fast_mul(mode)
                x, ox = both(G, a, deg)
                l, ol = both(G, lin, deg)
                results.append(((x * l), (l * x)))
            ctx.set_fast_mul(1)
            ref = ox * ol
            for r1, r2 in results:
                assert_same(r1, ref)
                assert_same(r2, ol * ox)


def test_python_scalar_operands_stay_alive(G):
    """`poly * 0.75` coerces the float to a temporary TaylorPoly; it must outlive the C call (it used to be freed
    between argument evaluation and the call)."""
    a = np.arange(12.0).reshape(3, 4) + 1.0
    g, o = both(G, a)
    for _ in range(50):
        assert_same(g * 0.75, o * O().TaylorPoly.from_scalar(0.75))
        assert_same(g + 2.0, o + O().TaylorPoly.from_scalar(2.0))
        assert_same(g - 0.5, o - O().TaylorPoly.from_scalar(0.5))
        assert_same(g / 4.0, o / O().TaylorPoly.from_scalar(4.0))


# ---------------------------------------------------------------------------------------------
# edge cases: 0-d scalars, many variables with unit axes, the maximum rank, IEEE specials
# ---------------------------------------------------------------------------------------------
def same_with_nans(g, o):
    assert_meta_equal(g, o)
    ga, oa = g.array(), o.array()
    assert np.array_equal(np.isnan(ga), np.isnan(oa)), f"\n gpu={ga}\n ref={oa}"
    ok = ~np.isnan(oa)
    assert np.array_equal(ga[ok].view(np.uint64), oa[ok].view(np.uint64)), f"\n gpu={ga}\n ref={oa}"


def test_scalar_polynomials_every_operator(G):
    """0-dimensional operands (TaylorPoly::from(c), :626-630) through every operator."""
    for x, y in ((2.5, -0.75), (0.0, 3.0), (1.0, 1.0), (-0.0, 2.0)):
        gx, ox = G.TaylorPoly.from_scalar(x), O().TaylorPoly.from_scalar(x)
        gy, oy = G.TaylorPoly.from_scalar(y), O().TaylorPoly.from_scalar(y)
        for f in (lambda a, b: a + b, lambda a, b: a - b, lambda a, b: a * b, lambda a, b: a / b, lambda a, b: -a,
                  lambda a, b: a.pow(3), lambda a, b: (a * a + b * b).exp(), lambda a, b: (a * a + b * b).log()):
            g, o = f(gx, gy), f(ox, oy)
            assert_meta_equal(g, o)
            assert_close(g, o)
        assert gx.is_zero() == ox.is_zero() and gx.is_one() == ox.is_one()
        assert gx.extract_constant() == x and ox.constant_term() == x


def test_many_variables_with_unit_axes(G):
    """Ten program variables, most axes of stored length 1 (the compact supports of the prodigy programs)."""
    rng = np.random.default_rng(10)
    shape = (2, 1, 3, 1, 1, 2, 1, 1, 1, 2)
    deg = (3, 2, 3, 4, 1, 2, 5, 1, 2, 3)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    (ga, oa), (gb, ob) = both(G, a, deg), both(G, b, deg)
    assert_same(ga + gb, oa + ob)
    assert_same(ga - gb, oa - ob)
    assert_same(ga * gb, oa * ob)
    assert_same(ga.derivative(2, 1), oa.derivative(2, 1))
    assert_same(ga.shift_down(9, 1), oa.shift_down(9, 1))
    assert_same(ga.shift_down(0, 1), oa.shift_down(0, 1))
    assert_same(ga.coefficients_of_term(5, 1), oa.coefficients_of_term(5, 1))
    assert_same(ga.taylor_expansion_of_coeff(2, 2), oa.taylor_expansion_of_coeff(2, 2))
    assert_close(ga * gb * ga, oa * ob * oa)


def test_maximum_rank(G):
    """GTP_MAX_NDIM = 24 variables (two non-unit axes)."""
    shape = [1] * 24
    shape[3], shape[23] = 3, 4
    rng = np.random.default_rng(24)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    (ga, oa), (gb, ob) = both(G, a), both(G, b)
    assert_same(ga + gb, oa + ob)
    assert_same(ga * gb, oa * ob)
    assert_same(ga.derivative(23, 2), oa.derivative(23, 2))
    assert_same(ga.shift_down(3, 1), oa.shift_down(3, 1))


def test_division_by_zero_constant_term_propagates_ieee_specials(G):
    """:1167 -- plain f64 division: a divisor series with zero constant term yields inf / NaN, no trap, no error."""
    a = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    for b in (np.array([[0.0, 1.0, 0.0]]), np.array([[0.0]]), np.array([[0.0, 1.0], [2.0, 0.0]])):
        (ga, oa), (gb, ob) = both(G, a, (2, 3)), both(G, b, (2, 3))
        if b.shape == (2, 2):   # general N-D path: compare the exact-order kernels
            G.default_context().set_fast_mul(0)
        try:
            same_with_nans(ga / gb, oa / ob)
        finally:
            G.default_context().set_fast_mul(1)
    big = np.array([1e308, 1e308, 1.0])
    (gx, ox) = both(G, big)
    same_with_nans(gx * gx, ox * ox)            # overflow to +inf
    same_with_nans((gx * gx) - (gx * gx), (ox * ox) - (ox * ox))   # inf - inf = NaN
