"""TaylorPoly<Interval<F64>> on the device (gti_* in csrc/interval_api.cu) against the oracle's Interval<F64>
instantiation of the reference's generic TaylorPoly<T> (oracle kind 'iv': interval.rs arithmetic).

Bar: the element-wise / gather family and the truncated product visit their terms in the reference's order and must agree with
the oracle BIT FOR BIT (both bounds of every coefficient).  div / exp / log use coefficient-wise recurrences (a different but
equally valid widening sequence): their enclosures must overlap the oracle's on every coefficient, contain the f64 result, and be
of comparable width.  End to end, gtp_run_sgcl_bounds must enclose the GPU's own f64 results and overlap the oracle's enclosure.
The interval path is unpinned against the reference (no reference fixture runs with --bounds), like the oracle's.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sgcl")


@pytest.fixture(scope="module")
def ctx():
    import genfer_b200
    c = genfer_b200.Context(0)
    yield c
    c.close()


def rand_iv(rng, shape, width=1e-9, lo=-1.0, hi=1.0):
    mid = rng.uniform(lo, hi, size=shape)
    w = rng.uniform(0.0, width, size=shape)
    return np.stack([mid - w, mid + w], axis=-1)


def pair(ctx, arr, degrees):
    from genfer_b200.interval import IntervalPoly
    from oracle import oracle as O
    return IntervalPoly.new(arr, degrees, ctx), O.TaylorPoly.new(arr, degrees, "iv")


def same(g, o):
    assert g.array_shape() == o.array_shape(), (g.array_shape(), o.array_shape())
    assert g.degrees_p1() == tuple(o.shape()), (g.degrees_p1(), o.shape())
    ga, oa = g.array(), o.array()
    assert np.array_equal(ga, oa, equal_nan=True), (np.abs(ga - oa).max(), np.argwhere(ga != oa)[:4])


def compatible(g, o, point=None, slack=8.0):
    """Both are enclosures of the same exact value: they overlap; the device's is at most `slack` times as wide."""
    assert g.array_shape() == o.array_shape(), (g.array_shape(), o.array_shape())
    ga, oa = g.array(), o.array()
    assert np.all(ga[..., 0] <= ga[..., 1])
    assert np.all(np.maximum(ga[..., 0], oa[..., 0]) <= np.minimum(ga[..., 1], oa[..., 1])), "enclosures do not overlap"
    wg, wo = ga[..., 1] - ga[..., 0], oa[..., 1] - oa[..., 0]
    tiny = 64 * np.finfo(float).eps * np.maximum(np.abs(oa).max(axis=-1), 1e-300)
    assert np.all(wg <= slack * wo + tiny), float((wg / (wo + tiny)).max())
    if point is not None:
        assert np.all((ga[..., 0] <= point) & (point <= ga[..., 1])), "f64 result outside the enclosure"


BIN_SHAPES = [((5,), (5,), (5,), (5,)), ((4, 6), (4, 6), (3, 2), (4, 6)), ((3, 4, 5), (3, 4, 5), (3, 4, 5), (3, 4, 5)),
              ((7, 3), (9, 5), (2, 5), (9, 5)), ((6, 1, 4), (6, 3, 4), (1, 3, 2), (6, 3, 4)), ((2, 2, 2, 2, 2), (3,) * 5, (2,) * 5, (3,) * 5),
              ((40, 30), (40, 30), (25, 30), (40, 30)), ((12, 12, 12), (12,) * 3, (12,) * 3, (12,) * 3)]


@pytest.mark.parametrize("xs,xd,ys,yd", BIN_SHAPES)
def test_add_sub_mul_bit_exact(ctx, xs, xd, ys, yd):
    rng = np.random.default_rng(hash((xs, ys)) % 2**32)
    gx, ox = pair(ctx, rand_iv(rng, xs), xd)
    gy, oy = pair(ctx, rand_iv(rng, ys), yd)
    same(gx + gy, ox + oy)
    same(gx - gy, ox - oy)
    same(gy - gx, oy - ox)
    same(-gx, -ox)
    same(gx * gy, ox * oy)
    same(gy * gx, oy * ox)


def test_scalar_zero_one_paths(ctx):
    from genfer_b200.interval import IntervalPoly
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    gx, ox = pair(ctx, rand_iv(rng, (4, 5)), (6, 5))
    for c in ([0.0, 0.0], [1.0, 1.0], [-1.0, -1.0], [0.3, 0.30000001], [-2.0, 1.0]):
        gc, oc = IntervalPoly.from_scalar(c, ctx), O.TaylorPoly.from_scalar(c, "iv")
        same(gx * gc, ox * oc)
        same(gc * gx, oc * ox)
        same(gx + gc, ox + oc)
        same(gc - gx, oc - ox)
        same(gx - gc, ox - oc)
        if c != [0.0, 0.0]:
            same(gx / gc, ox / oc)
    gz, oz = IntervalPoly.zero_with((3, 3), ctx), O.TaylorPoly.zero_with((3, 3), "iv")
    same(gx * gz, ox * oz)
    same(gx + gz, ox + oz)
    gv, ov = IntervalPoly.var(1, (0.25, 0.26), 4, ctx), O.TaylorPoly.var(1, [0.25, 0.26], 4, "iv")
    same(gv, ov)
    same(gx * gv * gv + gv, ox * ov * ov + ov)
    same(IntervalPoly.var_at_zero(2, 3, ctx), O.TaylorPoly.var_at_zero(2, 3, "iv"))
    same(IntervalPoly.var_with_degrees_p1(1, 0.5, (3, 2, 4), ctx), O.TaylorPoly.var_with_degrees_p1(1, [0.5, 0.5], (3, 2, 4), "iv"))
    assert gx.constant_term() == tuple(ox.constant_term())
    assert (gx * gz).extract_constant() == (0.0, 0.0) and gx.extract_constant() is None


@pytest.mark.parametrize("shape,deg", [((6,), (6,)), ((5, 4), (7, 4)), ((3, 6, 4), (3, 9, 4)), ((9, 2, 2, 3), (9, 2, 2, 3))])
def test_gather_family_bit_exact(ctx, shape, deg):
    rng = np.random.default_rng(len(shape) * 17 + shape[0])
    g, o = pair(ctx, rand_iv(rng, shape), deg)
    for v in range(len(shape)):
        for n in (0, 1, 2, shape[v] - 1):
            if n >= deg[v]:
                continue
            same(g.derivative(v, n), o.derivative(v, n))
            same(g.taylor_expansion_of_coeff(v, n), o.taylor_expansion_of_coeff(v, n))
            same(g.shift_down(v, n), o.shift_down(v, n))
            same(g.coefficients_of_term(v, n), o.coefficients_of_term(v, n))
        same(g.taylor_polynomial_terms(v, [0, 2]), o.taylor_polynomial_terms(v, [0, 2]))
        same(g.taylor_polynomial_terms(v, [1]), o.taylor_polynomial_terms(v, [1]))
        got = g.gather_axis(v, shape[v] + 2)
        for i in range(shape[v] + 2):
            idx = [0] * len(shape)
            idx[v] = i
            want = o.coefficient(idx) if i < shape[v] else np.zeros(2)
            assert np.array_equal(got[i], want)
    same(g.truncate_to_degree_p1(2), o.truncate_to_degree_p1(2))
    same(g.extend_to_dim(len(shape) + 2, 5), o.extend_to_dim(len(shape) + 2, 5))
    same(g.remove_last_variable(), o.remove_last_variable())
    same(g.pow(0), o.pow(0))
    same(g.pow(1), o.pow(1))
    same(g.pow(3), o.pow(3))


REC_SHAPES = [((8,), (8,)), ((5, 6), (5, 6)), ((3, 1, 4), (6, 1, 7)), ((4, 4, 4), (4, 4, 4)), ((2, 3), (9, 8)), ((6, 5, 3, 2), (6, 5, 3, 2)),
              ((20, 20), (20, 20))]


@pytest.mark.parametrize("shape,deg", REC_SHAPES)
def test_div_exp_log_enclosures(ctx, shape, deg):
    import genfer_b200
    from oracle import oracle as O
    rng = np.random.default_rng(sum(shape) * 31 + len(shape))
    a = rng.uniform(-0.5, 0.5, size=shape)
    a.flat[0] = 1.5
    b = rng.uniform(-0.3, 0.3, size=shape)
    b.flat[0] = 2.0
    w = 1e-12
    ai, bi = np.stack([a - w, a + w], -1), np.stack([b - w, b + w], -1)
    (ga, oa), (gb, ob) = pair(ctx, ai, deg), pair(ctx, bi, deg)
    fa, fb = genfer_b200.TaylorPoly.new(a, deg, ctx), genfer_b200.TaylorPoly.new(b, deg, ctx)
    compatible(ga / gb, oa / ob, (fa / fb).array())
    compatible(ga.exp(), oa.exp(), fa.exp().array())
    compatible(ga.log(), oa.log(), fa.log().array())
    # division by a lower-dimensional / smaller series and by a truncated one
    small = tuple(min(s, 2) for s in shape)
    bs = b[tuple(slice(0, s) for s in small)]
    gs, os_ = pair(ctx, np.stack([bs - w, bs + w], -1), deg)
    compatible(ga / gs, oa / os_, (fa / genfer_b200.TaylorPoly.new(bs, deg, ctx)).array())
    # exp(log(a)) encloses a
    back = ga.log().exp().array()
    full = np.zeros(back.shape[:-1])
    full[tuple(slice(0, s) for s in shape)] = a
    assert np.all((back[..., 0] <= full) & (full <= back[..., 1]))


def test_point_inputs_enclose_exact_rationals(ctx):
    """1 / (1 - x - y) has the binomial coefficients C(i + j, i) as its Taylor coefficients: the enclosure contains them."""
    from math import comb
    from genfer_b200.interval import IntervalPoly
    one = IntervalPoly.from_scalar(1.0, ctx)
    den = one - IntervalPoly.var(0, 0.0, 12, ctx) - IntervalPoly.var(1, 0.0, 12, ctx).extend_to_dim(2, 12)
    q = (one / den).array()
    assert q.shape == (12, 12, 2)
    for i in range(12):
        for j in range(12):
            assert q[i, j, 0] <= comb(i + j, i) <= q[i, j, 1]
            assert q[i, j, 1] - q[i, j, 0] <= 1e-9 * comb(i + j, i)
    e = IntervalPoly.var(0, 0.0, 15, ctx).exp().array()    # 1 / k!
    f = 1.0
    for k in range(15):
        f = f * k if k else 1.0
        assert e[k, 0] <= 1.0 / f <= e[k, 1]


def test_subst_var_encloses(ctx):
    import genfer_b200
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, size=(4, 5, 3))
    s = rng.uniform(-0.5, 0.5, size=(3, 3, 3))
    s.flat[0] = 0.0
    deg = (6, 6, 6)
    (ga, oa), (gs, os_) = pair(ctx, np.stack([a, a], -1), deg), pair(ctx, np.stack([s, s], -1), deg)
    fa, fs = genfer_b200.TaylorPoly.new(a, deg, ctx), genfer_b200.TaylorPoly.new(s, deg, ctx)
    for v in range(3):
        compatible(ga.subst_var(v, gs), oa.subst_var(v, os_), fa.subst_var(v, fs).array())
    from genfer_b200.interval import IntervalPoly
    from oracle import oracle as O
    same(ga.subst_var(1, IntervalPoly.zero_with(deg, ctx)), oa.subst_var(1, O.TaylorPoly.zero_with(deg, "iv")))


def bounds_fixtures():
    from helpers import ENCLOSURE_SLOW
    out = [os.path.relpath(s, GOLD) for s in sorted(glob.glob(os.path.join(GOLD, "*", "*.sgcl")))
           if os.path.exists(s[:-5] + ".expect") and os.path.basename(s)[:-5] not in ENCLOSURE_SLOW and os.path.getsize(s) < 50_000]
    return out


@pytest.mark.parametrize("rel", bounds_fixtures())
def test_device_enclosure_end_to_end(ctx, rel):
    """north_star check 2 with the interval arithmetic itself on the device: the f64 results lie inside the GPU's enclosure, and
    the GPU's enclosure overlaps the oracle's (both enclose the exact value of the same DAG)."""
    import math
    import genfer_b200
    from genfer_b200.interval import run_sgcl_bounds
    from helpers import check_inside_enclosure
    from oracle import oracle as O
    src = open(os.path.join(GOLD, rel)).read()
    opts = genfer_b200.parse_flags(src)
    g = genfer_b200.run_sgcl(src, limit=opts["limit"], no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"],
                             unroll=opts["unroll"], ctx=ctx)
    b = run_sgcl_bounds(src, limit=len(g.probs), unroll=opts["unroll"], ctx=ctx)
    check_inside_enclosure(g, b)
    o = O.run_sgcl_bounds(src, limit=len(g.probs), unroll=opts["unroll"])
    for (glo, ghi), (olo, ohi) in zip([b.rest, b.total] + b.raw_moments + b.probs, [o.rest, o.total] + o.raw_moments + o.probs):
        if all(math.isfinite(x) for x in (glo, ghi, olo, ohi)):
            assert max(glo, olo) <= min(ghi, ohi), ((glo, ghi), (olo, ohi))
        else:
            assert math.isfinite(glo) == math.isfinite(olo) and math.isfinite(ghi) == math.isfinite(ohi), ((glo, ghi), (olo, ohi))


@pytest.mark.parametrize("rel", ["real_world/population2000.sgcl", "slow/two_populations2000.sgcl"])
def test_device_enclosure_on_baseline_programs(ctx, rel):
    """BASELINE C2 programs, where the oracle's interval run costs 20-60 s of CPU: the GPU f64 results are inside the GPU's enclosure."""
    import genfer_b200
    from genfer_b200.interval import run_sgcl_bounds
    from helpers import check_inside_enclosure
    path = os.path.join(GOLD, rel)
    if not os.path.exists(path):
        pytest.skip("fixture not committed")
    src = open(path).read()
    opts = genfer_b200.parse_flags(src)
    g = genfer_b200.run_sgcl(src, limit=opts["limit"], no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"],
                             unroll=opts["unroll"], ctx=ctx)
    b = run_sgcl_bounds(src, limit=len(g.probs), unroll=opts["unroll"], ctx=ctx)
    assert check_inside_enclosure(g, b) > 0


def report_skeleton(report):
    """The report with every number replaced: what must be identical between two --bounds runs whose enclosures differ in ulps."""
    import re
    return re.sub(r"-?(?:\d+\.\d+(?:e-?\d+)?|inf|NaN)", "#", report)


@pytest.mark.parametrize("rel", bounds_fixtures())
def test_bounds_report_matches_oracle(ctx, rel):
    """`--bounds` end to end through gtp_run_sgcl (flags & 4): same report layout as the oracle's interval instantiation, every
    printed enclosure overlaps the oracle's, and the GPU's own f64 results lie inside it."""
    import math
    import genfer_b200
    from oracle import oracle as O
    src = open(os.path.join(GOLD, rel)).read()
    opts = genfer_b200.parse_flags(src)
    kw = dict(limit=opts["limit"], no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"], unroll=opts["unroll"])
    g = genfer_b200.run_sgcl(src, ctx=ctx, bounds=True, **kw)
    o = O.run_sgcl(src, bounds=True, **kw)
    assert report_skeleton(g.report) == report_skeleton(o.report)
    f = genfer_b200.run_sgcl(src, ctx=ctx, **kw)
    assert len(g.prob_bounds) == len(o.prob_bounds) == len(f.probs)
    for name, (glo, ghi), (olo, ohi) in zip(["Z", "E", "raw2", "raw3", "raw4", "sigma", "V", "mu3", "mu4", "S", "K"] + [f"p({i})" for i in range(len(f.probs))],
                                            g.moment_bounds + g.prob_bounds, o.moment_bounds + o.prob_bounds):
        if all(math.isfinite(x) for x in (glo, ghi, olo, ohi)):
            assert max(glo, olo) <= min(ghi, ohi), (name, (glo, ghi), (olo, ohi))
    for name, value, (lo, hi) in zip(["Z", "E", "raw2", "raw3", "raw4"], f.moments[:5], g.moment_bounds[:5]):
        if math.isfinite(lo) and math.isfinite(hi) and math.isfinite(value):
            assert lo <= value <= hi, (name, value, (lo, hi))
    for i, (p, (lo, hi)) in enumerate(zip(f.probs, g.prob_bounds)):
        if math.isfinite(lo) and math.isfinite(hi):
            assert lo <= p <= hi, (i, p, (lo, hi))


def test_bounds_mode_encloses_exact_posteriors_on_gpu(ctx):
    """With ratio constants enclosed (Number::from_ratio) the device's intervals contain the exact posterior: C1's closed form
    Z = 2 e^-2, p(n) = e^-10 10^n / n! * n 0.2 0.8^(n-1), and Pr[burglary] = 2969983/992160802 of prodigy/burglar_alarm."""
    import math
    from decimal import Decimal, getcontext
    from fractions import Fraction
    import genfer_b200
    getcontext().prec = 50
    src = open(os.path.join(GOLD, "config", "example.sgcl")).read()
    b = genfer_b200.run_sgcl(src, limit=25, ctx=ctx, bounds=True)
    lo, hi = b.moment_bounds[0]
    assert Decimal(lo) <= 2 * Decimal(-2).exp() <= Decimal(hi) and hi - lo < 1e-14
    e10 = Decimal(-10).exp()
    for n, (lo, hi) in enumerate(b.prob_bounds):
        exact = e10 * Decimal(10) ** n / Decimal(math.factorial(n)) * n * Decimal(2) / 10 * (Decimal(8) / 10) ** (n - 1) if n else Decimal(0)
        assert Decimal(lo) <= exact <= Decimal(hi), (n, lo, exact, hi)
    src = open(os.path.join(GOLD, "config", "burglar_alarm.sgcl")).read()
    b = genfer_b200.run_sgcl(src, ctx=ctx, bounds=True)
    lo, hi = b.normalized_prob_bounds[1]
    assert Fraction(lo) <= Fraction(2969983, 992160802) <= Fraction(hi) and hi - lo < 1e-15
    assert "Normalized:   p(1) / Z ∈ [" in b.report
