"""Size-independent algebraic properties of the oracle (CPU only, hypothesis): the checker itself must be a consistent
truncated-power-series algebra before the GPU is compared with it.  Each property is the identity the reference's own
tests use at fixed vectors (multivariate_taylor.rs: `test_log` exp/log round trip :1409-1436, `test_2d_div` :1131-1159
multiplies back, `test_2d_subst_var` :807-828), here on random shapes."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O

T = O.TaylorPoly
RTOL = 1e-9


@st.composite
def series_pair(draw, max_dims=3, max_len=5):
    nd = draw(st.integers(1, max_dims))
    shape = tuple(draw(st.integers(1, max_len)) for _ in range(nd))
    deg = tuple(s + draw(st.integers(0, 2)) for s in shape)
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1.0, 1.0, shape)
    b = rng.uniform(-1.0, 1.0, shape)
    a.flat[0] = rng.uniform(1.0, 2.0)       # invertible / positive constant terms
    b.flat[0] = rng.uniform(1.0, 2.0)
    return a, b, deg


def close(x, y, rtol=RTOL):
    xa, ya = x.array(), y.array()
    assert xa.shape == ya.shape, (xa.shape, ya.shape)
    assert x.shape() == y.shape()
    scale = max(np.max(np.abs(ya)), 1.0)
    assert np.max(np.abs(xa - ya)) <= rtol * scale, float(np.max(np.abs(xa - ya)))


@settings(max_examples=60, deadline=None)
@given(series_pair())
def test_product_commutes_and_distributes(p):
    a, b, deg = p
    x, y = T.new(a, deg), T.new(b, deg)
    close(x * y, y * x)
    close(x * (y + x), x * y + x * x)


@settings(max_examples=60, deadline=None)
@given(series_pair())
def test_division_inverts_multiplication(p):
    a, b, deg = p
    x, y = T.new(a, deg), T.new(b, deg)
    q = x / y
    back = q * y
    # the quotient carries the full degrees; compare on the stored block of x (zero-extended)
    xa = np.zeros(back.array().shape)
    xa[tuple(slice(0, s) for s in a.shape)] = a[tuple(slice(0, s) for s in xa.shape)]
    assert np.max(np.abs(back.array() - xa)) <= RTOL * max(np.max(np.abs(xa)), 1.0)


@settings(max_examples=40, deadline=None)
@given(series_pair(max_dims=2, max_len=5))
def test_exp_log_round_trip(p):
    a, _, deg = p
    x = T.new(a, deg)
    back = x.log().exp()
    full = np.zeros(back.array().shape)
    full[tuple(slice(0, s) for s in a.shape)] = a[tuple(slice(0, s) for s in full.shape)]
    assert np.max(np.abs(back.array() - full)) <= 1e-8 * max(np.max(np.abs(full)), 1.0)


@settings(max_examples=60, deadline=None)
@given(series_pair(), st.integers(0, 2), st.integers(0, 3))
def test_derivative_is_factorial_times_coefficient_expansion(p, v, n):
    a, _, deg = p
    v = v % len(deg)
    if n >= deg[v]:
        n = deg[v] - 1
    x = T.new(a, deg)
    d = x.derivative(v, n).array()
    c = x.taylor_expansion_of_coeff(v, n).array()
    assert d.shape == c.shape
    assert np.max(np.abs(d - math.factorial(n) * c)) <= RTOL * max(np.max(np.abs(d)), 1.0)


@settings(max_examples=60, deadline=None)
@given(series_pair(), st.integers(0, 2))
def test_substituting_a_variable_for_itself_is_the_identity(p, v):
    a, _, deg = p
    v = v % len(deg)
    if deg[v] < 2:        # `var` of a degree-0 axis is the constant 0 (:239-248): substituting it is a slice, not the identity
        return
    x = T.new(a, deg)
    same = x.subst_var(v, T.var(v, 0.0, deg[v]))
    xa, sa = x.array(), same.array()
    full = np.zeros(sa.shape)
    full[tuple(slice(0, s) for s in xa.shape)] = xa[tuple(slice(0, s) for s in sa.shape)]
    assert np.max(np.abs(sa - full)) <= RTOL * max(np.max(np.abs(full)), 1.0)


@settings(max_examples=60, deadline=None)
@given(series_pair(), st.integers(0, 2), st.integers(1, 3))
def test_shift_down_preserves_the_sum_of_coefficients(p, v, n):
    """shift_down folds the first n slices into slice 0 (:514-536): evaluate_all_one is invariant."""
    a, _, deg = p
    v = v % len(deg)
    n = min(n, deg[v] - 1)
    if n == 0:
        return
    x = T.new(a, deg)
    s = x.shift_down(v, n)
    assert abs(s.evaluate_all_one() - x.evaluate_all_one()) <= RTOL * max(abs(x.evaluate_all_one()), 1.0)


@settings(max_examples=40, deadline=None)
@given(series_pair(max_dims=2), st.integers(0, 4))
def test_pow_is_repeated_multiplication(p, e):
    a, _, deg = p
    x = T.new(a, deg)
    acc = T.from_scalar(1.0)
    for _ in range(e):
        acc = acc * x
    close(x.pow(e), acc) if e > 0 else None
