"""Shared helpers for the parity tests: oracle <-> device comparison."""
import numpy as np

from oracle import oracle as O

RTOL = 1e-12  # north_star: results within 1e-12 relative of the reference f64 path


def to_oracle(a, degrees=None):
    a = np.asarray(a, dtype=np.float64)
    return O.TaylorPoly.new(a, a.shape if degrees is None else degrees)


def assert_meta_equal(g, o):
    """Integer work must be bit-exact: stored shape and degrees_p1."""
    assert g.array_shape() == o.array_shape(), (g.array_shape(), o.array_shape())
    assert g.shape() == o.shape(), (g.shape(), o.shape())


def assert_same(g, o):
    """Bit-exact agreement (shape, degrees and every coefficient incl. signed zeros / NaN patterns)."""
    assert_meta_equal(g, o)
    ga, oa = g.array(), o.array()
    assert np.array_equal(ga.view(np.uint64), oa.view(np.uint64)), f"\n gpu={ga}\n ref={oa}"


def assert_close(g, o, rtol=RTOL, atol_scale=None):
    """|gpu - ref| <= rtol * max(|ref|, floor): floor = rtol * max|ref| guards entries that are exact
    zeros in the reference but tiny non-zeros after FMA/re-ordering (SURVEY 7 'parity budget')."""
    assert_meta_equal(g, o)
    ga, oa = g.array(), o.array()
    scale = np.max(np.abs(oa)) if oa.size else 0.0
    floor = (atol_scale if atol_scale is not None else scale) * rtol
    err = np.abs(ga - oa)
    bound = rtol * np.abs(oa) + floor
    bad = ~(err <= bound) & ~(np.isnan(ga) & np.isnan(oa)) & ~((ga == oa))
    assert not bad.any(), f"max rel err {np.max(err / np.maximum(np.abs(oa), 1e-300))}\n gpu={ga}\n ref={oa}"


def splitmix64(seed: int, n: int) -> np.ndarray:
    """Documented PRNG for synthetic tensors: uniform [0,1) doubles from splitmix64."""
    out = np.empty(n, dtype=np.uint64)
    state = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + state
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    out[:] = z
    return (out >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_uniform(shape, seed):
    """Distribution U of SURVEY 8(d): iid uniform [0,1)."""
    return splitmix64(seed, int(np.prod(shape))).reshape(shape)


def synth_pgf(shape, seed):
    """Distribution P of SURVEY 8(d): separable Poisson-like decay times (1 + 0.1 u)."""
    from math import lgamma
    u = synth_uniform(shape, seed)
    t = np.ones(shape)
    for ax, d in enumerate(shape):
        lam = d / 4.0
        k = np.arange(d)
        w = np.exp(-lam + k * np.log(lam) - np.array([lgamma(i + 1.0) for i in k]))
        sh = [1] * len(shape)
        sh[ax] = d
        t = t * w.reshape(sh)
    return t * (1.0 + 0.1 * u)
