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


from genfer_b200.synth import splitmix64, synth_pgf, synth_uniform  # noqa: E402,F401
