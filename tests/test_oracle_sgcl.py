"""Pins the oracle end to end: the host evaluator over the CPU restatement reproduces the reference's stdout
BYTE FOR BYTE on the reference's own golden files (tests/golden/sgcl/**, copied by make_sgcl_fixtures.py from
test/expect/** and benchmarks/neurips2023/**; harness restated from tests/integration.rs:18-60).  CPU only."""
import glob
import os

import pytest

from genfer_b200.evaluator import parse_flags
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sgcl")
SLOW = {"mixture"}   # 35 s on one CPU core: covered on the GPU, and here only with RUN_SLOW_TESTS=1


def fixtures():
    out = []
    for src in sorted(glob.glob(os.path.join(GOLD, "*", "*.sgcl"))):
        if os.path.exists(src[:-5] + ".expect"):
            out.append(os.path.relpath(src, GOLD))
    return out


@pytest.mark.parametrize("rel", fixtures())
def test_oracle_reproduces_reference_stdout(rel):
    name = os.path.basename(rel)[:-5]
    if name in SLOW and not os.environ.get("RUN_SLOW_TESTS"):
        pytest.skip("slow on CPU")
    src = open(os.path.join(GOLD, rel)).read()
    expect = open(os.path.join(GOLD, rel[:-5] + ".expect")).read()
    opts = parse_flags(src)
    assert not opts["unsupported"]
    got = O.run_sgcl(src, limit=opts["limit"], no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"],
                     unroll=opts["unroll"])
    assert got.report == expect


def test_example_config_c1_closed_form():
    """example.sgcl --limit 25 (BASELINE config C1): Z = 2 e^-2, E = 9, p(n) = e^-10 10^n/n! * n 0.2 0.8^(n-1)."""
    import math
    src = open(os.path.join(GOLD, "config", "example.sgcl")).read()
    r = O.run_sgcl(src, limit=25)
    assert abs(r.total - 2 * math.exp(-2)) <= 1e-15
    assert abs(r.mean - 9.0) <= 1e-12
    for n, p in enumerate(r.probs):
        exact = math.exp(-10) * 10.0 ** n / math.factorial(n) * n * 0.2 * 0.8 ** (n - 1) if n else 0.0
        assert abs(p - exact) <= 1e-15 + 1e-12 * exact


def test_prodigy_burglar_alarm_exact_rational():
    """benchmarks/prodigy/burglar_alarm.sgcl quotes Pr[burglary] = 2969983/992160802 in its `Original code` comment
    (also benchmarks/neurips2023/exact/alarm/alarm.expected)."""
    src = open(os.path.join(GOLD, "config", "burglar_alarm.sgcl")).read()
    r = O.run_sgcl(src)
    assert abs(r.normalized_probs[1] - 2969983 / 992160802) <= 1e-15


def enclosure_fixtures():
    from helpers import ENCLOSURE_SLOW
    return [rel for rel in fixtures() if os.path.basename(rel)[:-5] not in ENCLOSURE_SLOW
            and os.path.getsize(os.path.join(GOLD, rel)) < 50_000]


@pytest.mark.parametrize("rel", enclosure_fixtures())
def test_f64_results_inside_interval_enclosure(rel):
    """north_star check 2 on the CPU side: Z, the raw moments and p(n) of the f64 evaluation lie inside the enclosure the
    same host logic computes over TaylorPoly<Interval<F64>> (the --bounds arithmetic; unpinned: no reference fixture uses
    --bounds).  The GPU twin is tests/test_gpu_sgcl.py::test_gpu_results_inside_interval_enclosure."""
    from helpers import check_inside_enclosure
    src = open(os.path.join(GOLD, rel)).read()
    opts = parse_flags(src)
    r = O.run_sgcl(src, limit=opts["limit"], no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"], unroll=opts["unroll"])
    b = O.run_sgcl_bounds(src, limit=len(r.probs), unroll=opts["unroll"])
    check_inside_enclosure(r, b)


def test_bounds_mode_encloses_exact_posteriors():
    """`--bounds` (run_program_intervals::<F64>, main.rs:145-185): ratio constants enter as the enclosures Number::from_ratio
    builds, so the printed intervals contain the EXACT posterior -- checked on the closed form of C1 and the exact rational of
    prodigy/burglar_alarm.  (Unpinned against the reference: none of its fixtures runs with --bounds.)"""
    from fractions import Fraction
    from decimal import Decimal, getcontext
    getcontext().prec = 50
    src = open(os.path.join(GOLD, "config", "example.sgcl")).read()
    b = O.run_sgcl(src, limit=25, bounds=True)
    z = 2 * Decimal(-2).exp()
    lo, hi = b.moment_bounds[0]
    assert Decimal(lo) <= z <= Decimal(hi) and hi - lo < 1e-14
    lo, hi = b.moment_bounds[1]
    assert lo <= 9.0 <= hi and hi - lo < 1e-12
    e10 = Decimal(-10).exp()
    for n, (lo, hi) in enumerate(b.prob_bounds):
        exact = e10 * Decimal(10) ** n / Decimal(math_factorial(n)) * n * Decimal(2) / 10 * (Decimal(8) / 10) ** (n - 1) if n else Decimal(0)
        assert Decimal(lo) <= exact <= Decimal(hi), (n, lo, exact, hi)
    assert "Z ∈ [" in b.report and "p(1)     ∈ [" in b.report
    src = open(os.path.join(GOLD, "config", "burglar_alarm.sgcl")).read()
    b = O.run_sgcl(src, bounds=True)
    lo, hi = [float(x) for x in b.report.split("Normalized:   p(1) / Z ∈ [")[1].split("]")[0].split(", ")]
    assert Fraction(lo) <= Fraction(2969983, 992160802) <= Fraction(hi) and hi - lo < 1e-15


def math_factorial(n):
    import math
    return math.factorial(n)


def test_from_ratio_enclosures():
    """Number::from_ratio for Interval<F64> (number/number.rs:26-33): widened even when the quotient is exact; zero and
    denominators of one stay points (Interval's shortcuts, interval.rs:199-206)."""
    src = "X ~ Bernoulli(1/2);\nreturn X;\n"
    b = O.run_sgcl(src, bounds=True)
    lo, hi = b.prob_bounds[1]
    assert lo < 0.5 < hi and hi - lo <= 4 * 2.0 ** -53
    f = O.run_sgcl(src)
    assert f.probs[1] == 0.5


def test_from_ratio_with_a_denominator_beyond_32_bits():
    """Number::from_ratio splits numerator and denominator into 32-bit halves (number/number.rs:26-33); for Interval<F64> the
    recombination `hi * 2^32 + lo` is itself interval arithmetic, so the constant's enclosure is a few ulps wide -- it must still
    contain both the exact ratio and the f64 quotient the f64 path uses."""
    from fractions import Fraction
    src = "X ~ Bernoulli(0.0000000001);\nreturn X;\n"          # 1 / 10^10, 10^10 > 2^32
    b = O.run_sgcl(src, bounds=True)
    lo, hi = b.prob_bounds[1]
    assert Fraction(lo) <= Fraction(1, 10 ** 10) <= Fraction(hi)
    assert lo <= 1e-10 <= hi and (hi - lo) <= 16 * 2.0 ** -53 * 1e-10
    assert O.run_sgcl(src).probs[1] == 1.0 / 1e10
