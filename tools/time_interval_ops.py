"""Developer tool: device time of a few TaylorPoly<Interval<F64>> operators (gti_*), one JSON line each."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genfer_b200
from genfer_b200.interval import IntervalPoly

ctx = genfer_b200.Context(0)
rng = np.random.default_rng(1)


def iv(shape):
    m = rng.uniform(0.5, 1.5, size=shape)
    return np.stack([m - 1e-12, m + 1e-12], -1)


def timed(fn, reps=5):
    fn(); ctx.synchronize()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); ctx.synchronize(); best = min(best, time.perf_counter() - t)
    return best


for shape in [(48, 48, 48), (24, 24, 24, 24), (300, 300)]:
    a, b = IntervalPoly.new(iv(shape), shape, ctx), IntervalPoly.new(iv(shape), shape, ctx)
    macs = genfer_b200.mul_macs(shape, shape, shape)
    n = int(np.prod(shape))
    t_mul, t_add = timed(lambda: a * b), timed(lambda: a + b)
    rec = {"shape": list(shape), "coefficients": n, "mul_ms": t_mul * 1e3, "interval_gmac_per_s": macs / t_mul / 1e9,
           "add_ms": t_add * 1e3, "add_gb_per_s": 3 * n * 16 / t_add / 1e9}
    if len(shape) <= 3:
        rec["div_ms"] = timed(lambda: a / b, 2) * 1e3
        rec["exp_ms"] = timed(lambda: a.exp(), 2) * 1e3
    print(json.dumps(rec), flush=True)
ctx.close()
