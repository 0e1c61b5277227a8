"""Per-operator timing of the hot path's rows other than the plain product (a3, a6, a8-a16): the HBM-bound
gather / element-wise / reduction family against the measured HBM peak (algorithmic bytes = 8 x (doubles read + written),
SURVEY 8d Metric 1b / 1c), and the recurrences (div, exp, log, pow, subst_var) as time and product-equivalent FLOP/s, each
next to the CPU oracle on the same inputs where the oracle finishes in seconds.

usage: time_ops.py [--cpu] [--shape NxD] [--rec-shape NxD]        (developer tool; prints one JSON line per operator)
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import genfer_b200

ap = argparse.ArgumentParser()
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on the recurrence shapes")
ap.add_argument("--shape", default="6x16", help="tensor for the HBM-bound family")
ap.add_argument("--rec-shape", default="3x32,4x16", help="tensors for the recurrences")
args = ap.parse_args()

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = genfer_b200.Context(0, stream=stream.cuda_stream)
genfer_b200.set_default_context(ctx)
TP = genfer_b200.TaylorPoly


def timed(fns, reps=10, warm=8):
    """fns: callables run round-robin (different operands, so that reads come from HBM, not L2)."""
    keep = [None] * 3          # results stay alive for three calls: the pool must hand out a different output buffer each
    for i in range(warm):      # time (an output that is overwritten in place every call never leaves the 126 MB L2)
        keep[i % 3] = fns[i % len(fns)]()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    a.record()
    for i in range(reps):
        keep[i % 3] = fns[i % len(fns)]()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (ctx.launch_count - l0) / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- HBM-bound family on the big tensor -------------------------------------------------------------------
n, d = map(int, args.shape.split("x"))
shape = (d,) * n
N = d ** n
xs = [torch.rand(shape, dtype=torch.float64, device="cuda") for _ in range(3)]       # 3 x 128 MiB > L2
T = [TP.from_device(x.data_ptr(), shape, shape, ctx) for x in xs]
lin_last = TP.var(n - 1, 0.5, d)      # 0.5 + eps_v: the `mul_linear` operand (:611-623)
lin_first = TP.var(0, 0.5, d)
rows = [
    ("shift_down(last, D-1)  [a15]", lambda t: t.shift_down(n - 1, d - 1), N + N // d),
    ("shift_down(first, D-1) [a15]", lambda t: t.shift_down(0, d - 1), N + N // d),
    ("shift_down(last, 1)    [a15]", lambda t: t.shift_down(n - 1, 1), N + N // d * (d - 1)),
    ("shift_down(first, 1)   [a15]", lambda t: t.shift_down(0, 1), N + N // d * (d - 1)),
    ("derivative(last, 1)    [a13]", lambda t: t.derivative(n - 1, 1), 2 * (N // d) * (d - 1)),
    ("derivative(first, 2)   [a13]", lambda t: t.derivative(0, 2), 2 * (N // d) * (d - 2)),
    ("taylor_expansion_of_coeff(last, 1) [a14]", lambda t: t.taylor_expansion_of_coeff(n - 1, 1), 2 * (N // d) * (d - 1)),
    ("coefficients_of_term(first, 3) [a3]", lambda t: t.coefficients_of_term(0, 3), 2 * (N // d)),
    ("coefficients_of_term(last, 3)  [a3]", lambda t: t.coefficients_of_term(n - 1, 3), 2 * (N // d)),
    ("truncate_to_degree_p1(D/2) + materialise via neg [a3, a6]", lambda t: -t.truncate_to_degree_p1(d // 2), 2 * (d // 2) ** n),
    ("neg                    [a6]", lambda t: -t, 2 * N),
    ("mul_linear(last)       [a8]", lambda t: t * lin_last, 2 * N),
    ("mul_linear(first)      [a8]", lambda t: t * lin_first, 2 * N),
    ("mul by constant        [a7 fast path]", lambda t: t * 0.75, 2 * N),
]
for name, f, doubles in rows:
    ms, launches = timed([lambda t=t, f=f: f(t) for t in T])
    gbs = 8.0 * doubles / (ms * 1e-3) / 1e9
    emit(op=name, shape=args.shape, ms=round(ms, 4), algorithmic_bytes=8 * doubles, gbs=round(gbs, 1), frac_of_hbm=round(gbs / HBM, 3),
         hbm_peak=HBM, launches_per_call=launches)
ms, launches = timed([lambda: T[0] + T[1], lambda: T[1] + T[2], lambda: T[2] + T[0]])
gbs = 8.0 * 3 * N / (ms * 1e-3) / 1e9
emit(op="add                    [a6]", shape=args.shape, ms=round(ms, 4), algorithmic_bytes=24 * N, gbs=round(gbs, 1),
     frac_of_hbm=round(gbs / HBM, 3), hbm_peak=HBM, launches_per_call=launches)
del T, xs

# ---- stencil product (a7 with a small operand): the heavy product of the population models --------------------------
for xs_, ys_ in (((297, 282, 297), (2, 1, 2)), ((16,) * 6, (2, 1, 2, 1, 1, 2))):
    rs_ = tuple(a + b - 1 for a, b in zip(xs_, ys_))
    bigs = [torch.rand(xs_, dtype=torch.float64, device="cuda") for _ in range(3)]
    small = torch.rand(ys_, dtype=torch.float64, device="cuda")
    Tb = [TP.from_device(b.data_ptr(), xs_, rs_, ctx) for b in bigs]
    Ts = TP.from_device(small.data_ptr(), ys_, rs_, ctx)
    ms, launches = timed([lambda t=t: t * Ts for t in Tb])
    nx, nz = int(np.prod(xs_)), int(np.prod(rs_))
    gbs = 8.0 * (nx + nz) / (ms * 1e-3) / 1e9
    emit(op="mul by a small operand (stencil kernel) [a7]", shape=f"{list(xs_)} x {list(ys_)}", ms=round(ms, 4),
         algorithmic_bytes=8 * (nx + nz), gbs=round(gbs, 1), frac_of_hbm=round(gbs / HBM, 3), hbm_peak=HBM,
         macs=genfer_b200.mul_macs(xs_, ys_, rs_), launches_per_call=launches)
    del Tb, Ts, bigs, small

# ---- recurrences ----------------------------------------------------------------------------------------
peak_fl, _ = ctx.fp64_peak_probe(0, 16384)
for cfg in args.rec_shape.split(","):
    n, d = map(int, cfg.split("x"))
    shape = (d,) * n
    rng = np.random.default_rng(n * 100 + d)
    a = rng.uniform(0.5, 1.5, shape) / d ** n
    a.flat[0] = 1.0
    b = rng.uniform(0.5, 1.5, shape) / d ** n
    b.flat[0] = 1.25
    A, B = TP.new(a, shape), TP.new(b, shape)
    macs = genfer_b200.mul_macs(shape, shape, shape)
    ops = [("mul  [a7]", lambda x, y: x * y, 1.0), ("div  [a9]", lambda x, y: x / y, 1.0), ("exp  [a10]", lambda x, y: x.exp(), 1.0),
           ("log  [a11]", lambda x, y: x.log(), 2.0), ("pow(5) [a12]", lambda x, y: x.pow(5), 5.0),
           ("subst_var(last, y) [a16]", lambda x, y: x.subst_var(n - 1, y), float(d))]
    for name, f, products in ops:
        ms, launches = timed([lambda: f(A, B)], reps=3, warm=1)
        rec = dict(op=name, shape=cfg, ms=round(ms, 3), launches_per_call=launches, product_equivalents=products,
                   tflops_product_equivalent=round(2 * macs * products / (ms * 1e-3) / 1e12, 3), fp64_peak_tflops=round(peak_fl / 1e12, 2))
        if args.cpu:
            from oracle import oracle as O
            oa, ob = O.TaylorPoly.new(a, shape), O.TaylorPoly.new(b, shape)
            t0 = time.perf_counter()
            r = f(oa, ob)
            rec["cpu_oracle_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            g = f(A, B).array()
            ref = r.array()
            rec["max_rel_err_vs_oracle"] = float(np.max(np.abs(g - ref) / np.maximum(np.abs(ref), 1e-300)))
        emit(**rec)
ctx.close()
