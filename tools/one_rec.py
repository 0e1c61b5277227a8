"""Developer tool (ncu target): a few launches of the device-resident recurrences, the fused Horner loop (row-staged) and the
axis-convolution kernel.  usage: one_rec.py [wave|horner|axis]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genfer_b200
what = sys.argv[1] if len(sys.argv) > 1 else "wave"
ctx = genfer_b200.Context(0)
genfer_b200.set_default_context(ctx)
TP = genfer_b200.TaylorPoly
rng = np.random.default_rng(1)
if what == "wave":
    shape = (32, 32, 32)
    a = rng.uniform(0.5, 1.5, shape) / np.prod(shape); a.flat[0] = 1.0
    A = TP.new(a, shape)
    for _ in range(3):
        A.exp(); (A / (A + 0.25)); A.log()
elif what == "horner":
    shape, deg = (120, 110, 100), (140, 130, 140)
    A = TP.new(rng.standard_normal(shape), deg)
    S = TP.new(rng.standard_normal((2, 1, 2)), deg)
    for _ in range(3):
        A.subst_var(2, S)
else:
    x = TP.new(rng.standard_normal((336, 1)), (336, 336)); y = TP.new(rng.standard_normal((336, 336)), (336, 336))
    z = TP.new(rng.standard_normal((1, 336)), (336, 336))
    for _ in range(3):
        x * y; z * y
ctx.synchronize()
ctx.close()
