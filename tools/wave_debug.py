"""Developer tool: per-level timing of the device-resident recurrences (GTP_WAVE_DEBUG=1)."""
import os, sys
os.environ["GTP_WAVE_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genfer_b200
ctx = genfer_b200.Context(0)
genfer_b200.set_default_context(ctx)
for shape in ((32, 32, 32), (16, 16, 16, 16)):
    rng = np.random.default_rng(1)
    a = rng.uniform(0.5, 1.5, shape) / np.prod(shape); a.flat[0] = 1.0
    A = genfer_b200.TaylorPoly.new(a, shape)
    for rep in range(2):
        A.exp(); (A / (A + 0.25)); A.log()
ctx.close()
