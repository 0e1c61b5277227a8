"""Small-tensor operator mix for per-kernel duration measurements under ncu (launch-bound regime):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/time_small_ops.py; python tools/ncu_launch_agg.py L.csv"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, genfer_b200
ctx = genfer_b200.Context(0); genfer_b200.set_default_context(ctx)
TP = genfer_b200.TaylorPoly
for shape in [(27, 27), (20, 20, 20)]:
    a = TP.new(np.random.rand(*shape), shape); b = TP.new(np.random.rand(*shape), shape)
    s_dev = (a * b).coefficients_of_term(0, 0)   # some device tensor
    for i in range(30):
        r1 = a + b; r2 = a * 0.75; r3 = a - 0.5; r4 = -a; r5 = a + 0.25
ctx.synchronize(); ctx.close()
