"""Conditioning of the series division on tools/time_ops.py's inputs (x = 1 + tiny, y = 1.25 + tiny): the oracle's own per-coefficient
error against a long-double evaluation of the same recurrence, next to a double evaluation in another summation order.  The tiny
high-order coefficients are differences of nearly equal terms: per-coefficient relative error is ~1e-13 at 3x10 and grows with
the size, for ANY double-precision evaluation order -- the device result differs from the oracle by as much as the oracle differs
from the exact quotient (norm-wise all of them agree to 1e-16).  CPU only."""
import sys, itertools
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import oracle as O
def div_ld(x, y):
    # forward substitution in long double, wavefront order
    shape = x.shape; nd = len(shape)
    X = x.astype(np.longdouble); Y = y.astype(np.longdouble)
    R = np.zeros(shape, dtype=np.longdouble)
    idx = sorted(itertools.product(*[range(n) for n in shape]), key=sum)
    for k in idx:
        s = X[k]
        sl = tuple(slice(0, ki+1) for ki in k)
        Rb = R[sl]; Yb = Y[sl][tuple(slice(None,None,-1) for _ in k)]
        s = s - (Rb*Yb).sum() + R[k]*Y[(0,)*nd]   # R[k] is still 0 here
        R[k] = s / Y[(0,)*nd]
    return R
n, d = 3, int(sys.argv[1]) if len(sys.argv) > 1 else 10
shape=(d,)*n
rng=np.random.default_rng(n*100+d)
a=rng.uniform(0.5,1.5,shape)/d**n; a.flat[0]=1.0
b=rng.uniform(0.5,1.5,shape)/d**n; b.flat[0]=1.25
q=(O.TaylorPoly.new(a,shape)/O.TaylorPoly.new(b,shape)).array()
ex=div_ld(a,b)
rel=np.abs(q-ex.astype(np.float64))/np.abs(ex.astype(np.float64))
# the same recurrence in double with a different summation order (numpy pairwise sums)
def div_d(x,y):
    shape=x.shape; nd=len(shape); R=np.zeros(shape)
    for k in sorted(itertools.product(*[range(n) for n in shape]), key=sum):
        sl=tuple(slice(0,ki+1) for ki in k)
        s=x[k]-(R[sl]*y[sl][tuple(slice(None,None,-1) for _ in k)]).sum()
        R[k]=s/y[(0,)*nd]
    return R
q2=div_d(a,b)
rel2=np.abs(q2-ex.astype(np.float64))/np.abs(ex.astype(np.float64))
print("oracle vs long-double exact: max rel %.2e ; reordered double vs exact: %.2e ; reordered vs oracle %.2e"%(rel.max(), rel2.max(), (np.abs(q2-q)/np.abs(q)).max()))
print("normwise oracle err %.2e; min |q| %.2e max |q| %.2e"%(np.abs(q-ex).max()/np.abs(ex).max(), np.abs(ex).min(), np.abs(ex).max()))
