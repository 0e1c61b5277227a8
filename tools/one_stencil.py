"""Developer tool: a few launches of the small-operand product (ncu / A-B target).
usage: one_stencil.py [fast_mul mode: 1 default, 131073 gather kernel, 262145 row-walking kernel for any row length] [case]   cases: 0 = [297,282,297] x [2,1,2], 1 = 16^6 x [2,1,2,1,1,2], 2 = [52]^4 x [2,2,1,2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, genfer_b200
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
case = int(sys.argv[2]) if len(sys.argv) > 2 else 0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = genfer_b200.Context(0, stream=stream.cuda_stream)
ctx.set_fast_mul(mode)
TP = genfer_b200.TaylorPoly
xs, ys = [((297, 282, 297), (2, 1, 2)), ((16,) * 6, (2, 1, 2, 1, 1, 2)), ((52,) * 4, (2, 2, 1, 2))][case]
rs = tuple(a + b - 1 for a, b in zip(xs, ys))
big = torch.rand(xs, dtype=torch.float64, device="cuda"); small = torch.rand(ys, dtype=torch.float64, device="cuda")
B = TP.from_device(big.data_ptr(), xs, rs, ctx); S = TP.from_device(small.data_ptr(), ys, rs, ctx)
for _ in range(4):
    Z = B * S
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    Z = B * S
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
nbytes = (big.numel() + Z.array_shape()[0] * 0 + 1) * 8
import math
out_elems = math.prod(rs)
print(f"case {case} mode {mode} ctas {os.environ.get('GTP_DIRECT_CTAS', '-')}: ms per product {ms:.4f}  algorithmic {(big.numel() + out_elems) * 8 / ms / 1e9:.2f} TB/s")
ctx.close()
