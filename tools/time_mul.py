"""Quick kernel timing of the raw product on synthetic cubes (developer tool, not the bench)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import genfer_b200

cfgs = [tuple(map(int, a.split("x"))) for a in sys.argv[1:]] or [(5, 16), (6, 12), (6, 16)]
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
ctx = genfer_b200.Context(0, stream=tstream.cuda_stream)
for n, d in cfgs:
    shape = (d,) * n
    x = torch.rand(shape, dtype=torch.float64, device="cuda")
    y = torch.rand(shape, dtype=torch.float64, device="cuda")
    z = torch.empty(shape, dtype=torch.float64, device="cuda")
    macs = genfer_b200.mul_macs(shape, shape, shape)
    kind = ctx.mul_kernel_kind(shape, shape, shape)
    reps = 3 if macs > 1e12 else 10
    for it in range(2):
        ctx.mul_rows_raw(shape, x.data_ptr(), shape, y.data_ptr(), shape, 0, 1, d, z.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(reps):
        ctx.mul_rows_raw(shape, x.data_ptr(), shape, y.data_ptr(), shape, 0, 1, d, z.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"n": n, "d": d, "kind": kind, "ms": ms, "tflops": 2 * macs / ms / 1e9}), flush=True)
ctx.close()
