"""Quick kernel timing of the raw product on synthetic cubes (developer tool, not the bench).
usage: time_mul.py [--mode M] NxD ...   (mode 1 = auto, 2 = prefer the generic blocked kernel)"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import genfer_b200

args = sys.argv[1:]
modes = [1]
if args and args[0] == "--mode":
    modes = [int(m) for m in args[1].split(",")]
    args = args[2:]
cfgs = [tuple(map(int, a.split("x"))) for a in args] or [(4, 32), (5, 16), (6, 12), (5, 24), (6, 16)]
tstream = torch.cuda.Stream()
torch.cuda.set_stream(tstream)
ctx = genfer_b200.Context(0, stream=tstream.cuda_stream)
for n, d in cfgs:
    for mode in modes:
        ctx.set_fast_mul(mode)
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
        print(json.dumps({"n": n, "d": d, "mode": mode, "kind": kind, "ms": round(ms, 4), "tflops": round(2 * macs / ms / 1e9, 3)}), flush=True)
ctx.close()
