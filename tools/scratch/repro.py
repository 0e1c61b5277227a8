import faulthandler, sys, os
faulthandler.enable()
sys.path.insert(0, os.getcwd())
import torch, genfer_b200
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = genfer_b200.Context(0, stream=stream.cuda_stream)
genfer_b200.set_default_context(ctx)
TP = genfer_b200.TaylorPoly
for shape in [(4,4),(16,)*4,(16,)*6]:
    x = torch.rand(shape, dtype=torch.float64, device="cuda")
    t = TP.from_device(x.data_ptr(), shape, shape, ctx)
    print("shape", shape, flush=True)
    r = t * 0.75
    print(" mul ok", flush=True)
    torch.cuda.synchronize()
    print(" sync ok", float((r.array().ravel()[1])), float(x.view(-1)[1])*0.75, flush=True)
    r2 = t * TP.from_scalar(0.75, ctx)
    torch.cuda.synchronize(); print(" explicit ok", flush=True)
