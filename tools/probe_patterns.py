import sys; sys.path.insert(0,'.')
import genfer_b200
ctx = genfer_b200.Context(0)
for kind,name,it in ((0,'dfma',65536),(1,'dmma',16384),(2,'rowconv 16w/SM',4096),(4,'rowconv 8w/SM',4096),(3,'block22 8w/SM',1024),(5,'block22 12w/SM',1024)):
    fl,ms = ctx.fp64_peak_probe(kind, it)
    print(f"{name:18s} {fl/1e12:7.2f} TF/s  {ms:.2f} ms")
