"""End-to-end time per .sgcl on the GPU vs the CPU oracle (BASELINE metric 2; best of N, the protocol of the
reference's benchmarks/neurips2023/exact/bench.py:33).  usage: time_sgcl.py [--reps N] file.sgcl[:limit] ..."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfer_b200
from oracle import oracle as O

args = sys.argv[1:]
reps, cpu_reps = 3, 2
while args and args[0] in ("--reps", "--cpu-reps"):
    if args[0] == "--reps":
        reps = int(args[1])
    else:
        cpu_reps = int(args[1])
    args = args[2:]
ctx = genfer_b200.Context(0)
for a in args:
    path, _, lim = a.partition(":")
    src = open(path).read()
    opts = genfer_b200.parse_flags(src)
    limit = int(lim) if lim else opts["limit"]
    kw = dict(limit=limit, no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"], unroll=opts["unroll"])
    tg, to = [], []
    for _ in range(reps):
        l0 = ctx.launch_count
        t = time.perf_counter(); g = genfer_b200.run_sgcl(src, ctx=ctx, **kw); tg.append(time.perf_counter() - t)
        launches = ctx.launch_count - l0
    for _ in range(max(1, min(reps, cpu_reps))):
        t = time.perf_counter(); o = O.run_sgcl(src, **kw); to.append(time.perf_counter() - t)
    rel = abs(g.total - o.total) / abs(o.total) if o.total else abs(g.total)
    print(json.dumps({"program": os.path.relpath(path), "limit": limit, "gpu_s": round(min(tg), 4), "cpu_oracle_s": round(min(to), 4),
                      "gpu_launches": launches, "nodes_evaluated": g.nodes_evaluated, "byte_identical": g.report == o.report,
                      "Z_rel_err": rel}), flush=True)
ctx.close()
