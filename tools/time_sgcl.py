"""End-to-end time per .sgcl on the GPU vs the CPU oracle (BASELINE metric 2; best of N, the protocol of the
reference's benchmarks/neurips2023/exact/bench.py:33).
usage: time_sgcl.py [--reps N] [--cpu-reps N] file.sgcl[:limit[:probs]] ...   (`:probs` overrides a `--no-probs` flag line;
--cpu-reps 0 skips the oracle)"""
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
    lim, _, probs = lim.partition(":")
    src = open(path).read()
    opts = genfer_b200.parse_flags(src)
    limit = int(lim) if lim else opts["limit"]
    kw = dict(limit=limit, no_probs=opts["no_probs"] and probs != "probs", no_simplify_gf=opts["no_simplify_gf"], unroll=opts["unroll"])
    tg, to = [], []
    for _ in range(reps):
        l0 = ctx.launch_count
        t = time.perf_counter(); g = genfer_b200.run_sgcl(src, ctx=ctx, **kw); tg.append(time.perf_counter() - t)
        launches = ctx.launch_count - l0
    o = None
    for _ in range(min(reps, cpu_reps)):
        t = time.perf_counter(); o = O.run_sgcl(src, **kw); to.append(time.perf_counter() - t)
    rec = {"program": os.path.relpath(path), "limit": limit, "no_probs": kw["no_probs"], "gpu_s": round(min(tg), 4),
           "gpu_launches": launches, "nodes_evaluated": g.nodes_evaluated}
    if o is not None:
        rel = abs(g.total - o.total) / abs(o.total) if o.total else abs(g.total)
        prel = max((abs(a - b) / abs(b) for a, b in zip(g.probs, o.probs) if b), default=0.0)
        rec.update({"cpu_oracle_s": round(min(to), 4), "byte_identical": g.report == o.report, "Z_rel_err": rel,
                    "max_prob_rel_err": prel})
    print(json.dumps(rec), flush=True)
ctx.close()
