"""Developer tool: per-kernel launch histogram and host-time split of .sgcl programs (GTP_LAUNCH_HIST=1 prints at context teardown)."""
import os, sys, time
os.environ["GTP_LAUNCH_HIST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genfer_b200
for a in sys.argv[1:]:
    path, _, lim = a.partition(":")
    src = open(path).read()
    opts = genfer_b200.parse_flags(src)
    ctx = genfer_b200.Context(0)
    kw = dict(limit=int(lim) if lim else opts["limit"], no_probs=opts["no_probs"], no_simplify_gf=opts["no_simplify_gf"], unroll=opts["unroll"])
    genfer_b200.run_sgcl(src, ctx=ctx, **kw)     # warm
    l0 = ctx.launch_count
    t = time.perf_counter(); genfer_b200.run_sgcl(src, ctx=ctx, **kw); dt = time.perf_counter() - t
    sys.stderr.write(f"=== {path} limit={kw['limit']}: {dt*1e3:.3f} ms, {ctx.launch_count - l0} launches (histogram below counts BOTH runs)\n")
    sys.stderr.flush()
    ctx.close()
