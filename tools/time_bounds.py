"""Wall time of the interval-enclosure run (gtp_run_sgcl_bounds: host evaluator over TaylorPoly<Interval<F64>> on the device)
next to the oracle's interval instantiation on one host core, with a check that the two enclosures overlap and the f64 result of
the GPU lies inside the GPU's enclosure.  One JSON line per program.

    python tools/time_bounds.py [--cpu-budget SECONDS] > profiles/rNN_bounds.jsonl
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SGCL = os.path.join(ROOT, "tests", "golden", "sgcl")

# (label, fixture, limit override, oracle interval run expected to fit the default budget)
PROGRAMS = [
    ("example --limit 25", "config/example.sgcl", 25, True),
    ("burglar_alarm", "config/burglar_alarm.sgcl", None, True),
    ("population2000", "real_world/population2000.sgcl", None, True),
    ("population_modified2000", "real_world/population_modified2000.sgcl", None, True),
    ("two_populations2000", "slow/two_populations2000.sgcl", None, True),
    ("hmm", "real_world/hmm.sgcl", None, True),
    ("population_50_3vars --limit 60", "slow/population_50_3vars.sgcl", 60, True),
    ("population_50_3vars --limit 120", "slow/population_50_3vars.sgcl", 120, False),
    ("population_50_4vars --limit 30", "slow/population_50_4vars.sgcl", 30, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="skip the oracle run of programs marked as long unless >= 600")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import genfer_b200
    from genfer_b200.interval import run_sgcl_bounds
    from helpers import check_inside_enclosure
    from oracle import oracle as O
    ctx = genfer_b200.Context(0)
    for label, rel, limit, cheap in PROGRAMS:
        if args.only and args.only not in label:
            continue
        path = os.path.join(SGCL, rel)
        if not os.path.exists(path):
            print(json.dumps({"program": label, "error": "fixture missing"}), flush=True)
            continue
        src = open(path).read()
        opts = genfer_b200.parse_flags(src)
        lim = limit if limit is not None else opts["limit"]
        g = genfer_b200.run_sgcl(src, limit=lim, no_probs=False, no_simplify_gf=opts["no_simplify_gf"], unroll=opts["unroll"], ctx=ctx)
        n = len(g.probs)
        tg, launches, b = [], 0, None
        for _ in range(2):
            l0 = ctx.launch_count
            t = time.perf_counter()
            b = run_sgcl_bounds(src, limit=n, unroll=opts["unroll"], ctx=ctx)
            tg.append(time.perf_counter() - t)
            launches = ctx.launch_count - l0
        checked = check_inside_enclosure(g, b)
        rec = {"program": label, "limit": n, "gpu_bounds_s": min(tg), "gpu_launches": int(launches), "f64_inside_enclosure": checked,
               "Z": list(b.total), "Z_width": b.total[1] - b.total[0],
               "max_p_width": max((hi - lo for lo, hi in b.probs if math.isfinite(hi - lo)), default=0.0)}
        if cheap or args.cpu_budget >= 600:
            t = time.perf_counter()
            o = O.run_sgcl_bounds(src, limit=n, unroll=opts["unroll"])
            rec["cpu_oracle_bounds_s"] = time.perf_counter() - t
            rec["speedup"] = rec["cpu_oracle_bounds_s"] / rec["gpu_bounds_s"]
            overlap = all(max(a[0], c[0]) <= min(a[1], c[1]) for a, c in zip([b.total] + b.raw_moments + b.probs, [o.total] + o.raw_moments + o.probs)
                          if all(math.isfinite(x) for x in a + c))
            rec["overlaps_oracle"] = bool(overlap)
            rec["oracle_Z_width"] = o.total[1] - o.total[0]
            rec["oracle_max_p_width"] = max((hi - lo for lo, hi in o.probs if math.isfinite(hi - lo)), default=0.0)
        print(json.dumps(rec), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
