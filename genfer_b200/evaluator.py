"""Python front of the host evaluator (gtp_run_sgcl): run an SGCL program end to end on the GPU.

Mirrors what the reference's CLI does for the default f64 Taylor mode (/root/reference/src/main.rs:108-227).  The
parser, GF translation, DAG evaluation and report live in C++ (genfer_b200/csrc/evaluator/), every Taylor-polynomial
operation goes through the CUDA library; this module only marshals the result.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

from . import _lib
from .taylor import Context, TaylorPanic, default_context

MOMENT_NAMES = ("total", "mean", "raw2", "raw3", "raw4", "stddev", "variance", "central3", "central4", "skewness",
                "kurtosis")


class SgclResult:
    def __init__(self, report: str, moments: List[float], probs: List[float], normalized_probs: List[float],
                 is_normalized: bool, nodes_evaluated: int, cache_hits: int, moment_bounds=None, prob_bounds=None,
                 normalized_prob_bounds=None):
        self.moment_bounds = moment_bounds                      # 11 (lo, hi) pairs behind `moments`
        self.prob_bounds = prob_bounds                          # (lo, hi) of the unnormalised p(i)
        self.normalized_prob_bounds = normalized_prob_bounds    # (lo, hi) of p(i) / Z
        self.report = report
        self.moments = moments
        for name, value in zip(MOMENT_NAMES, moments):
            setattr(self, name, value)
        self.probs = probs
        self.normalized_probs = normalized_probs
        self.is_normalized = is_normalized
        self.nodes_evaluated = nodes_evaluated
        self.cache_hits = cache_hits


def parse_flags(source: str):
    """The `# flags: ...` first-line convention of the reference's test harness (tests/integration.rs:18-33)."""
    first = source.split("\n", 1)[0]
    opts = {"limit": None, "no_probs": False, "no_simplify_gf": False, "unroll": 8, "bounds": False, "symbolic": False, "unsupported": []}
    if "flags:" not in first:
        return opts
    toks = first.split("flags:", 1)[1].split()
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "--no-probs":
            opts["no_probs"] = True
        elif t == "--no-simplify-gf":
            opts["no_simplify_gf"] = True
        elif t in ("--limit", "-l"):
            opts["limit"] = int(toks[i + 1]); i += 1
        elif t in ("--unroll", "-u"):
            opts["unroll"] = int(toks[i + 1]); i += 1
        elif t == "--bounds":
            opts["bounds"] = True
        elif t in ("-s", "--symbolic"):
            opts["symbolic"] = True
        else:   # --rational, --precision, --big-float: number modes that stay on the reference's CPU code
            opts["unsupported"].append(t)
        i += 1
    return opts


def run_sgcl(source: str, limit: Optional[int] = None, no_probs: bool = False, no_simplify_gf: bool = False,
             unroll: int = 8, ctx: Optional[Context] = None, bounds: bool = False, symbolic: bool = False) -> SgclResult:
    """`bounds`: the reference's `--bounds` mode -- the evaluator runs over TaylorPoly<Interval<F64>> on the device (gti_*) and
    the report prints the enclosures.  `symbolic`: the reference's `-s` mode -- the generating function becomes one univariate
    computation DAG (host) that is evaluated over TaylorExpansion<F64> on the device (gtu_*)."""
    ctx = ctx or default_context()
    lib = ctx.lib
    h = C.c_void_p()
    err = C.create_string_buffer(2048)
    flags = (1 if no_probs else 0) | (2 if no_simplify_gf else 0) | (4 if bounds else 0) | (8 if symbolic else 0)
    rc = lib.gtp_run_sgcl(ctx.h, source.encode(), -1 if limit is None else int(limit), flags, unroll, C.byref(h), err, 2048)
    if rc != 0:
        raise TaylorPanic(rc, err.value.decode())
    try:
        m = (C.c_double * 11)()
        lib.gtp_sgcl_moments(h, m)
        n = int(lib.gtp_sgcl_limit(h))
        p, q = (C.c_double * max(n, 1))(), (C.c_double * max(n, 1))()
        lib.gtp_sgcl_probs(h, p, q)
        nodes, hits = C.c_uint64(), C.c_uint64()
        lib.gtp_sgcl_stats(h, C.byref(nodes), C.byref(hits))
        mb, pb, qb = (C.c_double * 22)(), (C.c_double * max(2 * n, 2))(), (C.c_double * max(2 * n, 2))()
        lib.gtp_sgcl_moment_bounds(h, mb)
        lib.gtp_sgcl_prob_bounds(h, pb, qb)
        return SgclResult(lib.gtp_sgcl_report(h).decode(), list(m), list(p)[:n], list(q)[:n],
                          bool(lib.gtp_sgcl_is_normalized(h)), int(nodes.value), int(hits.value),
                          [(mb[2 * i], mb[2 * i + 1]) for i in range(11)], [(pb[2 * i], pb[2 * i + 1]) for i in range(n)],
                          [(qb[2 * i], qb[2 * i + 1]) for i in range(n)])
    finally:
        lib.gtp_sgcl_free(h)
