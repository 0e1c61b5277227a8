#!/usr/bin/env python
"""Collects the reference's end-to-end fixtures for the f64 Taylor mode into tests/golden/sgcl/.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_sgcl_fixtures.py

Copied verbatim (test DATA, not source code): every `<name>.sgcl` + `<name>.expect` pair of the reference's golden-file
suites (tests/integration.rs:93-154: test/expect/{sample,observe,if,assign,while,normalize,examples,former_bugs,
real_world,slow} and benchmarks/neurips2023/{approx,exact}) whose `# flags:` line selects the default number type and
evaluator -- i.e. no --rational / --precision / -s / --bounds / --big-float (those modes stay on the reference's CPU code)
and no `skip integration test`.  `.expect` is the reference's stdout with --no-timing.
Also copied: example.sgcl (BASELINE config C1) and the benchmarks/prodigy programs of config C5 (they have no .expect;
their "Original code" comments quote exact rationals, checked in tests/test_sgcl_*.py).
The `-s` (symbolic mode) fixtures go to tests/golden/sgcl_symbolic/.
"""
import glob
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "sgcl")
OTHER_MODES = {"--rational", "-r", "--precision", "-p", "-s", "--symbolic", "--bounds", "-b", "--big-float"}


def wanted(path):
    first = open(path).readline()
    if "skip integration test" in first:
        return False
    if "flags:" in first and OTHER_MODES & set(first.split("flags:", 1)[1].split()):
        return False
    return True


def main():
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    n = 0
    seen = set()
    suites = [(f"test/expect/{d}", d) for d in ("sample", "observe", "if", "assign", "while", "normalize", "examples",
                                                "former_bugs", "real_world", "slow")]
    suites += [("benchmarks/neurips2023/approx", "bench_approx"), ("benchmarks/neurips2023/exact", "bench_exact")]
    for rel, name in suites:
        for src in sorted(glob.glob(os.path.join(REF, rel, "**", "*.sgcl"), recursive=True)):
            exp = src[:-5] + ".expect"
            if not os.path.exists(exp) or not wanted(src):
                continue
            key = (open(src).read(), open(exp).read())
            if key in seen:          # the benchmark trees repeat some real_world / slow programs
                continue
            seen.add(key)
            dst_dir = os.path.join(OUT, name)
            os.makedirs(dst_dir, exist_ok=True)
            base = os.path.basename(src)
            shutil.copy(src, os.path.join(dst_dir, base))
            shutil.copy(exp, os.path.join(dst_dir, base[:-5] + ".expect"))
            n += 1
    os.makedirs(os.path.join(OUT, "config"), exist_ok=True)
    shutil.copy(os.path.join(REF, "example.sgcl"), os.path.join(OUT, "config", "example.sgcl"))
    for prog in ("burglar_alarm", "max", "monty_hall", "monty_hall_nested", "grass", "fuzzy_or"):
        shutil.copy(os.path.join(REF, "benchmarks", "prodigy", prog + ".sgcl"), os.path.join(OUT, "config", prog + ".sgcl"))
    print(f"{n} fixture pairs + 7 config programs -> {OUT}")
    symbolic()


def symbolic():
    """The reference's `-s` (symbolic mode) fixtures, kept apart from the f64 Taylor-mode ones: tests/golden/sgcl_symbolic/."""
    out = os.path.join(HERE, "sgcl_symbolic")
    if os.path.exists(out):
        shutil.rmtree(out)
    os.makedirs(out)
    n = 0
    for src in sorted(glob.glob(os.path.join(REF, "test", "expect", "**", "*.sgcl"), recursive=True)):
        first = open(src).readline()
        exp = src[:-5] + ".expect"
        if "skip integration test" in first or "flags:" not in first or not os.path.exists(exp):
            continue
        flags = set(first.split("flags:", 1)[1].split())
        if not ({"-s", "--symbolic"} & flags) or (flags & (OTHER_MODES - {"-s", "--symbolic"})):
            continue
        shutil.copy(src, os.path.join(out, os.path.basename(src)))
        shutil.copy(exp, os.path.join(out, os.path.basename(exp)))
        n += 1
    print(f"{n} symbolic-mode fixture pairs -> {out}")


if __name__ == "__main__":
    main()
