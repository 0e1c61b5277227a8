#!/bin/bash
# Final validation round without the ncu passes: tools/gpu_final.sh TAG
tag=${1:-r02z}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/pytest.log; cat $out/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-300 $out/bench.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; echo "reference arm rc=$?"; cut -c1-300 $out/bench_reference.json
timeout 600 python tools/time_bounds.py > $out/bounds.jsonl 2> $out/bounds.err; cut -c1-220 $out/bounds.jsonl
timeout 200 python - > $out/symbolic.jsonl 2>&1 <<'PY'
import glob, json, os, time
import genfer_b200
from oracle import oracle as O
ctx = genfer_b200.Context(0)
for f in sorted(glob.glob("tests/golden/sgcl_symbolic/*.sgcl")):
    src = open(f).read(); exp = open(f[:-5] + ".expect").read()
    o = genfer_b200.parse_flags(src)
    kw = dict(limit=o["limit"], no_probs=o["no_probs"], no_simplify_gf=o["no_simplify_gf"], unroll=o["unroll"])
    tg, tc = [], []
    for _ in range(3):
        l0 = ctx.launch_count; t = time.perf_counter(); g = genfer_b200.run_sgcl(src, ctx=ctx, symbolic=True, **kw); tg.append(time.perf_counter() - t); n = ctx.launch_count - l0
        t = time.perf_counter(); c = O.run_sgcl(src, symbolic=True, **kw); tc.append(time.perf_counter() - t)
    print(json.dumps({"program": os.path.basename(f), "gpu_s": min(tg), "cpu_oracle_s": min(tc), "gpu_launches": n, "gpu_byte_identical": g.report == exp, "oracle_byte_identical": c.report == exp}), flush=True)
ctx.close()
PY
cat $out/symbolic.jsonl
(cd tests/golden/sgcl && timeout 700 python ../../../tools/time_sgcl.py --reps 3 --cpu-reps 1 config/example.sgcl:25 real_world/population2000.sgcl \
    slow/two_populations2000.sgcl real_world/hmm.sgcl real_world/switchpoint.sgcl slow/population_50_3vars.sgcl:120:probs slow/population_50_4vars.sgcl:50:probs \
    config/burglar_alarm.sgcl config/grass.sgcl > ../../../$out/time_sgcl.jsonl 2>&1
 timeout 200 python ../../../tools/time_sgcl.py --reps 3 --cpu-reps 0 slow/population_50_3vars.sgcl:300:probs slow/population_50_4vars.sgcl:60:probs >> ../../../$out/time_sgcl.jsonl 2>&1); cut -c1-260 $out/time_sgcl.jsonl
ls -la $out
