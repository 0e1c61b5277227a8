#!/bin/bash
# Final validation round without the ncu passes: tools/gpu_final.sh TAG
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/pytest.log; cat $out/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-300 $out/bench.json
timeout 300 python tools/time_mul.py 4x32 5x16 6x12 5x24 6x16 3x32 3x64 4x16 4x24 5x12 > $out/sweep.jsonl 2>&1; cat $out/sweep.jsonl
timeout 300 python tools/time_ops.py --cpu > $out/time_ops.jsonl 2>&1; grep -E "stencil|mul_linear|shift_down\(last, D" $out/time_ops.jsonl | cut -c1-260
(cd tests/golden/sgcl && timeout 700 python ../../../tools/time_sgcl.py --reps 2 --cpu-reps 1 config/example.sgcl:25 real_world/population2000.sgcl \
    slow/two_populations2000.sgcl real_world/hmm.sgcl slow/mixture.sgcl real_world/switchpoint.sgcl slow/population_100_2vars.sgcl \
    slow/population_50_3vars.sgcl:80 slow/population_50_4vars.sgcl:60 slow/population_50_3vars.sgcl:150:probs slow/population_50_4vars.sgcl:60:probs \
    slow/nested_infer_expensive.sgcl config/monty_hall.sgcl config/burglar_alarm.sgcl > ../../../$out/time_sgcl.jsonl 2>&1
 timeout 200 python ../../../tools/time_sgcl.py --reps 2 --cpu-reps 0 slow/population_50_3vars.sgcl:300:probs >> ../../../$out/time_sgcl.jsonl 2>&1); cut -c1-260 $out/time_sgcl.jsonl
ls -la $out
