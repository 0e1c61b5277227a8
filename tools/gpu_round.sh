#!/bin/bash
# One gpurun call: GPU tests, the bench line, the sweep, per-operator and per-program timings, the ncu launch list of the
# bench command and one full capture of the product kernel.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r01h'
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/pytest.log; cat $out/pytest.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json
timeout 300 python tools/time_mul.py 4x32 5x16 6x12 5x24 6x16 3x32 3x64 4x16 4x24 5x12 > $out/sweep.jsonl 2>&1; cat $out/sweep.jsonl
timeout 300 python tools/time_ops.py --cpu > $out/time_ops.jsonl 2>&1; tail -3 $out/time_ops.jsonl | cut -c1-200
(cd tests/golden/sgcl && timeout 600 python ../../../tools/time_sgcl.py --reps 2 --cpu-reps 1 config/example.sgcl:25 real_world/population2000.sgcl \
    slow/two_populations2000.sgcl real_world/hmm.sgcl slow/mixture.sgcl real_world/switchpoint.sgcl slow/population_100_2vars.sgcl \
    slow/population_50_3vars.sgcl:80 slow/population_50_4vars.sgcl:60 slow/nested_infer_expensive.sgcl config/monty_hall.sgcl \
    config/burglar_alarm.sgcl > ../../../$out/time_sgcl.jsonl 2>&1); cat $out/time_sgcl.jsonl | cut -c1-260
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_mul_(blk|slide)' -s 1 -c 1 -f -o $out/prof_product \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-aux > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out
