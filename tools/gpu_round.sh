#!/bin/bash
# One gpurun call: the bench line, the ncu launch list of the same command and one full capture of the product kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01c'
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/smi.txt 2>&1
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json
timeout 300 python tools/time_mul.py 4x32 5x16 6x12 5x24 6x16 > $out/sweep.jsonl 2>&1; cat $out/sweep.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_mul_(blk|slide)' -s 1 -c 1 -o $out/prof_product \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-aux > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out
