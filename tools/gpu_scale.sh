#!/bin/bash
# Strong-scaling run of the bench on one box: tools/gpu_scale.sh TAG "2 4 8"   (under gpurun --gpus 8)
tag=${1:-r01}; ns=${2:-"2 4 8"}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $out/smi_scale.txt 2>&1
port=29511
for n in $ns; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err
  echo "N=$n rc=$?"; cat $out/bench_${n}gpu.json | cut -c1-400
  port=$((port + 1))
done
