#!/bin/bash
# Full ncu capture of the product kernel on the given cubes: tools/ncu_mul.sh TAG 6x16 5x24 ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for cfg in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_mul_(blk|slide)' -s 1 -c 1 -f -o $out/prof_$cfg \
      python tools/time_mul.py $cfg > $out/ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
done
ls -la $out
