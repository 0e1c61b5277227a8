#!/bin/bash
# compute-sanitizer over the kernels touched this round (small shapes): tools/sanitize.sh TAG
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_product.py \
    -k "tiled_sliding or ragged_leading or kernel_selection" > $out/memcheck_product.log 2>&1; echo "memcheck product rc=$?"
tail -4 $out/memcheck_product.log
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_taylor.py \
    -k "mul_linear or gathers or many_variables or maximum_rank or scalar or shift_down or binary_ops" > $out/memcheck_taylor.log 2>&1; echo "memcheck taylor rc=$?"
tail -4 $out/memcheck_taylor.log
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_sgcl.py -k "example or prodigy" > $out/memcheck_sgcl.log 2>&1; echo "memcheck sgcl rc=$?"
tail -4 $out/memcheck_sgcl.log
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest -q -x tests/test_gpu_product.py \
    -k "tiled_sliding and (4-8-34 or 4-12-34 or 5-8-34)" > $out/racecheck_product.log 2>&1; echo "racecheck product rc=$?"
tail -4 $out/racecheck_product.log
grep -c "ERROR SUMMARY" $out/*.log; grep -h "ERROR SUMMARY" $out/*check*.log | sort | uniq -c
