set -x
GTP_DIRECT_MIN=1 GTP_FAST_MUL=262145 timeout 900 python -m pytest tests/test_gpu_taylor.py tests/test_gpu_product.py -m gpu -q -x -k "horner or stencil or STENCIL or row_staged or small_operand or row_walking or subst" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_taylor.py tests/test_gpu_sgcl.py -m gpu -q -x -k "horner or subst or golden or row_walking" 2>&1 | tail -3
for c in 0 2; do
  for m in 1 262145; do python tools/one_stencil.py $m $c 2>&1 | tail -1; done
done
S=tests/golden/sgcl/slow
python tools/time_sgcl.py --cpu-reps 0 $S/population_50_3vars.sgcl:300:probs $S/population_50_3vars.sgcl:120:probs $S/population_50_4vars.sgcl:60:probs $S/population_50_4vars.sgcl:50:probs $S/two_populations2000.sgcl 2>&1 | cut -c1-200
