#!/bin/bash
# Static SASS instruction mix of one kernel: tools/sass_mix.sh genfer_b200/csrc/kernels_mul_slide.cu 'k_mul_slideILi16ELb0'
src=$1; pat=$2
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -cubin -o /tmp/sass_mix.cubin $src -Xptxas -v 2>&1 | grep -A2 "$pat" | grep -E "Used|spill" 
cuobjdump -sass /tmp/sass_mix.cubin | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on {print}' > /tmp/sass_mix.txt
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" /tmp/sass_mix.txt | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -14
