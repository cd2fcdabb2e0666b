#!/bin/bash
# Times every selectable scalar-mul kernel mapping (jj_set_scalar_mul_variant) on 2^20 units.
mkdir -p gpurun_out
out=gpurun_out/${1:-sweep}_variants.txt
: > $out
for v in 13 2 3 4 5 9 10 11 12 14 15 16 19 20 21 22 23 1 6; do
  timeout 120 python scripts/run_smul.py --logn 20 --variant $v --reps 3 2>&1 | tail -1 >> $out
done
cat $out
