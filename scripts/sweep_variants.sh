#!/bin/bash
# Times scalar-mul kernel mappings (jj_set_scalar_mul_variant) on 2^20 units.  The experimental mappings live in a
# -DJJ_EXPERIMENTS build of the same library: JJ_LIB=jubjub_b200/libjubjub_b200_exp.so (built by hand).
mkdir -p gpurun_out
out=gpurun_out/${1:-sweep}_variants.txt
: > $out
for v in ${VARIANTS:-13 24 5}; do
  timeout 120 python scripts/run_smul.py --logn 20 --variant $v --reps 4 2>&1 | tail -1 >> $out
done
cat $out
