#!/bin/bash
# Kernel experiments: run the scalar-mul timing loop against alternative builds of the SAME library
# (JJ_LIB override).  usage: exp_libs.sh "<variants>" lib...   Output -> gpurun_out/exp_libs.txt
mkdir -p gpurun_out
out=gpurun_out/exp_libs.txt
: > $out
variants=$1; shift
for lib in "$@"; do
  for v in $variants; do
    echo "== $lib variant $v" >> $out
    JJ_LIB=$PWD/$lib timeout 120 python scripts/run_smul.py --logn 20 --variant $v --reps 4 2>&1 | tail -2 >> $out
  done
done
cat $out
