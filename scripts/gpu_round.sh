#!/bin/bash
# One GPU-box pass for a milestone: parity tests, bench lines, per-kernel throughput, ncu launch list and
# a full ncu capture of the dominant kernel.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-run}
python -c "import __graft_entry__ as g; g.build()" || exit 1  # content-hash check: rebuilds only if the sources changed
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -x -q -m gpu -rs --durations=15 > gpurun_out/${tag}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.txt
tail -25 gpurun_out/${tag}_pytest_gpu.txt
fi
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python scripts/bench_ops.py > gpurun_out/${tag}_bench_ops.txt 2>&1; cp gpurun_out/bench_ops.json gpurun_out/${tag}_bench_ops.json 2>/dev/null; tail -32 gpurun_out/${tag}_bench_ops.txt
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scalar_mul -s 1 -c 1 -o gpurun_out/${tag}_smul_prof python scripts/run_smul.py --logn 20 --reps 3 > /dev/null 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_is_torsion_free -s 1 -c 1 -o gpurun_out/${tag}_torsion_prof python scripts/run_smul.py --logn 20 --reps 3 --what torsion > /dev/null 2>&1; echo "ncu torsion rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scalar_mul_fixed_gmem -s 1 -c 1 -o gpurun_out/${tag}_fixed_prof python scripts/run_smul.py --logn 20 --reps 3 --what fixed > /dev/null 2>&1; echo "ncu fixed rc=$?"
fi
cat gpurun_out/${tag}_bench_n1.json | cut -c1-900
if [ -n "$SANITIZE" ]; then
  for tool in memcheck racecheck; do
    timeout 400 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/${tag}_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/${tag}_sanitizer_$tool.txt
  done
fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
