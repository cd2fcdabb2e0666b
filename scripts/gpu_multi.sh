#!/bin/bash
# Multi-GPU pass: at N=2 the 2-GPU parity test (log kept), then bench.py at N ranks (fused P2P gather and, unless
# SKIP_NCCL, ncclAllGather).  Everything lands in gpurun_out/<tag>_*.
N=${1:-2}; tag=${2:-run}
python -c "import __graft_entry__ as g; g.build()" || exit 1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -rs > gpurun_out/${tag}_pytest_gpu_multi_n2.txt 2>&1
  echo "pytest multi rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu_multi_n2.txt; tail -4 gpurun_out/${tag}_pytest_gpu_multi_n2.txt
fi
export JJ_CPU_SAMPLE=${JJ_CPU_SAMPLE:-32768}
export NCCL_DEBUG=INFO
run() { # name, extra env
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/${tag}_bench_n${N}_$1.json 2> gpurun_out/${tag}_bench_n${N}_$1.err
  echo "$1 rc=$?"
  grep -m1 -o "nranks [0-9]*" gpurun_out/${tag}_bench_n${N}_$1.err | head -1
  grep -c "NVLS\|P2P" gpurun_out/${tag}_bench_n${N}_$1.err | sed 's/^/nccl P2P\/NVLS lines: /'
  tail -c 300000 gpurun_out/${tag}_bench_n${N}_$1.err > gpurun_out/${tag}_bench_n${N}_$1.err.tail && mv gpurun_out/${tag}_bench_n${N}_$1.err.tail gpurun_out/${tag}_bench_n${N}_$1.err
}
run p2p "JJ_GATHER=p2p" 29517
[ -n "$SKIP_NCCL" ] || run nccl "JJ_GATHER=nccl" 29518
[ -z "$WITH_BYTES" ] || run p2p_bytes "JJ_GATHER=p2p JJ_OUT=bytes" 29519
for f in gpurun_out/${tag}_bench_n${N}_p2p.json gpurun_out/${tag}_bench_n${N}_nccl.json gpurun_out/${tag}_bench_n${N}_p2p_bytes.json; do [ -s "$f" ] || continue
python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["n_gpus"], "%.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"], d["config"]["collective"][:40], d["clocks"], "parity", d["parity_check"]["ok"], d["parity_check"]["mismatches"], d["parity_check"]["ranks_checked"])
PY
done
