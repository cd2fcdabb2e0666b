#!/bin/bash
# Multi-GPU pass: bench.py at N ranks (fused P2P gather and, optionally, ncclAllGather); N=2 also runs the 2-GPU test.
N=${1:-2}; tag=${2:-run}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2; fi
export JJ_CPU_SAMPLE=32768
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n${N}_p2p.json 2> gpurun_out/${tag}_bench_n${N}_p2p.err; echo "p2p rc=$?"
[ -n "$SKIP_NCCL" ] || JJ_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n${N}_nccl.json 2> gpurun_out/${tag}_bench_n${N}_nccl.err; echo "nccl rc=$?"
for f in gpurun_out/${tag}_bench_n${N}_p2p.json gpurun_out/${tag}_bench_n${N}_nccl.json; do [ -s "$f" ] || continue
python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["n_gpus"], "%.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"], d["config"]["collective"][:40], d["clocks"])
PY
done
