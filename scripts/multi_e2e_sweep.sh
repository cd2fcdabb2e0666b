#!/bin/bash
N=$1
mkdir -p gpurun_out
export JJ_CPU_SAMPLE=16384
for r in 1 2 4; do
  JJ_SHARD_ROUNDS=$r timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$r bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02i_n${N}_rounds$r.json 2> /dev/null
  python - gpurun_out/r02i_n${N}_rounds$r.json $r <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("JJ_SHARD_ROUNDS", sys.argv[2], "value %.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "parity", d["parity_check"]["ok"])
PY
done
