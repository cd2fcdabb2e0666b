#!/bin/bash
# N-GPU A/B of the host NUMA binding of bench.py's ranks (JJ_NUMA_BIND=0 / 1), fused gather only; topology of the box first.
N=${1:-8}; tag=${2:-run}
python -c "import __graft_entry__ as g; g.build()" || exit 1
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; lscpu | grep -i "numa\|socket\|^CPU(s)\|model name"; echo; echo "cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.cpus 2>/dev/null)"; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"; for d in /sys/bus/pci/devices/*; do [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ] && echo "$(basename $d) numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; done; } > gpurun_out/${tag}_topology.txt 2>&1
export JJ_CPU_SAMPLE=${JJ_CPU_SAMPLE:-32768}
for b in 0 1; do
  JJ_NUMA_BIND=$b JJ_GATHER=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + b)) \
    bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/${tag}_bench_n${N}_p2p_numa$b.json 2> gpurun_out/${tag}_bench_n${N}_p2p_numa$b.err
  echo "bind=$b rc=$?"
  tail -c 100000 gpurun_out/${tag}_bench_n${N}_p2p_numa$b.err > gpurun_out/x.tail && mv gpurun_out/x.tail gpurun_out/${tag}_bench_n${N}_p2p_numa$b.err
  python - gpurun_out/${tag}_bench_n${N}_p2p_numa$b.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["n_gpus"], "%.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "parity", d["parity_check"]["ok"], d["parity_check"]["ranks_checked"])
print([ (x.get("bound"), x.get("numa_node"), x.get("usable"), x.get("why")) for x in d["config"].get("host_numa_binding", [])])
PY
done
cat gpurun_out/${tag}_topology.txt | head -40
