#!/bin/bash
# Runs build/pcie_ceiling on N GPUs at once (one process per GPU, same start second) for each pinned-allocation flag set,
# with and without pinning each process to the CPUs `nvidia-smi topo -m` lists as local to its GPU.
#   tools/pcie_ceiling_all.sh N [mib_per_array]        -> JSON lines on stdout
N=${1:-1}; MIB=${2:-512}
cd "$(dirname "$0")/.."
nvidia-smi topo -m 2>/dev/null | sed 's/\x1b\[[0-9;]*m//g' > /tmp/topo.txt
echo "{\"topo\": $(python3 -c 'import json,sys; print(json.dumps(open("/tmp/topo.txt").read()))'), \"nproc\": $(nproc), \"numa_nodes\": \"$(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l)\"}"
for AFF in none local; do
  for FLAGS in default portable wc portable+wc; do
    START=$(python3 -c 'import time; print(time.time()+3)')
    for ((d=0; d<N; d++)); do
      CPUS=""
      if [ "$AFF" = local ]; then
        CPUS=$(python3 - "$d" <<'PY'
import re, sys
d = int(sys.argv[1])
for line in open("/tmp/topo.txt"):
    if line.startswith(f"GPU{d}\t") or line.startswith(f"GPU{d} "):
        m = re.findall(r"\b(\d+(?:-\d+)?(?:,\d+(?:-\d+)?)*)\b", line.split("\t", 1)[1] if "\t" in line else line)
        cand = [x for x in m if "-" in x or "," in x]
        print(cand[0] if cand else "")
        break
PY
)
      fi
      if [ -n "$CPUS" ]; then
        taskset -c "$CPUS" build/pcie_ceiling --device $d --flags $FLAGS --mib-per-array $MIB --start-at $START | sed "s/^{/{\"n_procs\": $N, \"affinity\": \"$CPUS\", /" &
      else
        build/pcie_ceiling --device $d --flags $FLAGS --mib-per-array $MIB --start-at $START | sed "s/^{/{\"n_procs\": $N, \"affinity\": \"none\", /" &
      fi
    done
    wait
  done
done
