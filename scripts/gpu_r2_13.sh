#!/bin/bash
# round 2: does NUMA placement bound the host packer?  topology + the e2e leg pinned to either socket
mkdir -p gpurun_out
{ lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nvidia-smi topo -m; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status; } > gpurun_out/topo13.txt 2>&1
cat gpurun_out/topo13.txt | head -40
for node in 0 1; do
  cpus=$(cat /sys/devices/system/node/node$node/cpulist 2>/dev/null) || continue
  timeout 600 taskset -c $cpus python bench.py --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --multi-k 0 --steps 3 > gpurun_out/bench13_node$node.json 2> gpurun_out/bench13_node$node.err; echo "rc=$?"
  python - $node <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench13_node{sys.argv[1]}.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("node", sys.argv[1], "e2e", round(e["value"], 1), "packed", round(e["packed_host_input"]["value"], 1), e["last_host_batch"]["pack_gb_per_s"])
PY
done
