#!/bin/bash
# round 2: (1) multi-k extra after the buffer-trade fix, (2) e2e against the host batch size
mkdir -p gpurun_out
timeout 600 python bench.py --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --steps 3 > gpurun_out/bench12_mk.json 2> gpurun_out/bench12_mk.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench12_mk.json").read().strip().splitlines()[-1])
m = d["multi_k"]
print("value", d["value"], "allocs in timed region", d["device_allocations_in_timed_region"], "multi_k", m["ms_total"], m["ms_per_k"][:4], "first sweep", m["ms_per_k_first_sweep"][:4], m["allocations_rank0"])
PY
for b in 262144 524288 1000000; do
  timeout 600 python bench.py --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --multi-k 0 --steps 3 --e2e-batch $b > gpurun_out/bench12_e2e_$b.json 2> gpurun_out/bench12_e2e_$b.err; echo "rc=$?"
  python - $b <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench12_e2e_{sys.argv[1]}.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("batch", sys.argv[1], "e2e", round(e["value"], 1), "packed", round(e["packed_host_input"]["value"], 1), e["last_host_batch"])
PY
done
