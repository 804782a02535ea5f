#!/bin/bash
# round 2: insert kernel rate against the table size (does an L2-resident table insert faster?)
mkdir -p gpurun_out
for r in 25000 50000 100000 200000 400000; do
  timeout 300 python bench.py --reads $r --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --multi-k 0 --steps 5 > gpurun_out/bench14_$r.json 2> gpurun_out/bench14_$r.err; echo "rc=$?"
  python - $r <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench14_{sys.argv[1]}.json").read().strip().splitlines()[-1])
occ = d["check"]["kminmer_occurrences_total"]
print("reads", sys.argv[1], "windows", occ, "insert ms", round(d["kernels_ms"]["insert"], 4), "G windows/s", round(occ / d["kernels_ms"]["insert"] / 1e6, 1), "n_solid", d["check"]["n_solid_total"], d["table_phase_ms_profiled_step_rank0"])
PY
done
