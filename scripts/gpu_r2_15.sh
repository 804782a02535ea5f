#!/bin/bash
# round 2: insert kernel rate with an L2-resident table (few genomes => few distinct k-min-mers) against the same number of windows with a large table
mkdir -p gpurun_out
for cfg in "50000 2" "100000 2" "100000 100" "200000 2" "200000 100"; do
  set -- $cfg
  timeout 300 python bench.py --reads $1 --genomes $2 --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --multi-k 0 --steps 5 > gpurun_out/bench15_$1_$2.json 2> gpurun_out/bench15_$1_$2.err; echo "rc=$?"
  python - $1 $2 <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench15_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
occ = d["check"]["kminmer_occurrences_total"]
print("reads", sys.argv[1], "genomes", sys.argv[2], "windows", occ, "insert ms", round(d["kernels_ms"]["insert"], 4), "G windows/s", round(occ / d["kernels_ms"]["insert"] / 1e6, 1), "n_solid", d["check"]["n_solid_total"], d["table_phase_ms_profiled_step_rank0"])
PY
done
