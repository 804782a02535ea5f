#!/bin/bash
# round 2: is the insert pass bound by DRAM's random-access rate?  error-free reads of few genomes => a table of a few MB (L2-resident)
mkdir -p gpurun_out
for cfg in "400000 2 0" "400000 20 0" "400000 100 0" "400000 100 0.001"; do
  set -- $cfg
  timeout 300 python bench.py --reads $1 --genomes $2 --err $3 --no-e2e --no-ascii-leg --no-cpu-baseline --extras '' --no-autotune --no-edges --multi-k 0 --steps 5 > gpurun_out/bench17_$2_$3.json 2> gpurun_out/bench17_$2_$3.err; echo "rc=$?"
  python - $2 $3 <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench17_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
occ = d["check"]["kminmer_occurrences_total"]
print("genomes", sys.argv[1], "err", sys.argv[2], "windows", occ, "insert ms", round(d["kernels_ms"]["insert"], 4), "G windows/s", round(occ / d["kernels_ms"]["insert"] / 1e6, 1), "n_solid", d["check"]["n_solid_total"], d["table_phase_ms_profiled_step_rank0"])
PY
done
