#!/bin/bash
# round 2: is the insert bound by random DRAM sectors?  Same 53 M windows, 10x / 100x fewer distinct keys (table in L2)
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg --multi-k 0 --no-edges --steps 4 --warmup 3"
for G in 100 10 3; do
  timeout 300 python bench.py --genomes $G $COMMON > gpurun_out/ins3_g$G.json 2> gpurun_out/ins3_g$G.err
  python - $G <<'PY'
import json, sys
g = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ins3_g{g}.json").read().strip().splitlines()[-1])
    occ = d["check"]["kminmer_occurrences_total"]
    print(g, "genomes: insert ms", d["kernels_ms"]["insert"], "windows", occ, "Gwin/s", occ / d["kernels_ms"]["insert"] / 1e6, "solid", d["check"]["n_solid_total"], "step", d["ms_per_step"], d.get("table_phase_ms_profiled_step_rank0"))
except Exception as e:
    print(g, "failed", e, open(f"gpurun_out/ins3_g{g}.err").read()[-500:])
PY
done
