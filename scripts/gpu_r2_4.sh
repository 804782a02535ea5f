#!/bin/bash
# round 2: (1) insert rate vs table size (is an L2-resident table region worth a partitioned insert?), (2) ncu of the unitig kernels
mkdir -p gpurun_out
COMMON="--no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg --multi-k 0 --no-edges --steps 5 --warmup 3"
for R in 12500 25000 50000 100000 200000 400000; do
  timeout 300 python bench.py --reads $R $COMMON > gpurun_out/ins_$R.json 2> gpurun_out/ins_$R.err
  python - $R <<'PY'
import json, sys
r = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ins_{r}.json").read().strip().splitlines()[-1])
    occ = d["check"]["kminmer_occurrences_total"]
    print(r, "reads: insert ms", d["kernels_ms"]["insert"], "windows", occ, "Gwin/s", occ / d["kernels_ms"]["insert"] / 1e6, "sketch ms", d["kernels_ms"]["sketch"], "step", d["ms_per_step"], "phases", d.get("table_phase_ms_profiled_step_rank0"))
except Exception as e:
    print(r, "failed", e)
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unitig_" -c 40 -f -o gpurun_out/prof_unitig python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras= --no-autotune --no-ascii-leg --multi-k 0 > gpurun_out/ncu_unitig.log 2>&1; echo "unitig capture rc=$?"
ls -la gpurun_out | tail -5
